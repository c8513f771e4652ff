"""Which kernels of the current build are instruction-for-instruction the kernels of an earlier revision?

    python tools/sass_diff.py 3e993bd            # 3e993bd = last revision whose build ran on a B200 in round 1

Compiles that revision's videocof_b200/csrc/*.cu (sm_100a, the build's own flags) into a scratch directory, dumps the SASS
of both builds and matches kernels by their instruction text (addresses and encodings stripped).  A kernel reported
IDENTICAL executes exactly the instruction sequence that was validated on hardware; SAME-CODE = the same opcodes,
immediates and order with some register numbers swapped (ptxas does that between two compilations of one and the same
source); SAME-INSTRUCTIONS = additionally a few independent neighbours scheduled in another order — the evidence behind "the default
paths did not change" for work done without a GPU (profiles/r1_sass_vs_validated.txt).  CPU only."""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "videocof_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fns, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = fns.setdefault(m.group(1), [])
            continue
        m = re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x", line)
        if m and cur is not None:
            cur.append(m.group(1).strip())
    return fns


def skeleton(body):
    """Instruction text with register / predicate numbers blanked: two compilations of one source can differ in the
    numbering of equivalent registers (ptxas picks between symmetric choices differently from run to run)."""
    return [re.sub(r"\b(UR|UP|R|P)\d+\b", r"\1#", ins) for ins in body]


def demangle(name):
    out = subprocess.run(["cu++filt", "-p", name], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\((bool|int)\)", "", out).replace("vcof::", "")


def main(rev):
    files = subprocess.run(["git", "-C", ROOT, "ls-tree", "--name-only", rev, "videocof_b200/csrc/"], capture_output=True,
                           text=True, check=True).stdout.split()
    print(f"# python tools/sass_diff.py {rev} — kernels of the working tree vs the build of {rev} (instruction text, "
          "addresses / encodings stripped)")
    with tempfile.TemporaryDirectory() as tmp:
        # same relative depth as the tree: the sources include "../../include/vcof.h"
        src = os.path.join(tmp, "videocof_b200", "csrc")
        os.makedirs(src)
        os.makedirs(os.path.join(tmp, "include"))
        for f in files + ["include/vcof.h"]:
            data = subprocess.run(["git", "-C", ROOT, "show", f"{rev}:{f}"], capture_output=True, check=True).stdout
            with open(os.path.join(tmp, f), "wb") as fh:
                fh.write(data)
        for f in sorted(files):
            if not f.endswith(".cu"):
                continue
            base = os.path.basename(f)
            new_obj = os.path.join(CSRC, base[:-3] + ".o")
            old_obj = os.path.join(src, base[:-3] + ".o")
            subprocess.run(["nvcc", *FLAGS, "-c", os.path.join(src, base), "-o", old_obj], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            old, new = sass(old_obj), sass(new_obj)
            exact = {tuple(v): k for k, v in old.items()}
            skel = {tuple(skeleton(v)): k for k, v in old.items()}
            bag = {tuple(sorted(skeleton(v))): k for k, v in old.items()}
            print(f"\n## {base}")
            for name in sorted(new, key=demangle):
                body = new[name]
                sk = skeleton(body)
                if tuple(body) in exact:
                    verdict, other = "IDENTICAL     ", exact[tuple(body)]
                elif tuple(sk) in skel:
                    other = skel[tuple(sk)]
                    k = sum(a != b for a, b in zip(body, old[other]))
                    verdict = f"SAME-CODE ({k} instr. with renamed registers)"
                elif tuple(sorted(sk)) in bag:
                    other = bag[tuple(sorted(sk))]
                    k = sum(a != b for a, b in zip(sk, skeleton(old[other])))
                    verdict = f"SAME-INSTRUCTIONS ({k} positions reordered)"
                else:
                    print(f"new/changed    {demangle(name)}  ({len(body)} instructions)")
                    continue
                print(f"{verdict} {demangle(name)}  ({len(body)} instructions) == {demangle(other)}")
        newer = sorted(set(os.path.basename(p) for p in os.listdir(CSRC) if p.endswith(".cu")) -
                       set(os.path.basename(f) for f in files))
        for base in newer:
            print(f"\n## {base}\nnew file (not in {rev})")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "HEAD"))
