"""Static evidence from the built library, no GPU needed: per kernel of videocof_b200/csrc/libvcof.so the register /
stack use (`cuobjdump -res-usage`) and the count of the SASS mnemonics that prove which hardware path it takes
(`cuobjdump -sass`; /opt/skills/guides/B200_PROFILING.md lists them): UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA
load / store, LDTM / STTM = tcgen05.ld / .st (TMEM), HMMA = mma.sync (the text encoder's 64-wide heads), MUFU.EX2 =
the softmax exponentials, LDL / STL = local-memory traffic (spills or indexed local arrays).

    python tools/sass_report.py            # table on stdout (committed as profiles/r1_sass_summary.txt)

tests/test_sass_cpu.py asserts the properties DESIGN.md §3 claims for the hot kernels from the same data."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "videocof_b200", "csrc", "libvcof.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "MUFU.EX2", "LDL", "STL", "SYNCS")


def _tool(name):
    for cand in (shutil.which(name), os.path.join("/usr/local/cuda/bin", name)):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError(f"{name} not found")


def demangle(names):
    out = subprocess.run([_tool("cu++filt"), "-p"] + list(names), capture_output=True, text=True, check=True).stdout
    return [re.sub(r"\((bool|int)\)", "", line.strip()).replace("vcof::", "") for line in out.splitlines()]


def kernels(lib=LIB):
    """-> {demangled kernel name: {"REG": n, "STACK": n, "instructions": n, mnemonic: count, ...}}"""
    cuobjdump = _tool("cuobjdump")
    res = subprocess.run([cuobjdump, "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    info, cur = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            info[cur] = {"REG": int(m.group(1)), "STACK": int(m.group(2)), "instructions": 0,
                         **{k: 0 for k in MNEMONICS}}
            cur = None
    sass = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur = None
    op = re.compile(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = info.get(m.group(1))
            continue
        m = op.match(line)
        if m and cur is not None:
            cur["instructions"] += 1
            name = m.group(1)
            for k in MNEMONICS:
                if name == k or name.startswith(k + "."):
                    cur[k] += 1
    mangled = list(info)
    return dict(zip(demangle(mangled), (info[m] for m in mangled)))


def main():
    ks = kernels()
    cols = ("REG", "STACK", "instructions") + MNEMONICS
    print("# python tools/sass_report.py — static resource / SASS-mnemonic table of libvcof.so (sm_100a), no GPU involved")
    print("kernel," + ",".join(cols))
    for name in sorted(ks):
        print(f'"{name}",' + ",".join(str(ks[name][c]) for c in cols))
    return 0


if __name__ == "__main__":
    sys.exit(main())
