"""Stage the UNMODIFIED reference package under the git-ignored `baseline/_ref/` so that it travels to the GPU box.

    python tools/stage_reference.py            # idempotent; no-op when /root/reference is absent (the GPU box)

The base contract's recipe is `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref
/root/reference`; it fails at metadata generation here ("Multiple top-level packages discovered in a flat-layout:
['config', 'assets', 'videox_fun']" — the reference's pyproject.toml names no package list), so this script does what
that install would have done: it places the `videox_fun` package (Python files only, 1.2 MB, byte-identical — a sha256
manifest is written next to it) under `baseline/_ref/`.  `baseline/_ref/` is listed in .gitignore (never enters the
history) and not in .gpurunignore (travels with the snapshot like the built .so files).

Who reads it: tools/ref_loader.py (the diffusers shim) when `/root/reference` is not there, i.e. on the GPU box, for
  * tools/gpu_reference.py — the reference's own CUDA path (bf16 autocast, flash-attn 2 / cuBLAS / cuDNN) timed and
    compared beside libvcof (bench.py's `gpu_reference` object, tests/test_widen_y_gpu_reference.py).
Nothing in the product (`videocof_b200/`, `videox_fun/` overlay) imports it.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("VCOF_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True):
    pkg = os.path.join(SRC, "videox_fun")
    if not os.path.isdir(pkg):
        if verbose:
            print(f"stage_reference: {SRC} not present; keeping {DST} as it is "
                  f"({'present' if os.path.isdir(DST) else 'absent'})")
        return os.path.isdir(os.path.join(DST, "videox_fun"))
    manifest = {}
    for dirpath, _dirs, files in os.walk(pkg):
        rel = os.path.relpath(dirpath, SRC)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(dirpath, f)
            d = os.path.join(DST, rel, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            data = open(s, "rb").read()
            manifest[os.path.join(rel, f)] = hashlib.sha256(data).hexdigest()
            if not os.path.exists(d) or open(d, "rb").read() != data:
                shutil.copyfile(s, d)
    lic = os.path.join(SRC, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(DST, "LICENSE"))
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=0, sort_keys=True)
    if verbose:
        print(f"stage_reference: {len(manifest)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
