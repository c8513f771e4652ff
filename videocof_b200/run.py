"""Run a reference CLI UNCHANGED on the B200-native hot path.

    python -m videocof_b200.run /path/to/VideoCoF/fast_infer.py --video_path ... --prompt ...
    torchrun --nproc_per_node=1 -m videocof_b200.run /path/to/VideoCoF/inference.py ...

Both reference CLIs put their own directory at the FRONT of sys.path before importing `videox_fun`
(fast_infer.py:14-22, inference.py:9-17), which shadows a PYTHONPATH overlay.  This launcher imports this repository's
`videox_fun` overlay first — a package already in sys.modules wins over any later sys.path edit — points the overlay at
the script's checkout for every module it does not override (VIDEOCOF_REFERENCE_ROOT), and then executes the script as
`__main__` with its own argv through runpy.  No reference file is edited; INTEGRATION.md §1.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_checkout(script):
    """The reference checkout a script belongs to: the nearest parent directory that holds a `videox_fun` package."""
    d = os.path.dirname(os.path.abspath(script))
    for _ in range(4):
        if os.path.isdir(os.path.join(d, "videox_fun", "models")) and os.path.abspath(d) != ROOT:
            return d
        d = os.path.dirname(d)
    return None


def install_overlay(script):
    """Make `import videox_fun` resolve to this repository for the rest of the process; returns the overlay module."""
    if "videox_fun" in sys.modules and not os.path.abspath(
            getattr(sys.modules["videox_fun"], "__file__", "") or "").startswith(ROOT):
        raise RuntimeError("videox_fun was imported from %r before the overlay could be installed"
                           % sys.modules["videox_fun"].__file__)
    checkout = os.environ.get("VIDEOCOF_REFERENCE_ROOT") or find_checkout(script)
    if checkout:
        os.environ["VIDEOCOF_REFERENCE_ROOT"] = checkout
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import videox_fun                                   # noqa: F401  the overlay, now pinned in sys.modules
    import videox_fun.models                            # noqa: F401  (libvcof-backed DiT / VAE / umT5)
    import videox_fun.pipeline                          # noqa: F401
    import videox_fun.utils                             # noqa: F401
    return sys.modules["videox_fun"]


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    script = argv[0]
    if not os.path.isfile(script):
        raise SystemExit(f"videocof_b200.run: no such script: {script}")
    install_overlay(script)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
