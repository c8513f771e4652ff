"""The `videox_fun` overlay package resolves every import the reference CLIs make (fast_infer.py:24-40,
inference.py:19-28): the hot-path modules to this repository, everything else to the reference checkout named by
VIDEOCOF_REFERENCE_ROOT.  Run against a stand-in checkout (three one-line modules) so the test does not depend on the
reference tree or on its third-party requirements being installed."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CLI_IMPORTS = {          # module -> names, as written in the two CLIs
    "videox_fun.models": ["AutoencoderKLWan", "WanT5EncoderModel", "AutoTokenizer", "WanTransformer3DModel"],
    "videox_fun.pipeline": ["WanPipeline"],
    "videox_fun.utils.fp8_optimization": ["convert_model_weight_to_float8", "replace_parameters_by_name",
                                          "convert_weight_dtype_wrapper"],
    "videox_fun.utils.lora_utils": ["merge_lora", "unmerge_lora"],
    "videox_fun.utils.utils": ["filter_kwargs", "save_videos_grid"],
    "videox_fun.data.dataset_image_video": ["derive_ground_object_from_instruction"],
    "videox_fun.utils.fm_solvers": ["FlowDPMSolverMultistepScheduler"],
    "videox_fun.utils.fm_solvers_unipc": ["FlowUniPCMultistepScheduler"],
}
OURS = {"videox_fun.models", "videox_fun.pipeline", "videox_fun.utils.lora_utils", "videox_fun.utils.utils",
        "videox_fun.utils.fm_solvers_unipc"}


def test_cli_imports_resolve(tmp_path):
    ref = tmp_path / "VideoCoF"
    for sub in ("models", "utils", "data", "pipeline", "dist"):
        (ref / "videox_fun" / sub).mkdir(parents=True)
        (ref / "videox_fun" / sub / "__init__.py").write_text("")
    (ref / "videox_fun" / "__init__.py").write_text("")
    (ref / "videox_fun" / "utils" / "fp8_optimization.py").write_text(textwrap.dedent("""
        def convert_model_weight_to_float8(*a, **k): return "ref"
        def replace_parameters_by_name(*a, **k): return "ref"
        def convert_weight_dtype_wrapper(*a, **k): return "ref"
    """))
    (ref / "videox_fun" / "utils" / "utils.py").write_text("def get_image_latent():\n    return 'ref'\n"
                                                            "def save_videos_grid():\n    return 'ref'\n")
    (ref / "videox_fun" / "utils" / "fm_solvers.py").write_text("class FlowDPMSolverMultistepScheduler: pass\n")
    (ref / "videox_fun" / "utils" / "lora_utils.py").write_text("def merge_lora(): return 'ref'\n")
    (ref / "videox_fun" / "data" / "dataset_image_video.py").write_text(
        "def derive_ground_object_from_instruction(s):\n    return 'ref:' + s\n")
    prog = textwrap.dedent(f"""
        import importlib, json
        out = {{}}
        for mod, names in {CLI_IMPORTS!r}.items():
            m = importlib.import_module(mod)
            out[mod] = [m.__file__, [n for n in names if not hasattr(m, n)]]
        import videox_fun.utils.utils as u
        out["forwarded_helper"] = u.get_image_latent()
        out["own_save"] = u.save_videos_grid.__module__
        print(json.dumps(out))
    """)
    env = dict(os.environ, PYTHONPATH=ROOT, VIDEOCOF_REFERENCE_ROOT=str(ref))
    r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    for mod in CLI_IMPORTS:
        path, missing = out[mod]
        assert not missing, (mod, missing)
        assert path.startswith(ROOT if mod in OURS else str(ref)), (mod, path)
    assert out["forwarded_helper"] == "ref"                      # un-overridden helpers come from the reference file
    assert out["own_save"] == "videocof_b200.video_io"           # the byte-aware save_videos_grid is ours
