"""`python -m videocof_b200.run <reference CLI>`: the reference scripts run UNCHANGED on the overlay although they put
their own checkout at sys.path[0] (fast_infer.py:14-22).  The stand-in script below starts with the reference's own
path-insertion header, verbatim in behaviour, then makes the CLIs' imports (fast_infer.py:24-40) and reports where each
one resolved; diffusers / omegaconf / imageio are stubbed the way the CLIs import them."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEADER = '''
import os
import sys
import json
import argparse

import numpy as np
import torch
from diffusers import FlowMatchEulerDiscreteScheduler
from omegaconf import OmegaConf
import imageio

current_file_path = os.path.abspath(__file__)
project_roots = [
    os.path.dirname(current_file_path),
    os.path.dirname(os.path.dirname(current_file_path)),
    os.path.dirname(os.path.dirname(os.path.dirname(current_file_path))),
]
for project_root in project_roots:
    if project_root not in sys.path:
        sys.path.insert(0, project_root)

from videox_fun.models import (AutoencoderKLWan, WanT5EncoderModel, AutoTokenizer, WanTransformer3DModel)
from videox_fun.pipeline import WanPipeline
from videox_fun.utils.fp8_optimization import replace_parameters_by_name
from videox_fun.utils.lora_utils import merge_lora, unmerge_lora
from videox_fun.utils.utils import filter_kwargs, save_videos_grid
from videox_fun.utils.fm_solvers_unipc import FlowUniPCMultistepScheduler
'''


def _fake_checkout(tmp_path):
    ref = tmp_path / "VideoCoF"
    for sub in ("models", "utils", "pipeline"):
        (ref / "videox_fun" / sub).mkdir(parents=True)
        (ref / "videox_fun" / sub / "__init__.py").write_text("raise ImportError('the reference package must not win')\n")
    (ref / "videox_fun" / "__init__.py").write_text("REFERENCE = True\n")
    (ref / "videox_fun" / "utils" / "fp8_optimization.py").write_text("def replace_parameters_by_name(*a): return 'ref'\n")
    (ref / "videox_fun" / "utils" / "utils.py").write_text("def filter_kwargs(c, k): return k\n"
                                                            "def save_videos_grid(*a, **k): return 'ref'\n")
    stubs = tmp_path / "stubs"
    for name, body in (("diffusers", "class FlowMatchEulerDiscreteScheduler: pass\n"),
                       ("omegaconf", "class OmegaConf: pass\n"), ("imageio", "")):
        (stubs / name).mkdir(parents=True)
        (stubs / name / "__init__.py").write_text(body)
    body = HEADER + textwrap.dedent('''
        ap = argparse.ArgumentParser()
        ap.add_argument("--prompt")
        args = ap.parse_args()
        import videox_fun
        print(json.dumps({"argv0": sys.argv[0], "prompt": args.prompt, "name": __name__,
                          "pkg": videox_fun.__file__, "dit": WanTransformer3DModel.__module__,
                          "vae": AutoencoderKLWan.__module__, "pipe": WanPipeline.__module__,
                          "sched": FlowUniPCMultistepScheduler.__module__, "merge": merge_lora.__module__,
                          "fp8": replace_parameters_by_name(), "path_head": sys.path[:3]}))
    ''')
    (ref / "fast_infer.py").write_text(body)
    return ref, stubs


def test_reference_cli_runs_unchanged_on_the_overlay(tmp_path):
    ref, stubs = _fake_checkout(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, str(stubs)]))
    env.pop("VIDEOCOF_REFERENCE_ROOT", None)
    r = subprocess.run([sys.executable, "-m", "videocof_b200.run", str(ref / "fast_infer.py"), "--prompt", "make it blue"],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["name"] == "__main__" and out["prompt"] == "make it blue" and out["argv0"].endswith("fast_infer.py")
    assert str(ref) in out["path_head"]                              # the script DID put its checkout first ...
    assert out["pkg"].startswith(ROOT)                               # ... and still got the overlay
    assert out["dit"] == "videocof_b200.dit" and out["vae"] == "videocof_b200.vae"
    assert out["pipe"] == "videocof_b200.pipeline" and out["sched"] == "videocof_b200.scheduler"
    assert out["merge"] == "videocof_b200.lora"
    assert out["fp8"] == "ref"                                       # un-overridden modules come from the checkout


def test_without_the_launcher_the_checkout_shadows_the_overlay(tmp_path):
    """The failure mode the launcher exists for: run directly, the script's own sys.path edit wins."""
    ref, stubs = _fake_checkout(tmp_path)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, str(stubs)]))
    r = subprocess.run([sys.executable, str(ref / "fast_infer.py"), "--prompt", "x"], capture_output=True, text=True,
                       env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode != 0 and "the reference package must not win" in r.stderr
