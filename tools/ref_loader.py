"""Load the UNMODIFIED reference hot-path files from /root/reference for golden generation.

The reference package cannot be imported as a whole here (diffusers, accelerate, omegaconf …
are not installed and there is no network).  Its hot-path files import only a
handful of diffusers names, so this module installs tiny in-memory stand-ins for those names
and loads the files with importlib straight from the read-only mount.  Nothing is copied.

Used by tools/gen_golden*.py in the build container and — against the staged, git-ignored copy
baseline/_ref (tools/stage_reference.py) — by tools/gpu_reference.py on the GPU box, where it runs the
reference's own CUDA path as the GPU performance / parity reference.  The product never imports it.
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """$VCOF_REFERENCE, else the read-only mount of the build container, else the staged copy that travels to the
    GPU box (baseline/_ref, git-ignored; tools/stage_reference.py)."""
    env = os.environ.get("VCOF_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_ROOT, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "videox_fun")):
            return cand
    return "/root/reference"


REF = _find_reference()


class _Config(dict):
    __getattr__ = dict.get


def _register_to_config(init):
    import functools
    import inspect

    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        object.__setattr__(self, "_vcof_cfg", _Config(cfg))
        init(self, *args, **kwargs)
    return wrapper


class _ConfigMixin:
    @property
    def config(self):
        return self._vcof_cfg


class _ModelMixin(nn.Module):
    def __getattr__(self, name):
        """diffusers' ModelMixin resolves unknown attributes through the registered config
        (e.g. vae.spatial_compression_ratio, pipeline_wan.py:138)."""
        try:
            return super().__getattr__(name)
        except AttributeError:
            cfg = self.__dict__.get("_vcof_cfg")
            if cfg is not None and name in cfg:
                return cfg[name]
            raise

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _Dist:
    def __init__(self, h):
        self.h = h

    def mode(self):
        return self.h.chunk(2, dim=1)[0]


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, i):
        return list(self.__dict__.values())[i]


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_vcof_shim", False):
        return

    class _Log:
        def get_logger(self, *_a, **_k):
            import logging as _l
            return _l.getLogger("ref")

    _mod("diffusers", _vcof_shim=True)
    _mod("diffusers.configuration_utils", ConfigMixin=_ConfigMixin, register_to_config=_register_to_config)
    _mod("diffusers.loaders")
    _mod("diffusers.loaders.single_file_model", FromOriginalModelMixin=type("FromOriginalModelMixin", (), {}))
    _mod("diffusers.models")
    _mod("diffusers.models.modeling_utils", ModelMixin=_ModelMixin)
    _mod("diffusers.models.autoencoders")
    _mod("diffusers.models.autoencoders.vae", DecoderOutput=lambda sample: _Out(sample=sample),
         DiagonalGaussianDistribution=_Dist)
    _mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=lambda latent_dist: _Out(latent_dist=latent_dist))
    _mod("diffusers.utils", is_torch_version=lambda *_a: True, logging=_Log(), deprecate=lambda *a, **k: None,
         is_scipy_available=lambda: True)
    _mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    _mod("diffusers.schedulers")
    _mod("diffusers.schedulers.scheduling_utils", KarrasDiffusionSchedulers=[],
         SchedulerMixin=type("SchedulerMixin", (), {}),
         SchedulerOutput=lambda prev_sample: _Out(prev_sample=prev_sample))


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


_cache = {}


def load_reference(pkg="videox_fun"):
    """Returns a namespace with the reference's wan_transformer3d, wan_vae, fm_solvers_unipc modules.

    `pkg` is the top-level name the files are registered under in sys.modules (their imports of each other are all
    relative, so any name works): the golden generators keep "videox_fun"; tools/gpu_reference.py uses a private name
    so that the repo's own `videox_fun` overlay stays importable in the same process."""
    if pkg in _cache:
        return _cache[pkg]
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} not present: goldens can only be regenerated in the build container")
    if not torch.cuda.is_available():
        os.environ.setdefault("VIDEOX_ATTENTION_TYPE", "SDPA")  # flash-attn asserts CUDA (attention_utils.py:73)
    install_shims()
    pk = _mod(pkg)
    pk.__path__ = []
    for sub in ("models", "utils", "dist"):
        m = _mod(f"{pkg}.{sub}")
        m.__path__ = []
    d = sys.modules[f"{pkg}.dist"]
    for n in ("get_sequence_parallel_rank", "get_sequence_parallel_world_size", "get_sp_group",
              "usp_attn_forward", "xFuserLongContextAttention"):
        setattr(d, n, None)
    cfgopt = _load(f"{pkg}.utils.cfg_optimization", "videox_fun/utils/cfg_optimization.py")
    sys.modules[f"{pkg}.utils"].cfg_skip = cfgopt.cfg_skip
    attn = _load(f"{pkg}.models.attention_utils", "videox_fun/models/attention_utils.py")
    _load(f"{pkg}.models.cache_utils", "videox_fun/models/cache_utils.py")
    _mod(f"{pkg}.models.wan_camera_adapter", SimpleAdapter=None)
    dit = _load(f"{pkg}.models.wan_transformer3d", "videox_fun/models/wan_transformer3d.py")
    vae = _load(f"{pkg}.models.wan_vae", "videox_fun/models/wan_vae.py")
    unipc = _load(f"{pkg}.utils.fm_solvers_unipc", "videox_fun/utils/fm_solvers_unipc.py")
    ns = types.SimpleNamespace(dit=dit, vae=vae, unipc=unipc, attention_utils=attn, root=REF)
    _cache[pkg] = ns
    if pkg == "videox_fun":
        _cache["ns"] = ns
    return ns


# ------------------------------------------------------------------------------------------------------------
# The caller of the hot path: videox_fun/pipeline/pipeline_wan.py, executed unmodified.  It needs a few more
# diffusers names (DiffusionPipeline, randn_tensor, VideoProcessor …).  The stand-ins below restate only the
# behaviour WanPipeline.__call__ touches; randn_tensor follows diffusers' published semantics (draw on the
# generator's device, then move).  The file brackets the DiT call with `torch.cuda.device(device)`, which
# rejects a CPU device, so the loader swaps that context manager for a null one while the pipeline runs on CPU.
# ------------------------------------------------------------------------------------------------------------
class _DiffusionPipeline:
    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def _execution_device(self):
        return torch.device("cpu")

    def progress_bar(self, total=None):
        import contextlib

        class _Bar:
            def update(self, *_a):
                pass
        return contextlib.nullcontext(_Bar())

    def maybe_free_model_hooks(self):
        pass


def _randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    gen = generator[0] if isinstance(generator, (list, tuple)) else generator
    device = torch.device(device or "cpu")
    rand_device = device
    if gen is not None and gen.device.type != device.type and gen.device.type == "cpu":
        rand_device = torch.device("cpu")
    return torch.randn(shape, generator=gen, device=rand_device, dtype=dtype).to(device)


def load_reference_pipeline():
    """-> namespace with .pipeline (pipeline_wan module), .t5 (wan_text_encoder module) and the models of
    load_reference()."""
    if "pipe" in _cache:
        return _cache["pipe"]
    import contextlib
    ns = load_reference()
    d = sys.modules["diffusers"]
    d.FlowMatchEulerDiscreteScheduler = type("FlowMatchEulerDiscreteScheduler", (), {})
    _mod("diffusers.callbacks", MultiPipelineCallbacks=type("MultiPipelineCallbacks", (), {}),
         PipelineCallback=type("PipelineCallback", (), {}))
    _mod("diffusers.pipelines")
    _mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=_DiffusionPipeline)
    u = sys.modules["diffusers.utils"]
    u.BaseOutput = type("BaseOutput", (), {})
    u.replace_example_docstring = lambda _doc: (lambda f: f)
    _mod("diffusers.utils.torch_utils", randn_tensor=_randn_tensor)
    _mod("diffusers.video_processor", VideoProcessor=lambda **_k: None)
    t5 = _load("videox_fun.models.wan_text_encoder", "videox_fun/models/wan_text_encoder.py")
    m = sys.modules["videox_fun.models"]
    m.AutoencoderKLWan, m.WanTransformer3DModel = ns.vae.AutoencoderKLWan, ns.dit.WanTransformer3DModel
    m.WanT5EncoderModel, m.AutoTokenizer, m.CLIPModel = t5.WanT5EncoderModel, None, None
    _mod("videox_fun.utils.fm_solvers", FlowDPMSolverMultistepScheduler=type("FlowDPMSolverMultistepScheduler", (), {}),
         get_sampling_sigmas=None)
    sys.modules["videox_fun.utils.fm_solvers_unipc"] = ns.unipc
    p = _mod("videox_fun.pipeline")
    p.__path__ = []
    pipe = _load("videox_fun.pipeline.pipeline_wan", "videox_fun/pipeline/pipeline_wan.py")
    if not torch.cuda.is_available():
        torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    out = types.SimpleNamespace(pipeline=pipe, t5=t5, dit=ns.dit, vae=ns.vae, unipc=ns.unipc)
    _cache["pipe"] = out
    return out
