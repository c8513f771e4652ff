"""The parts of WanPipeline the CLIs touch around `__call__` (fast_infer.py:336-361, pipeline_wan.py:449-498): device /
offload surface and the argument checks, with their reference conditions.  Pure host logic."""
import pytest
import torch

from videocof_b200.pipeline import WanPipeline


class _M(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(2))
        self.moves = []

    def to(self, *a, **k):
        self.moves.append(a)
        return self

    @property
    def device(self):
        return self.w.device

    @property
    def dtype(self):
        return self.w.dtype


def _pipe():
    return WanPipeline(None, _M(), _M(), _M(), object())


def test_offload_requests_keep_everything_resident(capsys):
    p = _pipe()
    assert p.enable_sequential_cpu_offload(device="cpu") is p          # the CLIs' default memory mode (fast_infer.py:136)
    assert p.enable_model_cpu_offload(device="cpu") is p
    assert "offload request ignored" in capsys.readouterr().out
    for m in (p.text_encoder, p.vae, p.transformer):
        assert m.moves == [("cpu",), ("cpu",)]
    assert p._execution_device == torch.device("cpu")
    p.to("cpu", torch.bfloat16)
    assert p.vae.moves[-1] == ("cpu", torch.bfloat16)
    assert p.maybe_free_model_hooks() is None and p.interrupt is False and p.num_timesteps == 0


def test_execution_device_defaults_to_the_transformer():
    assert _pipe()._execution_device == torch.device("cpu")


@pytest.mark.parametrize("kw,needle", [
    (dict(prompt="a", height=30, width=48), "divisible by 8"),
    (dict(prompt="a", height=32, width=48, cb=["latents", "nope"]), "callback_on_step_end_tensor_inputs"),
    (dict(prompt="a", height=32, width=48, prompt_embeds=[torch.zeros(1, 4)]), "both `prompt` and `prompt_embeds`"),
    (dict(prompt=None, height=32, width=48), "Provide either"),
    (dict(prompt=3, height=32, width=48), "has to be of type"),
    (dict(prompt="a", height=32, width=48, negative_prompt_embeds=[torch.zeros(1, 4)]), "negative_prompt_embeds"),
    (dict(prompt=None, height=32, width=48, negative_prompt="b", prompt_embeds=[torch.zeros(1, 4)],
          negative_prompt_embeds=[torch.zeros(1, 4)]), "both `negative_prompt` and `negative_prompt_embeds`"),
    (dict(prompt=None, height=32, width=48, prompt_embeds=torch.zeros(1, 5, 4), negative_prompt_embeds=torch.zeros(1, 6, 4)),
     "same shape"),
])
def test_check_inputs_conditions_of_the_reference(kw, needle):
    p = _pipe()
    with pytest.raises(ValueError, match=needle):
        p.check_inputs(kw.get("prompt"), kw["height"], kw["width"], kw.get("negative_prompt"), kw.get("cb", ["latents"]),
                       kw.get("prompt_embeds"), kw.get("negative_prompt_embeds"))


def test_check_inputs_accepts_the_pipeline_forms():
    p = _pipe()
    p.check_inputs("a prompt", 32, 48, "a negative prompt", ["latents"])
    p.check_inputs(["a", "b"], 32, 48, None, None)
    # lists of per-sample embeddings with different token counts (the form encode_prompt returns)
    p.check_inputs(None, 32, 48, None, ("latents",), [torch.zeros(7, 4)], [torch.zeros(3, 4)])


def test_encode_prompt_argument_errors():
    p = _pipe()
    with pytest.raises(TypeError, match="same type"):
        p.encode_prompt(["a"], negative_prompt=3, prompt_embeds=[torch.zeros(1, 4)])
    with pytest.raises(ValueError, match="batch size"):
        p.encode_prompt(["a", "b"], negative_prompt=["x"], prompt_embeds=[torch.zeros(1, 4)] * 2)
    pe, ne = p.encode_prompt(None, do_classifier_free_guidance=False, prompt_embeds=[torch.zeros(1, 4)])
    assert ne is None and len(pe) == 1
