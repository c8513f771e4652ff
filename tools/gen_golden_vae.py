"""Golden generators for the scheduler and the VAE (split from gen_golden.py for size)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

import ref_loader  # noqa: E402

UNIPC_CASES = {"unipc_4step_shift3": (4, 3.0), "unipc_9step_shift5": (9, 5.0)}


def fake_model(x, t):
    """Deterministic stand-in for the DiT inside scheduler goldens."""
    return torch.tanh(x.float() * 0.7 + float(t) / 1000.0 - 0.3).to(x.dtype)


def unipc_trajectory(sched, steps, shift, dtype):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, 3, 6, 5, generator=g).to(dtype)
    sched.set_timesteps(steps, device="cpu", shift=shift)
    traj = [x.float().numpy().copy()]
    for t in sched.timesteps:
        v = fake_model(x, t)
        x = sched.step(v, t, x, return_dict=False)[0]
        traj.append(x.float().numpy().copy())
    return np.stack(traj), sched.timesteps.numpy().copy(), sched.sigmas.numpy().copy()


def gen_unipc():
    ns = ref_loader.load_reference()
    for name, (steps, shift) in UNIPC_CASES.items():
        out = {}
        for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            # the CLIs force shift=1 in the constructor and pass the real shift to set_timesteps
            # (fast_infer.py:333-334, pipeline_wan.py:613-615)
            sched = ns.unipc.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
            traj, ts, sig = unipc_trajectory(sched, steps, shift, dt)
            out["traj_" + tag] = traj
            out["timesteps"], out["sigmas"] = ts, sig
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print("wrote", name, out["timesteps"])


def gen_vae():
    from gen_golden_vae_impl import gen_vae_impl
    gen_vae_impl()
