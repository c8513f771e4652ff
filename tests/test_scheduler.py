"""videocof_b200.scheduler (host-side UniPC loop) against golden trajectories of the executed reference
scheduler (videox_fun/utils/fm_solvers_unipc.py), fp32 and bf16 latents."""
import os

import numpy as np
import pytest
import torch

from gen_golden_vae import UNIPC_CASES, unipc_trajectory
from videocof_b200.scheduler import FlowUniPCMultistepScheduler


@pytest.mark.parametrize("name", list(UNIPC_CASES))
def test_unipc_matches_reference_trajectory(name, golden_dir):
    steps, shift = UNIPC_CASES[name]
    gold = np.load(os.path.join(golden_dir, name + ".npz"))
    for dt, tag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
        traj, ts, sig = unipc_trajectory(sched, steps, shift, dt)
        assert np.array_equal(ts, gold["timesteps"])
        assert np.array_equal(sig, gold["sigmas"])
        # identical op order on identical fp32 CPU scalars -> bit-exact trajectories
        assert np.array_equal(traj, gold["traj_" + tag]), (tag, np.abs(traj - gold["traj_" + tag]).max())


def test_fast_infer_schedule_values():
    """4 steps, shift 3 -> timesteps [999, 899, 749, 499]."""
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, solver_order=2)
    s.set_timesteps(4, device="cpu", shift=3)
    assert s.timesteps.tolist() == [999, 899, 749, 499]
    assert float(s.sigmas[-1]) == 0.0
