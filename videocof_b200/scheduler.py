"""Flow-matching UniPC multistep sampler — host-side scheduler loop of the denoising path.

Same class name, constructor arguments, `set_timesteps(..., shift=)` / `step()` contract and
numerics as the reference's default sampler (videox_fun/utils/fm_solvers_unipc.py:
set_timesteps :160-227, convert_model_output :279-348, predictor :350-482, corrector :484-626,
step :655-737).  It stays in PyTorch on purpose (north star: "host code stays Python/PyTorch
for tensor plumbing and the scheduler loop"): per step it is a handful of latent-sized
elementwise ops.  Scalar coefficients are computed in fp32 on the CPU exactly like the
reference so that bf16 latents round identically.  Pinned by tests/test_scheduler.py against
golden trajectories of the executed reference.
"""
import numpy as np
import torch


class _Cfg(dict):
    __getattr__ = dict.get


class SchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample

    def __getitem__(self, i):
        return (self.prev_sample,)[i]


class FlowUniPCMultistepScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, solver_order=2, prediction_type="flow_prediction", shift=1.0,
                 use_dynamic_shifting=False, thresholding=False, dynamic_thresholding_ratio=0.995,
                 sample_max_value=1.0, predict_x0=True, solver_type="bh2", lower_order_final=True,
                 disable_corrector=(), solver_p=None, timestep_spacing="linspace", steps_offset=0,
                 final_sigmas_type="zero"):
        if solver_type in ("midpoint", "heun", "logrho"):
            solver_type = "bh2"
        if solver_type not in ("bh1", "bh2"):
            raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if prediction_type != "flow_prediction":
            raise ValueError("only flow_prediction is supported (the Wan / VideoCoF setting)")
        if thresholding or use_dynamic_shifting or solver_p is not None:
            raise NotImplementedError("thresholding / dynamic shifting / solver_p are not on the VideoCoF path")
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                           prediction_type=prediction_type, shift=shift, use_dynamic_shifting=False,
                           thresholding=False, predict_x0=predict_x0, solver_type=solver_type,
                           lower_order_final=lower_order_final, final_sigmas_type=final_sigmas_type,
                           timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.predict_x0 = predict_x0
        self.disable_corrector = list(disable_corrector)
        self.init_noise_sigma = 1.0
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sig = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()
        self.num_inference_steps = None
        self._reset()

    def _reset(self):
        self.model_outputs = [None] * self.config.solver_order
        self.timestep_list = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = None
        self._begin_index = None

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, shift=None):
        """sigma grid linspace(sigma_max, sigma_min, N+1)[:-1], then the call-time shift (:183-193)."""
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.config.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        if self.config.final_sigmas_type != "zero":
            raise ValueError("final_sigmas_type must be 'zero' on the VideoCoF path")
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)                        # kept on the CPU (:224-226)
        self.timesteps = torch.from_numpy(timesteps).to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(timesteps)
        self._reset()

    # ---- UniPC B(h) coefficients, fp32 scalars on the CPU -------------------------------------------
    def _lam(self, sigma):
        return torch.log(1 - sigma) - torch.log(sigma)

    def _bh_terms(self, idx_t, idx_s0, hist_idx, order):
        """Shared by predictor and corrector: returns (sigma_t, sigma_s0, alpha_t, h_phi_1, B_h, rks, R, b)."""
        sigma_t, sigma_s0 = self.sigmas[idx_t], self.sigmas[idx_s0]
        alpha_t = 1 - sigma_t
        lam_t, lam_s0 = self._lam(sigma_t), self._lam(sigma_s0)
        h = lam_t - lam_s0
        rks = []
        for si in hist_idx[:order - 1]:
            rks.append((self._lam(self.sigmas[si]) - lam_s0) / h)
        rk_list = list(rks)
        rks = torch.tensor(rks + [1.0])
        hh = -h if self.predict_x0 else h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.config.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return sigma_t, sigma_s0, alpha_t, h_phi_1, B_h, rk_list, torch.stack(R), torch.tensor(b)

    def convert_model_output(self, model_output, sample):
        """x0 = x - sigma * v (:319-321)."""
        if not self.predict_x0:
            raise NotImplementedError("predict_x0=False is not used by the reference pipelines")
        return sample - self.sigmas[self.step_index] * model_output

    def _predict(self, sample, order):
        """UniP (:350-482)."""
        m0 = self.model_outputs[-1]
        k = self.step_index
        hist = [k - i for i in range(1, order)]
        sigma_t, sigma_s0, alpha_t, h_phi_1, B_h, rks, R, b = self._bh_terms(k + 1, k, hist, order)
        x_t = sigma_t / sigma_s0 * sample - alpha_t * h_phi_1 * m0
        if order > 1:
            D1s = torch.stack([(self.model_outputs[-(i + 1)] - m0) / rks[i - 1] for i in range(1, order)], dim=1)
            if order == 2:
                rhos = torch.tensor([0.5], dtype=sample.dtype, device=sample.device)
            else:
                rhos = torch.linalg.solve(R[:-1, :-1], b[:-1]).to(sample.device).to(sample.dtype)
            x_t = x_t - alpha_t * B_h * torch.einsum("k,bkc...->bc...", rhos, D1s)
        else:
            x_t = x_t - alpha_t * B_h * 0
        return x_t.to(sample.dtype)

    def _correct(self, model_t, last_sample, order):
        """UniC (:484-626)."""
        m0 = self.model_outputs[-1]
        k = self.step_index
        hist = [k - (i + 1) for i in range(1, order)]
        sigma_t, sigma_s0, alpha_t, h_phi_1, B_h, rks, R, b = self._bh_terms(k, k - 1, hist, order)
        x = last_sample
        if order == 1:
            rhos = torch.tensor([0.5], dtype=x.dtype, device=x.device)
        else:
            rhos = torch.linalg.solve(R, b).to(x.device).to(x.dtype)
        x_t = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr = 0
        if order > 1:
            D1s = torch.stack([(self.model_outputs[-(i + 1)] - m0) / rks[i - 1] for i in range(1, order)], dim=1)
            corr = torch.einsum("k,bkc...->bc...", rhos[:-1], D1s)
        x_t = x_t - alpha_t * B_h * (corr + rhos[-1] * (model_t - m0))
        return x_t.to(x.dtype)

    def index_for_timestep(self, timestep):
        idx = (self.timesteps == timestep).nonzero()
        return idx[1 if len(idx) > 1 else 0].item()

    def step(self, model_output, timestep, sample, return_dict=True, generator=None):
        """One multistep UniPC update (:655-737): corrector on the stored history, then predictor."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after "
                             "creating the scheduler")
        if self._step_index is None:
            if self._begin_index is None:
                if isinstance(timestep, torch.Tensor):
                    timestep = timestep.to(self.timesteps.device)
                self._step_index = self.index_for_timestep(timestep)
            else:
                self._step_index = self._begin_index
        use_corrector = (self.step_index > 0 and self.step_index - 1 not in self.disable_corrector
                         and self.last_sample is not None)
        x0 = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self._correct(x0, self.last_sample, self.this_order)
        n = self.config.solver_order
        for i in range(n - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
            self.timestep_list[i] = self.timestep_list[i + 1]
        self.model_outputs[-1] = x0
        self.timestep_list[-1] = timestep
        this_order = min(n, len(self.timesteps) - self.step_index) if self.config.lower_order_final else n
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower_order_nums < n:
            self.lower_order_nums += 1
        self._step_index += 1
        return SchedulerOutput(prev) if return_dict else (prev,)
