"""Host-side mirrors of the two ``diffusers`` schedulers the reference instantiates
(``configs/modules/scheduler.yaml:1-14,31-39``; consumed at ``models/modeltype/ladiff.py:113-115,407-417,491-492,776``).

``diffusers`` is a third-party, unpinned dependency of the reference (``src/requirements.txt:23``) that is not vendored
and not installable here, so these classes restate the published algorithms (DDIM: Song et al. 2021 eq. 12; DDPM:
Ho et al. 2020) with the constructor arguments the YAML passes.  *Parity unpinned* for the scheduler: see DESIGN.md.

The fused CUDA loop does not call ``step``; it consumes ``fused_coefficients()`` (the eta=0 closed form
``x' = c1 x + c2 eps``).  ``step`` / ``add_noise`` keep the object usable wherever the reference uses it.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np
import torch


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas=None):
    if trained_betas is not None:
        return torch.as_tensor(trained_betas, dtype=torch.float32)
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    if beta_schedule == "squaredcos_cap_v2":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
        return torch.tensor([min(1 - f((i + 1) / num_train_timesteps) / f(i / num_train_timesteps), 0.999)
                             for i in range(num_train_timesteps)], dtype=torch.float32)
    raise NotImplementedError(f"{beta_schedule} does is not implemented for this scheduler")


class SchedulerOutput(SimpleNamespace):
    def __getitem__(self, i):
        return (self.prev_sample, self.pred_original_sample)[i]


class _Base:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, prediction_type="epsilon", **extra):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                      beta_schedule=beta_schedule, trained_betas=trained_betas, clip_sample=clip_sample,
                                      prediction_type=prediction_type, **extra)
        self.betas = _betas(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        """q(x_t | x_0) -- used by training only (ladiff.py:776)."""
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        a = acp[timesteps] ** 0.5
        s = (1 - acp[timesteps]) ** 0.5
        while a.dim() < original_samples.dim():
            a, s = a.unsqueeze(-1), s.unsqueeze(-1)
        return a * original_samples + s * noise

    def __len__(self):
        return self.config.num_train_timesteps


class DDIMScheduler(_Base):
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, clip_sample=True, set_alpha_to_one=True, steps_offset=0,
                 prediction_type="epsilon", **kwargs):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas, clip_sample,
                         prediction_type, set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, **kwargs)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(f"`num_inference_steps`: {num_inference_steps} cannot be larger than "
                             f"`num_train_timesteps`: {self.config.num_train_timesteps}")
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def _alpha_pair(self, t: int) -> Tuple[float, float]:
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].double().item()
        a_p = self.alphas_cumprod[prev_t].double().item() if prev_t >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_p

    def fused_coefficients(self, eta: float = 0.0) -> Tuple[List[int], List[float], List[float]]:
        """(timesteps, c1, c2) with  prev = c1 * sample + c2 * eps  for every inference step (eta = 0, epsilon
        prediction, no clipping -- the only configuration the reference's sampling YAML uses)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0 or self.config.clip_sample or self.config.prediction_type != "epsilon":
            raise ValueError("fused DDIM loop supports eta=0, clip_sample=False, prediction_type='epsilon' "
                             "(configs/modules/scheduler.yaml:3-13); use .step() for other settings")
        ts, c1, c2 = [], [], []
        for t in self.timesteps.tolist():
            a_t, a_p = self._alpha_pair(int(t))
            ts.append(int(t))
            c1.append(math.sqrt(a_p / a_t))
            c2.append(math.sqrt(1 - a_p) - math.sqrt(a_p) * math.sqrt(1 - a_t) / math.sqrt(a_t))
        return ts, c1, c2

    def step(self, model_output, timestep, sample, eta: float = 0.0, use_clipped_model_output: bool = False,
             generator=None, variance_noise=None, return_dict: bool = True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        t = int(timestep)
        a_t, a_p = self._alpha_pair(t)
        b_t = 1 - a_t
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.config.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        else:
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, `sample`")
        if self.config.clip_sample:
            x0 = x0.clamp(-1, 1)
        var = (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)
        std = eta * var ** 0.5
        if use_clipped_model_output:
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        prev = a_p ** 0.5 * x0 + (1 - a_p - std ** 2) ** 0.5 * eps
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                             dtype=model_output.dtype)
            prev = prev + std * variance_noise
        out = SchedulerOutput(prev_sample=prev, pred_original_sample=x0)
        return out if return_dict else (prev,)


class DDPMScheduler(_Base):
    """Training ``noise_scheduler`` (scheduler.yaml:31-39) and the commented-out 1000-step sampler (:16-29)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, variance_type="fixed_small", clip_sample=True, prediction_type="epsilon", **kwargs):
        super().__init__(num_train_timesteps, beta_start, beta_end, beta_schedule, trained_betas, clip_sample,
                         prediction_type, variance_type=variance_type, **kwargs)
        self.one = torch.tensor(1.0)

    def set_timesteps(self, num_inference_steps: int, device=None):
        num_inference_steps = min(self.config.num_train_timesteps, num_inference_steps)
        self.num_inference_steps = num_inference_steps
        ts = np.arange(0, self.config.num_train_timesteps, self.config.num_train_timesteps // num_inference_steps)[::-1]
        ts = ts.copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)

    def _alphas(self, t: int):
        """(alpha_bar_t, alpha_bar_prev, alpha_t, beta_t) for the step t -> t - train/inference (== t - 1 at 1000 steps, the
        only DDPM sampling configuration the reference's YAML lists, scheduler.yaml:16-29)."""
        ratio = self.config.num_train_timesteps // (self.num_inference_steps or self.config.num_train_timesteps)
        prev_t = t - ratio
        a_t = self.alphas_cumprod[t].double().item()
        a_p = self.alphas_cumprod[prev_t].double().item() if prev_t >= 0 else 1.0
        cur_alpha = a_t / a_p
        return a_t, a_p, cur_alpha, 1.0 - cur_alpha

    def _get_variance(self, t: int) -> float:
        a_t, a_p, _, cur_beta = self._alphas(t)
        var = (1 - a_p) / (1 - a_t) * cur_beta
        vt = self.config.variance_type
        if vt == "fixed_small":
            return max(var, 1e-20)
        if vt == "fixed_small_log":
            return math.exp(0.5 * math.log(max(var, 1e-20))) ** 2
        if vt == "fixed_large":
            return cur_beta
        if vt == "fixed_large_log":
            return math.exp(0.5 * math.log(cur_beta)) ** 2
        raise ValueError(f"variance_type {vt} not supported")

    def fused_coefficients(self, eta: float = 0.0) -> Tuple[List[int], List[float], List[float], List[float]]:
        """(timesteps, c1, c2, c3) with  prev = c1 * sample + c2 * eps + c3 * noise  (epsilon prediction, no clipping):
        x0 = (x - sqrt(1 - abar_t) eps) / sqrt(abar_t);  prev = k0 x0 + kt x + sqrt(var_t) noise  (Ho et al. 2020 eq. 7)."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.config.clip_sample or self.config.prediction_type != "epsilon":
            raise ValueError("fused DDPM loop supports clip_sample=False, prediction_type='epsilon' "
                             "(configs/modules/scheduler.yaml:16-29); use .step() for other settings")
        ts, c1, c2, c3 = [], [], [], []
        for t in self.timesteps.tolist():
            t = int(t)
            a_t, a_p, cur_alpha, cur_beta = self._alphas(t)
            k0 = math.sqrt(a_p) * cur_beta / (1 - a_t)
            kt = math.sqrt(cur_alpha) * (1 - a_p) / (1 - a_t)
            ts.append(t)
            c1.append(kt + k0 / math.sqrt(a_t))
            c2.append(-k0 * math.sqrt(1 - a_t) / math.sqrt(a_t))
            c3.append(math.sqrt(self._get_variance(t)) if t > 0 else 0.0)
        return ts, c1, c2, c3

    def step(self, model_output, timestep, sample, generator=None, return_dict: bool = True, variance_noise=None, **kwargs):
        t = int(timestep)
        a_t, a_p, cur_alpha, cur_beta = self._alphas(t)
        b_t, b_p = 1 - a_t, 1 - a_p
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.config.prediction_type == "sample":
            x0 = model_output
        else:
            raise ValueError(f"prediction_type given as {self.config.prediction_type} must be one of `epsilon`, `sample`")
        if self.config.clip_sample:
            x0 = x0.clamp(-1, 1)
        c0 = (a_p ** 0.5 * cur_beta) / b_t
        ct = cur_alpha ** 0.5 * b_p / b_t
        prev = c0 * x0 + ct * sample
        if t > 0:
            noise = variance_noise if variance_noise is not None else torch.randn(
                model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
            prev = prev + self._get_variance(t) ** 0.5 * noise
        out = SchedulerOutput(prev_sample=prev, pred_original_sample=x0)
        return out if return_dict else (prev,)
