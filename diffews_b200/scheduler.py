"""DDIMSchedulerCustomized — restatement of marigold/util/scheduler_customized.py:107-180 (reference) and of the
diffusers-0.25 DDIMScheduler.set_timesteps / step it inherits (upstream), limited to what the hot path exercises:
`scaled_linear` betas, `leading` spacing, `steps_offset`, v-/epsilon-/sample- prediction, eta = 0, no clipping.

With the reference's scheduler_1.0_1.0/scheduler_config.json (beta_start = beta_end = 1.0, v_prediction,
set_alpha_to_one = false) every alpha_cumprod is exactly 0, so `step(...).pred_original_sample == -model_output`
bit-for-bit; the pipeline detects this and fuses the negation into the VAE-decode prologue.
"""
from __future__ import annotations

import json
from types import SimpleNamespace

import torch

DEFAULT_CONFIG = {   # scheduler_1.0_1.0/scheduler_config.json:1-19
    "beta_start": 1.0, "beta_end": 1.0, "beta_schedule": "scaled_linear", "num_train_timesteps": 1000,
    "prediction_type": "v_prediction", "set_alpha_to_one": False, "steps_offset": 1, "timestep_spacing": "leading",
    "clip_sample": False, "thresholding": False,
}


class DDIMSchedulerCustomized:
    def __init__(self, **config):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update({k: v for k, v in config.items() if not k.startswith("_")})
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.beta_schedule == "scaled_linear":
            betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, c.num_train_timesteps, dtype=torch.float32) ** 2
        elif c.beta_schedule == "linear":
            betas = torch.linspace(c.beta_start, c.beta_end, c.num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(c.beta_schedule)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if c.set_alpha_to_one else self.alphas_cumprod[0]
        self.timesteps = torch.arange(c.num_train_timesteps - 1, -1, -1)
        self.num_inference_steps = None

    @classmethod
    def from_config_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing != "leading":
            raise NotImplementedError(c.timestep_spacing)
        ratio = c.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).to(torch.int64) + c.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def is_pure_negation(self, timestep) -> bool:
        """True when step() reduces to pred_original_sample = -model_output exactly (alpha_cumprod == 0, v-pred)."""
        return self.config.prediction_type == "v_prediction" and float(self.alphas_cumprod[int(timestep)]) == 0.0

    def step(self, model_output, timestep, sample, eta: float = 0.0):
        c = self.config
        t = int(timestep)
        prev_t = t - c.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].to(model_output.device)
        a_prev = (self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod).to(model_output.device)
        b_t = 1 - a_t
        if c.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif c.prediction_type == "sample":
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif c.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(c.prediction_type)
        if eta != 0.0:
            raise NotImplementedError("eta > 0")
        prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)
