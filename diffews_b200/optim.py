"""Training-step tail on the B200 kernels (first pieces): MSE loss, gradient-norm clipping, AdamW.

Reference: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py
  :1384        loss = F.mse_loss(model_pred.float(), target.float(), reduction="mean")
  :1393        accelerator.clip_grad_norm_(unet.parameters(), args.max_grad_norm)
  :1186-1194   optimizer = torch.optim.AdamW(params, lr, betas=(b1, b2), weight_decay=wd, eps=eps);  :1394 optimizer.step()

`AdamW` keeps torch.optim.AdamW's constructor / `step()` / `zero_grad()` / `state_dict()`-style state names
(`exp_avg`, `exp_avg_sq`, `step`), but every parameter tensor of the model is updated by ONE launch (descriptor table +
chunk table, built once), the clip coefficient stays on the device (no host sync between the norm and the step), and
the step can emit the 16-bit tensor-core operand copy of each parameter in the same pass.  fp32 CUDA tensors only —
there is no CPU fallback.
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np
import torch

from . import ops
from ._lib import check, lib

CHUNK_ELEMS = 1 << 16           # elements per CTA: 256 threads x 64 float4 iterations


def mse_loss(pred: torch.Tensor, target: torch.Tensor, upstream: float = 1.0, want_grad: bool = True):
    """F.mse_loss(pred.float(), target.float(), reduction='mean') and its gradient w.r.t. pred (x upstream).
    Returns (loss [1] fp32 on the device, dpred or None)."""
    ops._req(pred, torch.float32, "pred"); ops._req(target, torch.float32, "target")
    assert pred.shape == target.shape
    n = pred.numel()
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    dpred = torch.empty_like(pred) if want_grad else None
    ws = torch.empty(int(lib.dfw_mse_workspace_floats()), device=pred.device, dtype=torch.float32)
    check(lib.dfw_mse_loss(pred.data_ptr(), target.data_ptr(), n, float(upstream), loss.data_ptr(), ops._ptr(dpred),
                           ws.data_ptr(), ops._stream()), "dfw_mse_loss")
    return loss, dpred


class AdamW:
    """torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2) on one fused launch.

    `params`: iterable of fp32 CUDA tensors (or nn.Parameters) whose `.grad` is an fp32 CUDA tensor at `step()` time.
    `half_copies`: optional list of 16-bit tensors (same numel as each param, bf16 or fp16, all the same dtype) that
    receive the rounded updated parameters."""

    def __init__(self, params: Iterable[torch.Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, half_copies: Optional[list] = None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise TypeError("AdamW (B200 engine) needs contiguous fp32 CUDA parameters: there is no CPU fallback")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.device = self.params[0].device
        self.state = {i: {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
                      for i, p in enumerate(self.params)}
        self.half_copies = half_copies
        self.p16_format = 0
        if half_copies is not None:
            assert len(half_copies) == len(self.params)
            dt = half_copies[0].dtype
            assert dt in (torch.bfloat16, torch.float16) and all(h.dtype == dt and h.is_cuda and h.is_contiguous()
                                                                 and h.numel() == p.numel()
                                                                 for h, p in zip(half_copies, self.params))
            self.p16_format = 1 if dt == torch.bfloat16 else 2
        # chunk table: chunk c = elements [offset, offset + CHUNK_ELEMS) of tensor t
        ct, co = [], []
        for i, p in enumerate(self.params):
            for off in range(0, p.numel(), CHUNK_ELEMS):
                ct.append(i); co.append(off)
        self.n_chunks = len(ct)
        self._chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=self.device)
        self._chunk_offset = torch.tensor(co, dtype=torch.int64, device=self.device)
        self._partial = torch.empty(self.n_chunks, dtype=torch.float32, device=self.device)
        self._norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._coef = torch.ones(1, dtype=torch.float32, device=self.device)
        self._table = None
        self._table_key = None
        self._clip_pending = False

    # -- descriptor table (rebuilt only when a .grad tensor moved) ----------------------------------------------------
    def _descs(self):
        grads = []
        for p in self.params:
            g = p.grad
            if g is None:
                raise RuntimeError("every parameter needs a .grad (sparse / partial updates are not on the DiffewS path)")
            if not g.is_cuda or g.dtype != torch.float32 or not g.is_contiguous():
                raise TypeError("gradients must be contiguous fp32 CUDA tensors")
            grads.append(g)
        key = tuple(g.data_ptr() for g in grads)
        if key != self._table_key:
            rec = np.zeros(len(self.params), dtype=np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"),
                                                             ("p16", "<u8"), ("n", "<i8")]))
            for i, (p, g) in enumerate(zip(self.params, grads)):
                st = self.state[i]
                h = self.half_copies[i].data_ptr() if self.half_copies is not None else 0
                rec[i] = (p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), h, p.numel())
            self._table = torch.from_numpy(rec.view(np.uint8).copy()).to(self.device)
            self._table_key = key
        return self._table

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_(params, max_norm) (L2): returns the total norm ([1] fp32, on the device).  The
        gradients themselves are left untouched; the clip coefficient is applied inside the next `step()` -- so unlike
        torch the .grad tensors still hold the UNCLIPPED values afterwards, and the coefficient belongs to exactly these
        gradient values: `step()` raises if a gradient was written in between (accumulate first, clip last).  A
        non-finite norm makes the next `step()` a no-op on the device (found-inf guard)."""
        t = self._descs()
        self._clip_versions = tuple((p.grad.data_ptr(), p.grad._version) for p in self.params)
        check(lib.dfw_grad_norm_clip_coef(t.data_ptr(), self._chunk_tensor.data_ptr(), self._chunk_offset.data_ptr(),
                                          self.n_chunks, CHUNK_ELEMS, float(max_norm), self._partial.data_ptr(),
                                          self._norm.data_ptr(), self._coef.data_ptr(), ops._stream()),
              "dfw_grad_norm_clip_coef")
        self._clip_pending = True
        return self._norm

    @torch.no_grad()
    def step(self, skip_nonfinite: bool = False):
        """One AdamW update.  `skip_nonfinite`: read the clipped norm back (one host sync) and, GradScaler-style, skip the
        step -- including the step counter -- when it is inf / NaN; without it the device-side guard still leaves parameters
        and moments untouched, only the bias-correction step count advances."""
        t = self._descs()
        if self._clip_pending:
            now = tuple((p.grad.data_ptr(), p.grad._version) for p in self.params)
            if now != self._clip_versions:
                self._clip_pending = False
                raise RuntimeError("a gradient changed between clip_grad_norm_() and step(): the clip coefficient is stale "
                                   "(accumulate gradients first, clip last)")
            if skip_nonfinite and not bool(torch.isfinite(self._norm).item()):
                self._clip_pending = False
                return
        for st in self.state.values():
            st["step"] += 1
        step = self.state[0]["step"]
        scale = self._coef.data_ptr() if self._clip_pending else 0
        check(lib.dfw_adamw_step(t.data_ptr(), self._chunk_tensor.data_ptr(), self._chunk_offset.data_ptr(), self.n_chunks,
                                 CHUNK_ELEMS, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, step, scale,
                                 self.p16_format, ops._stream()), "dfw_adamw_step")
        self._clip_pending = False

    # -- torch.optim.Optimizer.state_dict() layout (accelerator.save_state, train...v3.py:1408-1414) ---------------------
    def state_dict(self) -> dict:
        state = {i: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}
                 for i, st in self.state.items() if st["step"] > 0}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict):
        """Accepts what `state_dict()` or torch.optim.AdamW(params).state_dict() produced for the same parameter list."""
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("state dict does not match this optimizer's single parameter group")
        g = groups[0]
        if g.get("amsgrad") or g.get("maximize"):
            raise NotImplementedError("amsgrad / maximize are not on the DiffewS path")
        self.lr, self.betas = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1]))
        self.eps, self.weight_decay = float(g["eps"]), float(g["weight_decay"])
        for i, p in enumerate(self.params):
            st = sd["state"].get(g["params"][i])
            if st is None:
                self.state[i]["step"] = 0
                self.state[i]["exp_avg"].zero_(); self.state[i]["exp_avg_sq"].zero_()
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"parameter {i}: moment shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
            self.state[i]["step"] = int(round(float(st["step"])))
            self.state[i]["exp_avg"].copy_(st["exp_avg"])           # in place: the descriptor table holds the pointers
            self.state[i]["exp_avg_sq"].copy_(st["exp_avg_sq"])
        self._clip_pending = False

    def zero_grad(self, set_to_none: bool = False):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
