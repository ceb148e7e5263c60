"""First correct path for the GEMM / convolution backward of the training step — assembled from the
forward tcgen05 kernels, not yet a tuned implementation.

Reference: `accelerator.backward(loss)` (train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1386) runs torch
autograd through every nn.Linear / nn.Conv2d of the UNet.  Here

  Linear   y = x W^T :   dx = dy W          -> `ops.linear(dy, W^T)`            (contraction over Nout)
                         dW = dy^T x        -> `ops.linear(dy^T, x^T)` fp32     (contraction over the M tokens)
  Conv 3x3 s1 p1     :   dx = conv(dy, W rotated by 180 degrees, in/out channels swapped)   -> `ops.conv2d`
                         dW[co, tap, ci] = sum_pixels dy[p, co] * x[p + tap, ci]            -> 9 x `ops.linear`

so every FLOP runs on the implicit-GEMM kernel; the operand transposes / shifted copies are torch layout ops (extra HBM
passes — the dedicated dgrad / wgrad mainloops that read the tensors in place are the round-2 item).  16-bit operands,
fp32 accumulation, fp32 weight gradients.  CUDA only.
"""
from __future__ import annotations

import torch

from . import ops


def _pad_rows(t: torch.Tensor, mult: int) -> torch.Tensor:
    r = (-t.shape[0]) % mult
    if r == 0:
        return t
    return torch.cat([t, t.new_zeros((r,) + tuple(t.shape[1:]))], dim=0)


def linear_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, need_dx: bool = True):
    """x [M, K], w [Nout, K], dy [M, Nout] (one 16-bit dtype) -> (dx [M, K] 16-bit or None, dW [Nout, K] fp32,
    dbias [Nout] fp32).  K % 64 == 0 and Nout % 64 == 0 (true for every Linear of the UNet)."""
    assert x.is_cuda and x.dtype in ops.OPERAND_DTYPES and w.dtype == x.dtype and dy.dtype == x.dtype
    M, K = x.shape
    Nout = w.shape[0]
    assert w.shape == (Nout, K) and dy.shape == (M, Nout) and K % 64 == 0 and Nout % 64 == 0
    dx = ops.linear(dy.contiguous(), w.t().contiguous()) if need_dx else None
    xp, dyp = _pad_rows(x, 64), _pad_rows(dy, 64)                    # zero rows add nothing to the sums
    dw = ops.linear(dyp.t().contiguous(), xp.t().contiguous(), out_f32=True)
    return dx, dw, dy.float().sum(dim=0)


def conv3x3_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, need_dx: bool = True):
    """3x3 / stride 1 / pad 1 convolution.  x [N, H, W, Cin], w [Cout, 9 * Cin] (tap-major GEMM layout, see
    weights.conv_weight_to_gemm), dy [N, H, W, Cout] (one 16-bit dtype) -> (dx [N,H,W,Cin] or None, dW [Cout, 9*Cin] fp32,
    dbias [Cout] fp32).  Cin % 64 == 0, Cout % 64 == 0."""
    assert x.is_cuda and x.dtype in ops.OPERAND_DTYPES and w.dtype == x.dtype and dy.dtype == x.dtype
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    assert w.shape == (Cout, 9 * Cin) and dy.shape == (N, H, W, Cout) and Cin % 64 == 0 and Cout % 64 == 0
    dx = None
    if need_dx:
        w_rot = w.view(Cout, 3, 3, Cin).flip(1, 2).permute(3, 1, 2, 0).contiguous().view(Cin, 9 * Cout)
        dx = ops.conv2d(dy.contiguous(), w_rot, ksize=3)
    P = N * H * W
    dyt = _pad_rows(dy.reshape(P, Cout), 64).t().contiguous()       # [Cout, P]
    xpad = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))            # [N, H+2, W+2, Cin]
    taps = []
    for kh in range(3):
        for kw in range(3):
            xs = _pad_rows(xpad[:, kh:kh + H, kw:kw + W, :].reshape(P, Cin), 64).t().contiguous()     # [Cin, P]
            taps.append(ops.linear(dyt, xs, out_f32=True))                                           # [Cout, Cin]
    dw = torch.stack(taps, dim=1).reshape(Cout, 9 * Cin)
    return dx, dw, dy.float().sum(dim=(0, 1, 2))
