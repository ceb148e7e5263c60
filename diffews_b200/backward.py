"""Stand-alone GEMM / convolution backward on the gradient kernels.

Reference: `accelerator.backward(loss)` (train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1386) runs torch
autograd through every nn.Linear / nn.Conv2d of the UNet.  Here

  Linear   y = x W^T :   dx = dy W          -> the forward GEMM kernel on W^T (dfw_weight_permute: one transpose kernel)
                         dW = dy^T x        -> dfw_conv_wgrad (tcgen05, dy and x read in place as MN-major TMA tiles)
  Conv 3x3 s1 p1     :   dx = conv(dy, W rotated by 180 degrees, in/out channels swapped)   (same permute kernel)
                         dW[co, tap, ci] = sum_pixels dy[p, co] * x[p + tap, ci]            (same wgrad kernel, 9 taps)
                         dbias = column sums of dy                                          (dfw_colsum)

No torch layout ops, no operand copies: the round-1 composition (pad / flip / permute / 9 GEMM launches per conv wgrad)
is gone.  The training step (diffews_b200/train.py) drives the same kernels through autograd nodes and keeps the permuted
weights cached between optimizer steps; these two functions are the per-layer form used by the kernel tests.
16-bit operands, fp32 accumulation, fp32 weight / bias gradients.  CUDA only.
"""
from __future__ import annotations

import torch

from . import ops


def linear_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, need_dx: bool = True):
    """x [M, K], w [Nout, K], dy [M, Nout] (one 16-bit dtype) -> (dx [M, K] 16-bit or None, dW [Nout, K] fp32,
    dbias [Nout] fp32).  K % 64 == 0 and Nout % 64 == 0 (true for every Linear of the UNet); any M."""
    assert x.is_cuda and x.dtype in ops.OPERAND_DTYPES and w.dtype == x.dtype and dy.dtype == x.dtype
    M, K = x.shape
    Nout = w.shape[0]
    assert w.shape == (Nout, K) and dy.shape == (M, Nout) and K % 64 == 0 and Nout % 64 == 0
    x, w, dy = x.contiguous(), w.contiguous(), dy.contiguous()
    dx = ops.linear(dy, ops.weight_permute(w, Nout, 1, K, [0]).view(K, Nout)) if need_dx else None
    dw = ops.linear_wgrad(x, dy, torch.empty((Nout, K), device=x.device, dtype=torch.float32))
    return dx, dw, ops.colsum(dy).view(Nout)


def conv3x3_backward(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, need_dx: bool = True):
    """3x3 / stride 1 / pad 1 convolution.  x [N, H, W, Cin], w [Cout, 9 * Cin] (tap-major GEMM layout, see
    weights.conv_weight_to_gemm), dy [N, H, W, Cout] (one 16-bit dtype) -> (dx [N,H,W,Cin] or None, dW [Cout, 9*Cin] fp32,
    dbias [Cout] fp32).  Cin % 64 == 0, Cout % 64 == 0."""
    assert x.is_cuda and x.dtype in ops.OPERAND_DTYPES and w.dtype == x.dtype and dy.dtype == x.dtype
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    assert w.shape == (Cout, 9 * Cin) and dy.shape == (N, H, W, Cout) and Cin % 64 == 0 and Cout % 64 == 0
    x, w, dy = x.contiguous(), w.contiguous(), dy.contiguous()
    dx = None
    if need_dx:
        w_rot = ops.weight_permute(w, Cout, 9, Cin, [8 - t for t in range(9)]).view(Cin, 9 * Cout)
        dx = ops.conv2d(dy, w_rot, ksize=3)
    dw = ops.conv_wgrad(x, dy, torch.empty((Cout, 9 * Cin), device=x.device, dtype=torch.float32), ksize=3)
    return dx, dw, ops.colsum(dy).view(Cout)
