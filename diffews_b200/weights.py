"""Weight re-layouts for the kernels (done once at model-load time, on whatever device the weights live on)."""
from __future__ import annotations

import torch


def conv_weight_to_gemm(w_oihw: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> [Cout, kh*kw*Cin] (K index = (kh*kw_size + kw)*Cin + c), the implicit-GEMM B operand."""
    co, ci, kh, kw = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()


def geglu_permute(w: torch.Tensor, b: torch.Tensor | None, block: int = 128):
    """GEGLU proj weight [8C, C] = [values(4C) ; gates(4C)] -> rows interleaved per 256-row tile as
    [128 values | 128 gates] so that one 128x256 accumulator tile holds a value and its gate (DFW_EPI_GEGLU)."""
    half = w.shape[0] // 2
    assert half % block == 0
    idx = []
    for t in range(half // block):
        idx.append(torch.arange(t * block, (t + 1) * block))
        idx.append(torch.arange(half + t * block, half + (t + 1) * block))
    idx = torch.cat(idx).to(w.device)
    wp = w.index_select(0, idx).contiguous()
    bp = b.index_select(0, idx).contiguous() if b is not None else None
    return wp, bp
