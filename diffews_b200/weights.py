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


def upconv_phase_weights(w_oihw: torch.Tensor) -> torch.Tensor:
    """Weights of `nearest-2x upsample -> 3x3 conv (pad 1)` collapsed per output phase.

    For output row 2i+ph the three kernel rows read input rows {i-1, i, i} (ph = 0) or {i, i, i+1} (ph = 1), so along
    each axis the 3 taps fold into 2:  ph = 0 -> [W0, W1+W2] at input offsets (-1, 0);  ph = 1 -> [W0+W1, W2] at
    offsets (0, +1).  Returns fp32 [4, Cout, 4*Cin] with phase = ph*2+pw and K index = (a*2+b)*Cin + c."""
    w = w_oihw.detach().float()
    co, ci = w.shape[0], w.shape[1]

    def fold(t, phase, dim):
        a, b, c = t.unbind(dim)
        return torch.stack([a, b + c], dim) if phase == 0 else torch.stack([a + b, c], dim)
    out = []
    for ph in range(2):
        for pw in range(2):
            f = fold(fold(w, ph, 2), pw, 3)                      # [co, ci, 2, 2]
            out.append(f.permute(0, 2, 3, 1).reshape(co, 4 * ci))
    return torch.stack(out, 0).contiguous()
