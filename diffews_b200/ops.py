"""Torch-tensor front end of the C ABI (device memory + current stream come from torch; the math does not).

Every function validates layout/dtype, passes raw pointers to libdiffews_b200.so and raises on a non-zero status.
Activations are channels-last: a conv input is a contiguous bf16 tensor of shape [N, H, W, C].
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import EPI_F16, EPI_GEGLU, EPI_OUT_F32, EPI_RES_F32, EPI_SILU, check, lib

bf16 = torch.bfloat16
f16 = torch.float16
OPERAND_DTYPES = (bf16, f16)      # 16-bit tensor-core operand formats; one format per call (A and B must match)


def _xd(t) -> int:
    """x_dtype code of the C ABI: 0 = bf16, 1 = fp32, 2 = fp16."""
    return {bf16: 0, torch.float32: 1, f16: 2}[t.dtype]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.DfwError(f"{name}: expected a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous tensor")
    return t


def launch_count() -> int:
    return int(lib.dfw_launch_count())


def set_option(option: int, value: int) -> int:
    """dfw_set_option: process-wide kernel-selection switch (_lib.OPT_*); returns the previous value."""
    old = int(lib.dfw_get_option(option))
    check(lib.dfw_set_option(option, int(value)), "dfw_set_option")
    return old


class KernelTimer:
    """CUDA-event timing of individual tensor-core launches (bench.py's roofline leg).  When installed with
    `set_timer`, conv2d / linear / attn_kvfused bracket their launch with two events on the launching stream."""

    def __init__(self):
        self.records = []   # (kind, flops, ev0, ev1)

    def add(self, kind, flops, ev0, ev1, shape=None):
        self.records.append((kind, flops, ev0, ev1, shape))

    def by_shape(self):
        out = {}
        for kind, flops, e0, e1, shape in self.records:
            d = out.setdefault((kind, shape), {"launches": 0, "flops": 0.0, "ms": 0.0})
            d["launches"] += 1; d["flops"] += flops; d["ms"] += e0.elapsed_time(e1)
        return out

    def summary(self):
        out = {}
        for kind, flops, e0, e1, _ in self.records:
            ms = e0.elapsed_time(e1)
            d = out.setdefault(kind, {"launches": 0, "flops": 0.0, "ms": 0.0})
            d["launches"] += 1; d["flops"] += flops; d["ms"] += ms
        return out


_timer = None


def set_timer(t):
    global _timer
    _timer = t


class _Timed:
    def __init__(self, kind, flops, shape=None):
        self.kind, self.flops, self.shape = kind, flops, shape

    def __enter__(self):
        if _timer is not None:
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _timer is not None:
            self.e1.record()
            _timer.add(self.kind, self.flops, self.e0, self.e1, self.shape)
        return False


def _nb(*ts) -> float:
    """algorithmic bytes of the HBM-bound kernels: every operand read or written once"""
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


MEM_KINDS = ("groupnorm", "layernorm", "softmax", "elementwise", "im2col", "cross_attn", "metric")   # work = bytes


# ---------------------------------------------------------------------------------------------------------------------
_sm_count = {}


def _sms(device) -> int:
    n = _sm_count.get(device)
    if n is None:
        n = torch.cuda.get_device_properties(device).multi_processor_count
        _sm_count[device] = n
    return n


def gn_stats_supported(N, Ho, Wo, Cout, out_f32, residual, bias_per_sample) -> bool:
    """Can the conv epilogue emit the GroupNorm(32) statistics of its output (dfw_conv2d_igemm_gnstats)?"""
    if bias_per_sample or (Cout * (4 if out_f32 else 2)) % 16:
        return False
    if residual is not None and (residual.dtype == torch.float32) != out_f32:
        return False
    return bool(lib.dfw_conv_gnstats_supported(N, Ho, Wo, Cout))      # group width in {4,8,16}, one image per tile


def conv2d(x, w, bias=None, *, ksize, stride=1, pad_mode=0, residual=None, out_scale=1.0, out_f32=False,
           silu=False, geglu=False, bias_per_sample=False, gn_stats=False):
    """x bf16 [N,H,W,Cin]; w bf16 [Cout, ks*ks*Cin]; bias fp32 [Cout] or [N,Cout]; returns [N,H/s,W/s,Cout_eff]."""
    assert x.dtype in OPERAND_DTYPES and w.dtype == x.dtype, (x.dtype, w.dtype)
    _req(x, x.dtype, "x"); _req(w, w.dtype, "w")
    h16 = x.dtype
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    assert w.numel() == Cout * ksize * ksize * Cin, (w.shape, Cin, ksize)
    flags = EPI_F16 if h16 == f16 else 0
    if out_f32: flags |= EPI_OUT_F32
    if silu: flags |= EPI_SILU
    if geglu: flags |= EPI_GEGLU
    cout_eff = Cout // 2 if geglu else Cout
    Ho, Wo = H // stride, W // stride
    y = torch.empty((N, Ho, Wo, cout_eff), device=x.device, dtype=torch.float32 if out_f32 else h16)
    bss = 0
    if bias is not None:
        _req(bias, torch.float32, "bias")
        if bias_per_sample:
            assert bias.shape == (N, Cout)
            bss = Cout
    if residual is not None:
        assert residual.shape == y.shape and residual.is_contiguous()
        if residual.dtype == torch.float32: flags |= EPI_RES_F32
        else: assert residual.dtype == h16
    want_gn = gn_stats and not (silu or geglu) and gn_stats_supported(N, Ho, Wo, Cout, out_f32, residual, bias_per_sample)
    tag = ""
    if _timer is not None:
        # which kernel the library will pick (igemm.cu try_launch_t128): only evaluated for the instrumented pass
        t128 = (stride == 1 and pad_mode == 0 and not (out_f32 or silu or geglu or bias_per_sample)
                and (residual is None or residual.dtype != torch.float32) and (not want_gn or Cout <= 512)
                and bool(lib.dfw_conv_t128_eligible(N, H, W, Cin, Cout, ksize)))
        tag = "t128 " if t128 else ""
    if want_gn:
        partial = torch.empty(int(lib.dfw_gn_partial_floats(N)), device=x.device, dtype=torch.float32)
        with _Timed("igemm", 2.0 * N * Ho * Wo * Cout * ksize * ksize * Cin,
                    f"{tag}conv N{N} {H}x{W} {Cin}->{Cout} k{ksize} s{stride} f{flags} +gn"):
            check(lib.dfw_conv2d_igemm_gnstats(x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), N,
                                               H, W, Cin, Cout, ksize, stride, pad_mode, flags, float(out_scale),
                                               partial.data_ptr(), _stream()), "dfw_conv2d_igemm_gnstats")
        y._gn_partial = (partial, partial.numel() // (N * 64))
        return y
    with _Timed("igemm", 2.0 * N * Ho * Wo * Cout * ksize * ksize * Cin,
                f"{tag}conv N{N} {H}x{W} {Cin}->{Cout} k{ksize} s{stride} f{flags}"):
        check(lib.dfw_conv2d_igemm(x.data_ptr(), w.data_ptr(), _ptr(bias), bss, _ptr(residual), y.data_ptr(), N, H, W,
                                   Cin, Cout, ksize, stride, pad_mode, flags, float(out_scale), _stream()),
              "dfw_conv2d_igemm")
    return y


def conv_gn_in_supported(x, cout, ksize) -> bool:
    """Can conv2d_gn_in consume x (which must carry the statistics its producing conv emitted)?"""
    if getattr(x, "_gn_partial", None) is None or x.dtype not in OPERAND_DTYPES or x.dim() != 4:
        return False
    N, H, W, Cin = x.shape
    return bool(lib.dfw_conv_gnin_supported(N, H, W, Cin, cout, ksize))


_gnin_scratch = {}


def _gnin_scratch_for(device):
    """The ring of transformed tiles dfw_conv2d_igemm_gnin keeps in L2: one per (device, stream), allocated once."""
    key = (device.index, _stream())
    buf = _gnin_scratch.get(key)
    if buf is None:
        buf = _gnin_scratch[key] = torch.empty(int(lib.dfw_conv_gnin_scratch_bytes()), device=device, dtype=torch.uint8)
    return buf


def conv2d_gn_in(x, gamma, beta, eps, w, bias=None, *, ksize, residual=None, gn_stats=False, groups=32):
    """conv(silu(groupnorm(x))) + bias (+ residual) with the normalisation applied to the conv operand on the fly:
    x 16-bit [N,H,W,Cin] with x._gn_partial (its statistics, from the conv that produced it); returns 16-bit [N,H,W,Cout]."""
    assert x.dtype in OPERAND_DTYPES and w.dtype == x.dtype and x.is_contiguous()
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    partial, nchunks = x._gn_partial
    ss = torch.empty((N, 2, Cin), device=x.device, dtype=torch.float32)
    check(lib.dfw_gn_scale_shift(partial.data_ptr(), nchunks, gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), N, H * W, Cin,
                                 groups, float(eps), _stream()), "dfw_gn_scale_shift")
    y = torch.empty((N, H, W, Cout), device=x.device, dtype=x.dtype)
    flags = EPI_F16 if x.dtype == f16 else 0
    if residual is not None:
        assert residual.shape == y.shape and residual.is_contiguous() and residual.dtype == x.dtype
    if bias is not None: _req(bias, torch.float32, "bias")
    part_out = None
    if gn_stats and gn_stats_supported(N, H, W, Cout, False, residual, False):
        part_out = torch.empty(int(lib.dfw_gn_partial_floats(N)), device=x.device, dtype=torch.float32)
    with _Timed("igemm", 2.0 * N * H * W * Cout * ksize * ksize * Cin,
                f"conv N{N} {H}x{W} {Cin}->{Cout} k{ksize} s1 gn-in" + (" +gn" if part_out is not None else "")):
        check(lib.dfw_conv2d_igemm_gnin(x.data_ptr(), ss.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(),
                                        N, H, W, Cin, Cout, ksize, flags, _ptr(part_out),
                                        _gnin_scratch_for(x.device).data_ptr(), _stream()),
              "dfw_conv2d_igemm_gnin")
    if part_out is not None:
        y._gn_partial = (part_out, part_out.numel() // (N * 64))
    return y


def linear(x, w, bias=None, *, residual=None, out_scale=1.0, out_f32=False, silu=False, geglu=False, out=None):
    """x bf16 [..., K]; w bf16 [Nout, K]; returns [..., Nout_eff] (written into `out` if given)."""
    assert x.dtype in OPERAND_DTYPES and w.dtype == x.dtype, (x.dtype, w.dtype)
    _req(x, x.dtype, "x"); _req(w, w.dtype, "w")
    h16 = x.dtype
    K = x.shape[-1]
    M = x.numel() // K
    Nout = w.shape[0]
    assert w.shape[1] == K
    flags = EPI_F16 if h16 == f16 else 0
    if out_f32: flags |= EPI_OUT_F32
    if silu: flags |= EPI_SILU
    if geglu: flags |= EPI_GEGLU
    nout_eff = Nout // 2 if geglu else Nout
    odt = torch.float32 if out_f32 else h16
    if out is None:
        y = torch.empty(x.shape[:-1] + (nout_eff,), device=x.device, dtype=odt)
    else:
        y = out
        assert y.is_contiguous() and y.dtype == odt and y.numel() == M * nout_eff and y.is_cuda
    if bias is not None: _req(bias, torch.float32, "bias")
    if residual is not None:
        assert residual.numel() == y.numel() and residual.is_contiguous()
        if residual.dtype == torch.float32: flags |= EPI_RES_F32
        else: assert residual.dtype == h16
    tag = ""
    if _timer is not None and residual is not None and M % 256 == 0 and not (out_f32 or silu or geglu) \
            and residual.dtype != torch.float32 and bool(lib.dfw_conv_t128_eligible(1, M // 16, 16, K, Nout, 1)):
        tag = "t128 "                  # residual projections run on the channel-major kernel (igemm.cu try_launch_t128)
    with _Timed("igemm", 2.0 * M * K * Nout, f"{tag}linear M{M} K{K} N{Nout} f{flags}"):
        check(lib.dfw_linear(x.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(residual), y.data_ptr(), M, K, Nout, flags,
                             float(out_scale), _stream()), "dfw_linear")
    return y


def upconv2x(x, w4, bias=None, *, out_f32=False, gn_stats=False):
    """nearest-2x upsample + 3x3 conv, fused: x 16-bit [N,H,W,Cin]; w4 [4,Cout,4*Cin] -> [N,2H,2W,Cout]."""
    assert x.dtype in OPERAND_DTYPES and w4.dtype == x.dtype
    _req(x, x.dtype, "x"); _req(w4, w4.dtype, "w4")
    N, H, W, Cin = x.shape
    Cout = w4.shape[1]
    assert w4.shape == (4, Cout, 4 * Cin)
    flags = (EPI_F16 if x.dtype == f16 else 0) | (EPI_OUT_F32 if out_f32 else 0)
    y = torch.empty((N, 2 * H, 2 * W, Cout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    if bias is not None: _req(bias, torch.float32, "bias")
    partial = None
    if gn_stats and gn_stats_supported(N, H, W, Cout, out_f32, None, False):
        partial = torch.empty(4 * int(lib.dfw_gn_partial_floats(N)), device=x.device, dtype=torch.float32)
    with _Timed("igemm", 2.0 * N * H * W * 16 * Cin * Cout, f"upconv N{N} {H}x{W} {Cin}->{Cout} f{flags}"):
        check(lib.dfw_upconv2x_igemm(x.data_ptr(), w4.data_ptr(), _ptr(bias), y.data_ptr(), N, H, W, Cin, Cout,
                                     flags, _ptr(partial), _stream()), "dfw_upconv2x_igemm")
    if partial is not None:
        y._gn_partial = (partial, partial.numel() // (N * 64))
    return y


def bmm_nt(x, w, bias=None, *, out_f32=False, out_scale=1.0):
    """y[b] = x[b] @ w[b]^T (+ bias).  x [B, M, K] contiguous; w [B, Nout, K] with unit stride on K and arbitrary
    row / batch strides (so a column slice of a wider buffer works)."""
    assert x.dtype in OPERAND_DTYPES and w.dtype == x.dtype and x.is_cuda and x.is_contiguous()
    B, M, K = x.shape
    assert w.shape[0] == B and w.shape[2] == K and w.stride(2) == 1
    Nout = w.shape[1]
    flags = (EPI_F16 if x.dtype == f16 else 0) | (EPI_OUT_F32 if out_f32 else 0)
    y = torch.empty((B, M, Nout), device=x.device, dtype=torch.float32 if out_f32 else x.dtype)
    if bias is not None: _req(bias, torch.float32, "bias")
    with _Timed("igemm", 2.0 * B * M * K * Nout, f"bmm B{B} M{M} K{K} N{Nout} f{flags}"):
        check(lib.dfw_bmm_nt(x.data_ptr(), w.data_ptr(), w.stride(1), w.stride(0), _ptr(bias), y.data_ptr(), B, M, K,
                             Nout, flags, float(out_scale), _stream()), "dfw_bmm_nt")
    return y


def attn_kvfused(q, k_self, v_self, k_bank, v_bank, heads, scale, return_lse=False):
    """q [B,Lq,C]; k_self/v_self [B,Ls,C]; k_bank/v_bank [B,Lb,C] or None.  Tensors may be column slices of a wider
    (e.g. fused QKV) buffer: only the last dim must be unit-stride.  `return_lse`: also return the fp32 [B, heads, Lq]
    log2-domain logsumexp of the scaled logits (what attn_kvfused_backward needs)."""
    B, Lq, C = q.shape
    assert C == heads * 64
    h16 = q.dtype
    assert h16 in OPERAND_DTYPES
    for t in (q, k_self, v_self):
        assert t.dtype == h16 and t.is_cuda and t.stride(2) == 1
    Ls = k_self.shape[1]
    assert k_self.stride() == v_self.stride()
    o = torch.empty((B, Lq, C), device=q.device, dtype=h16)
    if k_bank is not None:
        Lb = k_bank.shape[1]
        assert k_bank.shape[0] == B and k_bank.stride() == v_bank.stride() and k_bank.stride(2) == 1
        assert k_bank.dtype == h16 and v_bank.dtype == h16
        kb, vb, kbs, krs = k_bank.data_ptr(), v_bank.data_ptr(), k_bank.stride(0), k_bank.stride(1)
    else:
        Lb, kb, vb, kbs, krs = 0, 0, 0, 0, 0
    with _Timed("attn", 4.0 * B * heads * Lq * (Ls + Lb) * 64, f"attn B{B} h{heads} Lq{Lq} Lk{Ls + Lb}"):
        if return_lse:
            lse = torch.empty((B, heads, Lq), device=q.device, dtype=torch.float32)
            check(lib.dfw_attn_kvfused_fwd_lse(q.data_ptr(), q.stride(0), q.stride(1), k_self.data_ptr(), v_self.data_ptr(),
                                               k_self.stride(0), k_self.stride(1), kb, vb, kbs, krs, o.data_ptr(), o.stride(0),
                                               o.stride(1), B, heads, Lq, Ls, Lb, float(scale), int(h16 == f16),
                                               lse.data_ptr(), _stream()), "dfw_attn_kvfused_fwd_lse")
            return o, lse
        check(lib.dfw_attn_kvfused_fwd(q.data_ptr(), q.stride(0), q.stride(1), k_self.data_ptr(), v_self.data_ptr(),
                                       k_self.stride(0), k_self.stride(1), kb, vb, kbs, krs, o.data_ptr(), o.stride(0),
                                       o.stride(1), B, heads, Lq, Ls, Lb, float(scale), int(h16 == f16), _stream()),
              "dfw_attn_kvfused_fwd")
    return o


def attn_kvfused_backward(q, k_self, v_self, k_bank, v_bank, o, d_o, heads, scale, lse=None):
    """Backward of attn_kvfused: contiguous [B, L, heads*64] tensors of one 16-bit dtype (o = forward output, d_o its
    gradient, lse = the statistics of attn_kvfused(..., return_lse=True); without them one extra forward recomputes them).
    Returns (dq, dk_self, dv_self, dk_bank, dv_bank); the bank gradients are None without a bank."""
    B, Lq, C = q.shape
    assert C == heads * 64 and q.dtype in OPERAND_DTYPES
    h16 = q.dtype
    Ls = k_self.shape[1]
    Lb = 0 if k_bank is None else k_bank.shape[1]
    ts = [q, k_self, v_self, o, d_o] + ([k_bank, v_bank] if Lb else [])
    for t in ts:
        assert t.dtype == h16 and t.is_cuda and t.is_contiguous()
    assert o.shape == q.shape and d_o.shape == q.shape and v_self.shape == k_self.shape
    dq = torch.empty_like(q); dks = torch.empty_like(k_self); dvs = torch.empty_like(v_self)
    dkb = torch.empty_like(k_bank) if Lb else None
    dvb = torch.empty_like(v_bank) if Lb else None
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and tuple(lse.shape) == (B, heads, Lq)
    ws = torch.empty(int(lib.dfw_attn_bwd_workspace_bytes(B, heads, Lq, Ls, Lb)), device=q.device, dtype=torch.uint8)
    kb = (k_bank.data_ptr(), v_bank.data_ptr(), k_bank.stride(0), k_bank.stride(1)) if Lb else (0, 0, 0, 0)
    with _Timed("attn", 2.5 * 4.0 * B * heads * Lq * (Ls + Lb) * 64, f"attn-bwd B{B} h{heads} Lq{Lq} Lk{Ls + Lb}"):
        check(lib.dfw_attn_kvfused_bwd(q.data_ptr(), q.stride(0), q.stride(1), k_self.data_ptr(), v_self.data_ptr(),
                                       k_self.stride(0), k_self.stride(1), kb[0], kb[1], kb[2], kb[3], o.data_ptr(),
                                       d_o.data_ptr(), o.stride(0), o.stride(1), _ptr(lse), dq.data_ptr(), dks.data_ptr(), dvs.data_ptr(),
                                       _ptr(dkb), _ptr(dvb), B, heads, Lq, Ls, Lb, float(scale), int(h16 == f16), ws.data_ptr(),
                                       _stream()), "dfw_attn_kvfused_bwd")
    return dq, dks, dvs, dkb, dvb


def cross_attn(q, k, v, heads, scale):
    """q [B,L,C] bf16; k,v [Bk,Lctx,C] with Bk in {1,B}."""
    h16 = q.dtype
    assert h16 in OPERAND_DTYPES
    _req(q, h16, "q"); _req(k, h16, "k"); _req(v, h16, "v")
    B, L, C = q.shape
    Lctx = k.shape[1]
    kvs = 0 if k.shape[0] == 1 else Lctx * C
    o = torch.empty_like(q)
    with _Timed("cross_attn", _nb(q, o), f"cross_attn B{B} L{L} h{heads} ctx{Lctx}"):
        check(lib.dfw_cross_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), kvs, o.data_ptr(), B, L, heads, Lctx,
                                     float(scale), int(h16 == f16), _stream()), "dfw_cross_attn_fwd")
    return o


_gn_ws = {}


def groupnorm(x, gamma, beta, *, groups=32, eps=1e-5, silu=False, out_dtype=bf16):
    """x bf16|fp16|fp32 [N, ..., C] channels-last; returns bf16 (or fp16) of the same shape."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (bf16, f16, torch.float32)
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    pre = getattr(x, "_gn_partial", None)
    if pre is not None and groups == 32:
        # the producing convolution already reduced the statistics in its epilogue: normalise-and-store pass only
        partial, nchunks = pre
        y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        with _Timed("groupnorm", _nb(x, y), f"gn-apply N{N} HW{HW} C{C} {x.dtype}".replace("torch.", "")):
            check(lib.dfw_groupnorm_from_partial(x.data_ptr(), _xd(x), partial.data_ptr(), nchunks, gamma.data_ptr(),
                                                 beta.data_ptr(), y.data_ptr(), int(out_dtype == f16), N, HW, C, groups,
                                                 float(eps), int(silu), _stream()), "dfw_groupnorm_from_partial")
        return y
    need = int(lib.dfw_groupnorm_workspace_bytes(N, HW, C, groups))
    key = (x.device, torch.cuda.current_stream().cuda_stream)
    ws = _gn_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 20), device=x.device, dtype=torch.uint8)
        _gn_ws[key] = ws
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    with _Timed("groupnorm", _nb(x, y), f"gn-full N{N} HW{HW} C{C} {x.dtype}".replace("torch.", "")):
        check(lib.dfw_groupnorm_silu(x.data_ptr(), _xd(x), gamma.data_ptr(), beta.data_ptr(),
                                     y.data_ptr(), int(out_dtype == f16), N, HW, C, groups, float(eps), int(silu),
                                     ws.data_ptr(), _stream()),
              "dfw_groupnorm_silu")
    return y


def cross_attn_collapsed(logits, U, bias, residual, heads, lctx, out_dtype):
    """logits fp32 [B, L, >= heads*lctx]; U fp32 [heads*lctx, C]; bias fp32 [C]; residual [B, L, C] or None -> [B, L, C]."""
    _req(logits, torch.float32, "logits"); _req(U, torch.float32, "U"); _req(bias, torch.float32, "bias")
    B, L, ld = logits.shape
    C = U.shape[1]
    out = torch.empty((B, L, C), device=logits.device, dtype=out_dtype)
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == out_dtype and residual.is_contiguous()
    with _Timed("cross_attn", _nb(logits, out, residual), f"cross_attn_collapsed B{B} L{L} h{heads} ctx{lctx}"):
        check(lib.dfw_cross_attn_collapsed(logits.data_ptr(), ld, U.data_ptr(), bias.data_ptr(), _ptr(residual),
                                           out.data_ptr(), _xd(out), B * L, C, heads, lctx, _stream()),
              "dfw_cross_attn_collapsed")
    return out


def groupnorm_backward(x, dy, gamma, beta, *, groups=32, eps=1e-5, silu=False):
    """Backward of y = [silu](group_norm(x)): x, dy [N, ..., C] channels-last of one dtype (bf16 | fp16 | fp32);
    returns (dx like x, dgamma fp32 [C], dbeta fp32 [C])."""
    assert x.is_cuda and x.is_contiguous() and dy.is_contiguous() and dy.shape == x.shape and dy.dtype == x.dtype
    assert x.dtype in (bf16, f16, torch.float32)
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    ws = torch.empty(int(lib.dfw_groupnorm_bwd_workspace_bytes(N, HW, C, groups)), device=x.device, dtype=torch.uint8)
    dx = torch.empty_like(x)
    dg = torch.empty(C, device=x.device, dtype=torch.float32); db = torch.empty(C, device=x.device, dtype=torch.float32)
    with _Timed("groupnorm", 2 * _nb(x, dy) + _nb(dx), f"gn-bwd N{N} HW{HW} C{C} {x.dtype}".replace("torch.", "")):
        check(lib.dfw_groupnorm_silu_bwd(x.data_ptr(), dy.data_ptr(), _xd(x), gamma.data_ptr(), beta.data_ptr(),
                                         dx.data_ptr(), dg.data_ptr(), db.data_ptr(), N, HW, C, groups, float(eps), int(silu),
                                         ws.data_ptr(), _stream()), "dfw_groupnorm_silu_bwd")
    return dx, dg, db


def layernorm(x, gamma, beta, eps=1e-5, out_dtype=bf16):
    assert x.is_cuda and x.is_contiguous() and x.dtype in (bf16, f16, torch.float32)
    C = x.shape[-1]
    M = x.numel() // C
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    with _Timed("layernorm", _nb(x, y), f"ln M{M} C{C}"):
        check(lib.dfw_layernorm(x.data_ptr(), _xd(x), gamma.data_ptr(), beta.data_ptr(),
                                y.data_ptr(), int(out_dtype == f16), M, C, float(eps), _stream()), "dfw_layernorm")
    return y


def layernorm_backward(x, dy, gamma, eps=1e-5):
    """Backward of `layernorm` (statistics recomputed from x).  x, dy [.., C] of one dtype (bf16 / fp16 / fp32);
    returns (dx like x, dgamma fp32 [C], dbeta fp32 [C])."""
    assert x.is_cuda and x.is_contiguous() and dy.is_contiguous() and x.dtype == dy.dtype and x.shape == dy.shape
    C = x.shape[-1]
    M = x.numel() // C
    dx = torch.empty_like(x)
    dg = torch.empty(C, device=x.device, dtype=torch.float32)
    db = torch.empty(C, device=x.device, dtype=torch.float32)
    ws = torch.empty(int(lib.dfw_layernorm_bwd_workspace_bytes(M, C)), device=x.device, dtype=torch.uint8)
    check(lib.dfw_layernorm_bwd(x.data_ptr(), dy.data_ptr(), _xd(x), gamma.data_ptr(), dx.data_ptr(), dg.data_ptr(),
                                db.data_ptr(), M, C, float(eps), ws.data_ptr(), _stream()), "dfw_layernorm_bwd")
    return dx, dg, db


def geglu_backward(h, dy):
    """Backward of diffusers GEGLU's activation: h [.., 2F] = (value | gate) pre-activations, dy [.., F] -> dh like h."""
    assert h.is_cuda and h.is_contiguous() and dy.is_contiguous() and h.dtype == dy.dtype
    F = dy.shape[-1]
    assert h.shape[-1] == 2 * F and h.shape[:-1] == dy.shape[:-1]
    dh = torch.empty_like(h)
    check(lib.dfw_geglu_bwd(h.data_ptr(), dy.data_ptr(), dh.data_ptr(), _xd(h), dy.numel() // F, F, _stream()),
          "dfw_geglu_bwd")
    return dh


def softmax_rows(s, scale, out_dtype=bf16):
    _req(s, torch.float32, "s")
    L = s.shape[-1]
    M = s.numel() // L
    p = torch.empty(s.shape, device=s.device, dtype=out_dtype)
    with _Timed("softmax", _nb(s, p), f"softmax M{M} L{L}"):
        check(lib.dfw_softmax_rows(s.data_ptr(), p.data_ptr(), int(out_dtype == f16), M, L, float(scale), _stream()),
              "dfw_softmax_rows")
    return p


# ---- fp32 evaluation mode (csrc/f32mode.cu): split operands for the tensor-core GEMMs, fp32 everything else ------------
def split3(x, role=0, dtype=f16):
    """fp32 [..., C] -> 16-bit [..., 3C]: role 0 = [hi | lo | hi] (activation side), role 1 = [hi | hi | lo] (weight
    side of an activation x activation product).  `x` may be a column slice of a wider contiguous buffer."""
    assert x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1 and dtype in OPERAND_DTYPES
    C = x.shape[-1]
    rows = x.numel() // C
    if x.is_contiguous():
        xs = C
    else:
        xs = x.stride(-2)
        for d in range(x.dim() - 2):
            assert x.stride(d) == x.stride(d + 1) * x.shape[d + 1], "split3: rows must have one uniform stride"
    y = torch.empty(x.shape[:-1] + (3 * C,), device=x.device, dtype=dtype)
    with _Timed("elementwise", _nb(x, y), f"split3 rows{rows} C{C}"):
        check(lib.dfw_split3_16(x.data_ptr(), xs, y.data_ptr(), rows, C, int(role), int(dtype == f16), _stream()),
              "dfw_split3_16")
    return y


def split3_host(w: torch.Tensor, taps: int, role: int, dtype) -> torch.Tensor:
    """Weight-side split at load time (torch on the host tensor, once): w [rows, taps*C] fp32 -> [rows, taps*3C]."""
    rows = w.shape[0]
    w3 = w.detach().float().reshape(rows, taps, -1)
    hi = w3.to(dtype)
    lo = (w3 - hi.float()).to(dtype)
    parts = [hi, hi, lo] if role == 1 else [hi, lo, hi]
    return torch.cat(parts, dim=-1).reshape(rows, -1).contiguous()


def groupnorm_f32(x, gamma, beta, *, groups=32, eps=1e-5, silu=False):
    _req(x, torch.float32, "x")
    N, C = x.shape[0], x.shape[-1]
    HW = x.numel() // (N * C)
    y = torch.empty_like(x)
    with _Timed("groupnorm", _nb(x, y), f"gn-f32 N{N} HW{HW} C{C}"):
        check(lib.dfw_groupnorm_f32(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), N, HW, C, groups,
                                    float(eps), int(silu), _stream()), "dfw_groupnorm_f32")
    return y


def layernorm_f32(x, gamma, beta, eps=1e-5):
    _req(x, torch.float32, "x")
    C = x.shape[-1]
    y = torch.empty_like(x)
    with _Timed("layernorm", _nb(x, y), f"ln-f32 M{x.numel() // C} C{C}"):
        check(lib.dfw_layernorm_f32(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), x.numel() // C, C,
                                    float(eps), _stream()), "dfw_layernorm_f32")
    return y


def softmax_rows_f32(s, scale):
    """In place: s fp32 [.., L] <- softmax(s * scale)."""
    _req(s, torch.float32, "s")
    L = s.shape[-1]
    with _Timed("softmax", 2 * _nb(s), f"softmax-f32 M{s.numel() // L} L{L}"):
        check(lib.dfw_softmax_rows_f32(s.data_ptr(), s.data_ptr(), s.numel() // L, L, float(scale), _stream()),
              "dfw_softmax_rows_f32")
    return s


def geglu_f32(h):
    _req(h, torch.float32, "h")
    F = h.shape[-1] // 2
    y = torch.empty(h.shape[:-1] + (F,), device=h.device, dtype=torch.float32)
    with _Timed("elementwise", _nb(h, y), f"geglu-f32 rows{y.numel() // F} F{F}"):
        check(lib.dfw_geglu_f32(h.data_ptr(), y.data_ptr(), y.numel() // F, F, _stream()), "dfw_geglu_f32")
    return y


def attn_f32(q, k_self, v_self, k_bank, v_bank, heads, scale):
    """fp32 attention, head dim 64: q [B,Lq,C]; k_self / v_self [Bk,Ls,C] (Bk = 1 shares them over the batch);
    k_bank / v_bank [B,Lb,C] or None.  Tensors may be column slices of wider buffers (unit stride on the last dim)."""
    B, Lq, C = q.shape
    assert C == heads * 64
    for t in (q, k_self, v_self):
        assert t.dtype == torch.float32 and t.is_cuda and t.stride(2) == 1
    assert k_self.stride() == v_self.stride() and k_self.shape[0] in (1, B)
    Ls = k_self.shape[1]
    ksb = 0 if k_self.shape[0] == 1 else k_self.stride(0)
    o = torch.empty((B, Lq, C), device=q.device, dtype=torch.float32)
    if k_bank is not None:
        Lb = k_bank.shape[1]
        assert k_bank.shape[0] == B and k_bank.stride() == v_bank.stride() and k_bank.stride(2) == 1
        kb, vb, kbs, krs = k_bank.data_ptr(), v_bank.data_ptr(), k_bank.stride(0), k_bank.stride(1)
    else:
        Lb, kb, vb, kbs, krs = 0, 0, 0, 0, 0
    with _Timed("attn_f32", 4.0 * B * heads * Lq * (Ls + Lb) * 64, f"attn-f32 B{B} h{heads} Lq{Lq} Lk{Ls + Lb}"):
        check(lib.dfw_attn_f32(q.data_ptr(), q.stride(0), q.stride(1), k_self.data_ptr(), v_self.data_ptr(), ksb,
                               k_self.stride(1), kb, vb, kbs, krs, o.data_ptr(), o.stride(0), o.stride(1), B, heads, Lq, Ls,
                               Lb, float(scale), _stream()), "dfw_attn_f32")
    return o


def upsample2x(x, out_dtype=bf16):
    """x 16-bit|fp32 [N,H,W,C] -> 16-bit [N,2H,2W,C] (fp32 input is cast to `out_dtype`)."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (bf16, f16, torch.float32)
    if x.dtype != torch.float32:
        out_dtype = x.dtype
    N, H, W, Cc = x.shape
    y = torch.empty((N, 2 * H, 2 * W, Cc), device=x.device, dtype=out_dtype)
    with _Timed("elementwise", _nb(x, y), f"upsample2x N{N} {H}x{W} C{Cc}"):
        check(lib.dfw_upsample2x_nhwc(x.data_ptr(), int(x.dtype == torch.float32), y.data_ptr(), int(out_dtype == f16),
                                      N, H, W, Cc, _stream()), "dfw_upsample2x_nhwc")
    return y


def concat_channels(a, b):
    assert a.is_cuda and a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype
    assert a.dtype in (bf16, f16, torch.float32) and a.shape[:-1] == b.shape[:-1]
    Ca, Cb = a.shape[-1], b.shape[-1]
    rows = a.numel() // Ca
    y = torch.empty(a.shape[:-1] + (Ca + Cb,), device=a.device, dtype=a.dtype)
    with _Timed("elementwise", 2 * _nb(y), f"concat rows{rows} {Ca}+{Cb} x{a.element_size()}B"):
        check(lib.dfw_concat_channels(a.data_ptr(), b.data_ptr(), y.data_ptr(), rows, Ca, Cb, a.element_size(),
                                      _stream()), "dfw_concat_channels")
    return y


def cast16(x, dtype=bf16):
    """fp32 -> bf16 / fp16 (no-op for an input that already has `dtype`)."""
    if x.dtype == dtype:
        return x
    _req(x, torch.float32, "x")
    y = torch.empty(x.shape, device=x.device, dtype=dtype)
    with _Timed("elementwise", _nb(x, y), f"cast16 n{x.numel()}"):
        check(lib.dfw_cast_f32_to_16(x.data_ptr(), y.data_ptr(), int(dtype == f16), x.numel(), _stream()),
              "dfw_cast_f32_to_16")
    return y


def cast_bf16(x):
    return cast16(x, bf16)


def conv3x3_small_cin(x_nchw, w, bias, out_dtype=bf16):
    """x fp32 NCHW [N,Cin<=8,H,W]; w fp32 [Cout,3,3,Cin]; returns NHWC [N,H,W,Cout] in bf16 / fp16 / fp32."""
    _req(x_nchw, torch.float32, "x"); _req(w, torch.float32, "w")
    N, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    y = torch.empty((N, H, W, Cout), device=x_nchw.device, dtype=out_dtype)
    with _Timed("elementwise", _nb(x_nchw, y), f"conv3x3_small N{N} {H}x{W} {Cin}->{Cout}"):
        check(lib.dfw_conv3x3_small_cin(x_nchw.data_ptr(), w.data_ptr(), _ptr(bias), y.data_ptr(), _xd(y), N, H, W,
                                        Cin, Cout, _stream()), "dfw_conv3x3_small_cin")
    return y


def im2col3x3_small(x_nchw, kpad, dtype=bf16):
    """x fp32 NCHW [N,Cin,H,W] -> 16-bit [N,H,W,kpad] im2col rows (k = tap*Cin + c, zero padded)."""
    _req(x_nchw, torch.float32, "x")
    N, Cin, H, W = x_nchw.shape
    y = torch.empty((N, H, W, kpad), device=x_nchw.device, dtype=dtype)
    with _Timed("im2col", _nb(x_nchw, y), f"im2col N{N} {H}x{W} Cin{Cin} K{kpad}"):
        check(lib.dfw_im2col3x3_small(x_nchw.data_ptr(), y.data_ptr(), int(dtype == f16), N, H, W, Cin, kpad, _stream()),
              "dfw_im2col3x3_small")
    return y


def pointwise_small(x, x_strides, w_host, b_host, y, y_strides, N, HW, in_scale=1.0, out_scale=1.0):
    """w_host [Cout,Cin] / b_host [Cout]: fp32 CPU tensors (kernel parameters). x, y fp32 CUDA tensors."""
    assert w_host.device.type == "cpu" and w_host.dtype == torch.float32 and w_host.is_contiguous()
    Cout, Cin = w_host.shape
    bptr = 0
    if b_host is not None:
        assert b_host.device.type == "cpu" and b_host.dtype == torch.float32
        bptr = b_host.data_ptr()
    with _Timed("elementwise", float(N) * HW * (Cin + Cout) * 4, f"pointwise N{N} HW{HW} {Cin}->{Cout}"):
        check(lib.dfw_pointwise_small(x.data_ptr(), *x_strides, w_host.data_ptr(), bptr, float(in_scale),
                                      float(out_scale), y.data_ptr(), *y_strides, N, HW, Cin, Cout, _stream()),
              "dfw_pointwise_small")
    return y


def nhwc_f32_to_nchw(x, C, H, W, scale=1.0, shift=0.0, lo=-3.0e38, hi=3.0e38):
    """x fp32 [N, H*W, row_stride>=C] -> fp32 [N, C, H, W]."""
    _req(x, torch.float32, "x")
    N = x.shape[0]
    row_stride = x.shape[-1]
    y = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
    with _Timed("elementwise", 2 * _nb(y), f"nhwc->nchw N{N} C{C} {H}x{W}"):
        check(lib.dfw_nhwc_f32_to_nchw_f32(x.data_ptr(), row_stride, y.data_ptr(), N, C, H * W, float(scale),
                                           float(shift), float(lo), float(hi), _stream()), "dfw_nhwc_f32_to_nchw_f32")
    return y


def seg_post(dec, H, W, want_f32=True, want_u8=True):
    """dec fp32 [N, H*W, row_stride>=3] -> (seg_f32 [N,3,H,W] in [0,255], seg_u8 [N,3,H,W])."""
    _req(dec, torch.float32, "dec")
    N = dec.shape[0]
    rs = dec.shape[-1]
    f = torch.empty((N, 3, H, W), device=dec.device, dtype=torch.float32) if want_f32 else None
    u = torch.empty((N, 3, H, W), device=dec.device, dtype=torch.uint8) if want_u8 else None
    with _Timed("elementwise", float(N) * H * W * 3 * 4 + _nb(f, u), f"seg_post N{N} {H}x{W}"):
        check(lib.dfw_seg_post(dec.data_ptr(), rs, _ptr(f), _ptr(u), N, H * W, _stream()), "dfw_seg_post")
    return f, u


def seg_head_prepare(w_oihw: torch.Tensor, dtype, device) -> torch.Tensor:
    """conv_out weight [3,128,3,3] -> the mma.sync B fragments dfw_seg_head_u8 reads (uint32 device tensor, int32 view)."""
    w = w_oihw.detach().float().cpu().contiguous()
    assert tuple(w.shape) == (3, 128, 3, 3)
    out = torch.empty(int(lib.dfw_seg_head_weight_u32()), dtype=torch.int32)
    check(lib.dfw_seg_head_prepare_weights(w.data_ptr(), int(dtype == f16), out.data_ptr()), "dfw_seg_head_prepare_weights")
    return out.to(device)


def seg_head_supported(x) -> bool:
    """dfw_seg_head_u8 needs the decoder's real last block: 128 channels, 16-bit, statistics from the producing conv."""
    return (x.dim() == 4 and x.shape[-1] == 128 and x.dtype in OPERAND_DTYPES and x.shape[2] % 16 == 0
            and getattr(x, "_gn_partial", None) is not None and bool(lib.dfw_get_option(_lib.OPT_SEG_HEAD)))


def seg_head_u8(x, gamma, beta, eps, wb, bias_host, want_f32=False, want_u8=True):
    """Fused decoder head: x 16-bit [N,H,W,128] carrying x._gn_partial -> (seg_f32 [N,3,H,W] in [0,255] | None,
    seg_u8 [N,3,H,W] | None).  GroupNorm(32) scale / shift come from the statistics the producing conv emitted."""
    N, H, W, C = x.shape
    partial, nchunks = x._gn_partial
    ss = torch.empty((N, 2, C), device=x.device, dtype=torch.float32)
    check(lib.dfw_gn_scale_shift(partial.data_ptr(), nchunks, gamma.data_ptr(), beta.data_ptr(), ss.data_ptr(), N, H * W, C,
                                 32, float(eps), _stream()), "dfw_gn_scale_shift")
    f = torch.empty((N, 3, H, W), device=x.device, dtype=torch.float32) if want_f32 else None
    u = torch.empty((N, 3, H, W), device=x.device, dtype=torch.uint8) if want_u8 else None
    assert bias_host.device.type == "cpu" and bias_host.dtype == torch.float32 and bias_host.numel() == 3
    with _Timed("elementwise", _nb(x, f, u), f"seg_head N{N} {H}x{W}"):
        check(lib.dfw_seg_head_u8(x.data_ptr(), int(x.dtype == f16), ss.data_ptr(), wb.data_ptr(), bias_host.data_ptr(),
                                  _ptr(u), _ptr(f), N, H, W, _stream()), "dfw_seg_head_u8")
    return f, u


def rthres_iou_hist(pred_u8, gt_u8, ignore_u8=None, r_threshold=0.25, want_mask=True):
    """pred_u8 [B,3,H,W] uint8 (or a binarised [B,H,W] {0,1} mask); gt_u8 [B,H,W] uint8 {0,1};
    returns (area_inter [B,2] i64, area_union [B,2] i64, mask)."""
    _req(pred_u8, torch.uint8, "pred"); _req(gt_u8, torch.uint8, "gt")
    is_mask = pred_u8.ndim == 3
    B, H, W = pred_u8.shape[0], pred_u8.shape[-2], pred_u8.shape[-1]
    assert gt_u8.shape == (B, H, W)
    dev = pred_u8.device
    inter = torch.empty((B, 2), device=dev, dtype=torch.int64)
    union = torch.empty((B, 2), device=dev, dtype=torch.int64)
    mask = torch.empty((B, H, W), device=dev, dtype=torch.uint8) if want_mask else None
    ws = torch.empty(int(lib.dfw_rthres_workspace_bytes(B)), device=dev, dtype=torch.uint8)
    if ignore_u8 is not None: _req(ignore_u8, torch.uint8, "ignore")
    with _Timed("metric", _nb(pred_u8, gt_u8, mask), f"rthres B{B} {H}x{W}"):
        check(lib.dfw_rthres_iou_hist(pred_u8.data_ptr(), int(is_mask), gt_u8.data_ptr(), _ptr(ignore_u8), float(r_threshold),
                                      inter.data_ptr(), union.data_ptr(), _ptr(mask), B, H, W, ws.data_ptr(), _stream()),
              "dfw_rthres_iou_hist")
    return inter, union, mask


def iou_accumulate(inter, union, class_id, inter_buf, union_buf):
    _req(inter, torch.int64, "inter"); _req(union, torch.int64, "union"); _req(class_id, torch.int64, "class_id")
    B = inter.shape[0]
    nclass = inter_buf.shape[1]
    check(lib.dfw_iou_accumulate(inter.data_ptr(), union.data_ptr(), class_id.data_ptr(), inter_buf.data_ptr(),
                                 union_buf.data_ptr(), B, nclass, _stream()), "dfw_iou_accumulate")


# ---- K9: episode preprocessing (dataset.py:36-40, coco.py:38-47) --------------------------------------------------
IMAGE_DESC_BYTES = 24      # sizeof(DfwImageDesc): int64 offset, int32 h, w, row_stride, param


def resize_normalize_u8(buf, desc_offset, n, max_h, max_w, out_h, out_w, mean=0.5, std=0.5, want_f32=True,
                        want_u8=False):
    """`buf`: uint8 device buffer holding n DfwImageDesc records at byte `desc_offset` and the decoded RGB (HWC uint8)
    images they point to.  Returns (fp32 [n,3,out_h,out_w] normalised, uint8 [n,out_h,out_w,3] resized) — the same
    bytes as torchvision's Resize -> ToTensor -> Normalize on the CPU."""
    _req(buf, torch.uint8, "buf")
    assert desc_offset % 8 == 0 and buf.data_ptr() % 8 == 0
    dev = buf.device
    dst = torch.empty((n, 3, out_h, out_w), device=dev, dtype=torch.float32) if want_f32 else None
    dst8 = torch.empty((n, out_h, out_w, 3), device=dev, dtype=torch.uint8) if want_u8 else None
    nbytes = int(lib.dfw_preproc_workspace_bytes(n, max_h, max_w, out_h, out_w))
    if nbytes < 0:
        raise ValueError("resize_normalize_u8: bad sizes")
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
    check(lib.dfw_resize_normalize_u8(buf.data_ptr(), buf.data_ptr() + desc_offset, n, max_h, max_w, _ptr(dst), _ptr(dst8),
                                      out_h, out_w, float(mean), float(std), ws.data_ptr(), nbytes, _stream()),
          "dfw_resize_normalize_u8")
    return dst, dst8


def mask_nearest(buf, desc_offset, n, out_h, out_w, mode=0, want_boundary=False):
    """Label masks (uint8 [h,w], described like images) -> fp32 {0,1} [n,out_h,out_w] at torch 'nearest' positions
    (+ the PASCAL ignore boundary floor(label / 255) when asked)."""
    _req(buf, torch.uint8, "buf")
    assert desc_offset % 8 == 0 and buf.data_ptr() % 8 == 0
    dev = buf.device
    m = torch.empty((n, out_h, out_w), device=dev, dtype=torch.float32)
    bd = torch.empty((n, out_h, out_w), device=dev, dtype=torch.float32) if want_boundary else None
    check(lib.dfw_mask_nearest(buf.data_ptr(), buf.data_ptr() + desc_offset, n, m.data_ptr(), _ptr(bd), out_h, out_w, int(mode),
                               _stream()), "dfw_mask_nearest")
    return m, bd


# ---- gradient kernels of the training step (csrc/grad.cu) ------------------------------------------------------------
_WS = {}


def _workspace(nbytes: int, device, key: str = "ws") -> torch.Tensor:
    """Grow-only scratch buffer per (device, key): stream-ordered reuse, so consecutive launches on one stream may share it."""
    k = (str(device), key)
    buf = _WS.get(k)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _WS.setdefault("retired", []).append(buf)      # a captured CUDA graph may still hold the old pointer
        buf = torch.empty(max(int(nbytes), 1 << 20), device=device, dtype=torch.uint8)
        _WS[k] = buf
    return buf


def conv_wgrad(x, dy, dw, *, ksize, stride=1, scale=1.0, accumulate=False, cout_store=None):
    """dw[co, tap*Cin + ci] (=|+=) scale * sum_pixels dy[p, co] * x[p*stride + tap - pad, ci].
    x 16-bit [N,H,W,Cin], dy 16-bit [N,H/s,W/s,Cout], dw fp32 [cout_store, ks*ks*Cin] (contiguous, written in place)."""
    assert x.dtype in OPERAND_DTYPES and dy.dtype == x.dtype
    _req(x, x.dtype, "x"); _req(dy, dy.dtype, "dy"); _req(dw, torch.float32, "dw")
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    cs = Cout if cout_store is None else int(cout_store)
    assert tuple(dy.shape) == (N, H // stride, W // stride, Cout), (x.shape, dy.shape, stride)
    assert dw.numel() == cs * ksize * ksize * Cin, (dw.shape, cs, ksize, Cin)
    ws = _workspace(int(lib.dfw_conv_wgrad_workspace_bytes(N, H, W, Cin, Cout, ksize, stride)), x.device, "wgrad")
    with _Timed("igemm", 2.0 * N * (H // stride) * (W // stride) * Cout * ksize * ksize * Cin,
                f"wgrad N{N} {H}x{W} {Cin}->{Cout} k{ksize} s{stride}"):
        check(lib.dfw_conv_wgrad(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, Cin, Cout, cs, ksize, stride,
                                 int(x.dtype == f16), float(scale), int(accumulate), ws.data_ptr(), _stream()), "dfw_conv_wgrad")
    return dw


def linear_wgrad(x, dy, dw, *, scale=1.0, accumulate=False):
    """dw[n, k] (=|+=) scale * sum_rows dy[r, n] * x[r, k]; x [.., K], dy [.., Nout] 16-bit, dw fp32 [Nout, K]."""
    K, Nout = x.shape[-1], dy.shape[-1]
    M = x.numel() // K
    return conv_wgrad(x.view(1, 1, M, K), dy.view(1, 1, M, Nout), dw, ksize=1, scale=scale, accumulate=accumulate)


def weight_permute(w, R, T, C, tap_map, out=None):
    """out[c, t', r] = w[r, tap_map[t'], c] (0 where tap_map[t'] < 0): 16-bit [R, T, C] -> [C, len(tap_map), R]."""
    assert w.dtype in OPERAND_DTYPES and w.is_cuda and w.is_contiguous() and w.numel() == R * T * C
    To = len(tap_map)
    if out is None:
        out = torch.empty((C, To * R), device=w.device, dtype=w.dtype)
    assert out.numel() == C * To * R and out.dtype == w.dtype and out.is_contiguous()
    import ctypes
    tm = (ctypes.c_int * To)(*[int(t) for t in tap_map])
    check(lib.dfw_weight_permute(w.data_ptr(), out.data_ptr(), R, T, C, To, ctypes.cast(tm, ctypes.c_void_p), _stream()),
          "dfw_weight_permute")
    return out


def colsum(x, groups=1, *, out=None, scale=1.0, accumulate=False):
    """out[g, c] (=|+=) scale * sum of the rows of group g: x [groups * rows, C] (any leading shape), fp32 out [groups, C]."""
    assert x.is_cuda and x.is_contiguous() and x.dtype in (bf16, f16, torch.float32)
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    assert rows % groups == 0
    if out is None:
        out = torch.empty((groups, Cc), device=x.device, dtype=torch.float32)
        assert not accumulate
    _req(out, torch.float32, "out")
    assert out.numel() == groups * Cc
    chunks = int(lib.dfw_colsum_chunks(rows // groups, groups))
    ws = _workspace(groups * chunks * Cc * 4, x.device, "colsum")
    with _Timed("elementwise", _nb(x), f"colsum rows{rows} C{Cc} g{groups}"):
        check(lib.dfw_colsum(x.data_ptr(), _xd(x), out.data_ptr(), rows // groups, groups, Cc, float(scale), int(accumulate),
                             ws.data_ptr(), _stream()), "dfw_colsum")
    return out


def downsum2x(dy):
    """Backward of upsample2x: dy 16-bit [N,2H,2W,C] -> dx [N,H,W,C]."""
    assert dy.dtype in OPERAND_DTYPES
    _req(dy, dy.dtype, "dy")
    N, H2, W2, Cc = dy.shape
    dx = torch.empty((N, H2 // 2, W2 // 2, Cc), device=dy.device, dtype=dy.dtype)
    with _Timed("elementwise", _nb(dy, dx), f"downsum2x N{N} {H2}x{W2} C{Cc}"):
        check(lib.dfw_downsum2x_nhwc(dy.data_ptr(), dx.data_ptr(), int(dy.dtype == f16), N, H2 // 2, W2 // 2, Cc, _stream()),
              "dfw_downsum2x_nhwc")
    return dx


def split_channels(y, Ca):
    """Backward of concat_channels: (y[..., :Ca], y[..., Ca:]) as two contiguous tensors."""
    assert y.dtype in OPERAND_DTYPES
    _req(y, y.dtype, "y")
    Cc = y.shape[-1]
    rows = y.numel() // Cc
    a = torch.empty(y.shape[:-1] + (Ca,), device=y.device, dtype=y.dtype)
    b = torch.empty(y.shape[:-1] + (Cc - Ca,), device=y.device, dtype=y.dtype)
    with _Timed("elementwise", 2 * _nb(y), f"split rows{rows} {Ca}+{Cc - Ca}"):
        check(lib.dfw_split_channels(y.data_ptr(), a.data_ptr(), b.data_ptr(), rows, Ca, Cc - Ca, _stream()), "dfw_split_channels")
    return a, b


def geglu(h):
    """y = h[..., :F] * gelu_erf(h[..., F:]) for 16-bit (value | gate) pre-activations h [.., 2F]."""
    assert h.dtype in OPERAND_DTYPES
    _req(h, h.dtype, "h")
    F = h.shape[-1] // 2
    y = torch.empty(h.shape[:-1] + (F,), device=h.device, dtype=h.dtype)
    with _Timed("elementwise", _nb(h, y), f"geglu rows{h.numel() // (2 * F)} F{F}"):
        check(lib.dfw_geglu_fwd(h.data_ptr(), y.data_ptr(), int(h.dtype == f16), h.numel() // (2 * F), F, _stream()), "dfw_geglu_fwd")
    return y


def nchw_to_nhwc_pad(x, cpad, dtype=f16, scale=1.0):
    """fp32 [N,C,H,W] -> 16-bit [N,H,W,cpad] (channels beyond C are zero), values scaled by `scale`."""
    _req(x, torch.float32, "x")
    N, Cc, H, W = x.shape
    y = torch.empty((N, H, W, cpad), device=x.device, dtype=dtype)
    check(lib.dfw_nchw_f32_to_nhwc16_pad(x.data_ptr(), y.data_ptr(), N, Cc, H, W, cpad, float(scale), int(dtype == f16), _stream()),
          "dfw_nchw_f32_to_nhwc16_pad")
    return y
