"""Prepared-weight layer objects shared by the UNet and VAE engines.

Each object owns device copies of one diffusers module's parameters in the layout the kernels want
(16-bit [Cout, taps*Cin] GEMM weights, fp32 biases / norm affine) and drives `ops.*`.  No torch math on the data path.
"""
from __future__ import annotations

from dataclasses import dataclass

import os

import torch

from . import ops
from .weights import conv_weight_to_gemm, geglu_permute, upconv_phase_weights

bf16 = torch.bfloat16


@dataclass(frozen=True)
class Precision:
    """Numerics policy (DESIGN.md "Numerics").

    half        the 16-bit format of every tensor-core operand (weights and activations) and of every 16-bit
                activation tensor: torch.float16 (default) or torch.bfloat16.  tcgen05 kind::f16 runs both at the same
                rate but needs A and B in the SAME format.  fp16 is also the reference's own half-precision mode
                (evaluation_util/main_oss.py:332-336 `--half_precision` -> torch.float16; training runs fp16 autocast,
                scripts/train_*_v3.sh).  Accumulation, softmax, statistics are always fp32.
    stream_f32  keep the residual stream / skip connections in fp32 (they only feed norm kernels and residual adds).
    mid_f32     keep conv1 -> norm2 intermediates in fp32.

    Measured, UNet latent rel-L2 vs the fp32 oracle on the 1-shot pipeline with random-init SD-2.1 weights (identical
    UNet inputs; bar 1e-2): bf16 operands 1.2e-2 .. 1.4e-2 (fails), fp16 operands ~1.3e-3.
    """
    half: torch.dtype = torch.float16
    stream_f32: bool = True
    mid_f32: bool = True
    # fp32 evaluation mode (the reference's shipped eval numerics, main_oss.py:332-336; bar 1e-4 vs the fp32 oracle): every
    # activation tensor fp32, norms / softmax / GELU / attention core in fp32 with exact transcendentals (csrc/f32mode.cu),
    # GEMMs and convolutions on the same tcgen05 kernels with split operands x = hi + lo in the `half` format
    # (hi_x hi_w + lo_x hi_w + hi_x lo_w as one GEMM over a 3x longer channel axis).  ~4x slower than the 16-bit path.
    f32: bool = False

    def __post_init__(self):
        if self.f32:
            object.__setattr__(self, "stream_f32", True)
            object.__setattr__(self, "mid_f32", True)


PURE_BF16 = Precision(half=torch.bfloat16)
F32 = Precision(f32=True)


def to_operand(h: torch.Tensor, prec: Precision) -> torch.Tensor:
    """A residual-stream tensor about to be consumed by a GEMM: 16-bit cast on the default path; left in fp32 in the fp32
    mode (the consuming Conv / Linear splits it)."""
    return h if prec.f32 else ops.cast16(h, prec.half)


def _prep_w(w2d: torch.Tensor, taps: int, device, wdtype, f32: bool, role: int = 1) -> torch.Tensor:
    """GEMM weight [rows, taps*Cin] -> device operand: a 16-bit cast, or in the fp32 mode the split [hi | hi | lo] per tap."""
    if f32:
        return ops.split3_host(w2d, taps, role, wdtype).to(device)
    return _dev(w2d, device, wdtype)


def _split_in(x: torch.Tensor, half) -> torch.Tensor:
    return ops.split3(x, 0, half) if x.dtype == torch.float32 else x

# Fold GroupNorm + SiLU into the consuming convolution (ops.conv2d_gn_in, bit-identical to norm kernel + conv: the
# normalised tensor never reaches HBM).  Measured on B200 (profiles/r02_gnin_bench.json): the fused kernel runs at ~0.8x the
# plain convolution's rate (its transform warps share the SM with the MMA pipeline), so it wins exactly where the GroupNorm
# pass it removes is large next to the convolution: +10 .. 14 % on 128->128 / 256->128 at 512^2 and 512->256 at 256^2
# without a residual (kernel timed alone), break-even at 256->256, a loss with a residual operand or at 512 channels.  In
# the whole step, which runs under the power cap, even the winning shapes give nothing back: 127.9 ms (off) / 129.2 ms
# ("auto") / 133.3 ms (True) per step at config 2.  OFF by default.
#   "auto": conv1 of a ResnetBlock when its input has >= 2^25 elements per 16 images;  True: wherever the kernel supports
#   the shape;  False (default): never.  Set before building the engines (no environment variable; bench.py --gn-fusion).
FUSE_GN_INTO_CONV = False
FUSE_GN_MIN_ELEMS_PER_IMAGE = (1 << 25) // 16


def _fuse_gn(x, has_residual: bool) -> bool:
    if FUSE_GN_INTO_CONV is True:
        return True
    if FUSE_GN_INTO_CONV == "auto":
        return (not has_residual) and x.shape[1] * x.shape[2] * x.shape[3] >= FUSE_GN_MIN_ELEMS_PER_IMAGE
    return False


def _dev(t: torch.Tensor, device, dtype) -> torch.Tensor:
    return t.detach().to(device=device, dtype=dtype).contiguous()


class Conv:
    """3x3 / 1x1 convolution on the tcgen05 implicit-GEMM kernel."""

    def __init__(self, sd, prefix, device, stride=1, pad_mode=0, wdtype=bf16, f32=False):
        w = sd[prefix + ".weight"]
        self.cout, self.cin, self.ksize, _ = w.shape
        self.stride, self.pad_mode = stride, pad_mode
        self.f32, self.half = f32, wdtype
        self.w = _prep_w(conv_weight_to_gemm(w), self.ksize * self.ksize, device, wdtype, f32)
        self.b = _dev(sd[prefix + ".bias"], device, torch.float32)

    def __call__(self, x, *, bias=None, bias_per_sample=False, residual=None, out_f32=False, out_scale=1.0,
                 gn_stats=False):
        if self.f32:
            x, out_f32, gn_stats = _split_in(x, self.half), True, False
        return ops.conv2d(x, self.w, self.b if bias is None else bias, ksize=self.ksize, stride=self.stride,
                          pad_mode=self.pad_mode, residual=residual, out_f32=out_f32, out_scale=out_scale,
                          bias_per_sample=bias_per_sample, gn_stats=gn_stats)


class UpsampleConv:
    """diffusers Upsample2D (nearest 2x + 3x3 conv) as four 2x2-tap phase convolutions (weights.upconv_phase_weights)."""

    def __init__(self, sd, prefix, device, wdtype=bf16, f32=False):
        w4 = upconv_phase_weights(sd[prefix + ".weight"])                       # [4, Cout, 4*Cin]
        self.f32, self.half = f32, wdtype
        self.w4 = _prep_w(w4.reshape(-1, w4.shape[-1]), 4, device, wdtype, f32).view(4, w4.shape[1], -1)
        self.b = _dev(sd[prefix + ".bias"], device, torch.float32)

    def __call__(self, h, out_f32, gn_stats=False):
        if self.f32:
            return ops.upconv2x(_split_in(h, self.half), self.w4, self.b, out_f32=True)
        return ops.upconv2x(ops.cast16(h, self.half), self.w4, self.b, out_f32=out_f32, gn_stats=gn_stats)


class SmallCinConv:
    """3x3 / pad 1 conv with Cin <= 16 reading the reference's NCHW fp32 tensor: im2col (K = 9*Cin padded to a multiple
    of 64) + one tcgen05 GEMM.  conv_in / conv_in_ref of the UNet, conv_in of the VAE encoder / decoder."""

    def __init__(self, sd, prefix, device, wdtype=bf16, f32=False):
        w = sd[prefix + ".weight"]
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.f32 = f32
        if f32:      # fp32 mode: the direct CUDA-core kernel (fp32 FMA) on the NCHW tensor
            self.w32 = _dev(w.detach().float().permute(0, 2, 3, 1), device, torch.float32)      # [Cout,3,3,Cin]
        k = 9 * self.cin
        self.kpad = (k + 63) // 64 * 64
        wg = torch.zeros(self.cout, self.kpad, dtype=torch.float32)
        wg[:, :k] = conv_weight_to_gemm(w.detach().float().cpu())
        self.w = _dev(wg, device, wdtype)
        self.b = _dev(sd[prefix + ".bias"], device, torch.float32)
        self.half = wdtype

    def __call__(self, x_nchw, out_f32, gn_stats=False):
        if self.f32:
            return ops.conv3x3_small_cin(x_nchw, self.w32, self.b, out_dtype=torch.float32)
        cols = ops.im2col3x3_small(x_nchw, self.kpad, self.half)      # [N,H,W,Kpad]: a 1x1 convolution from here on
        return ops.conv2d(cols, self.w, self.b, ksize=1, out_f32=out_f32, gn_stats=gn_stats)


class Linear:
    def __init__(self, sd, prefix, device, geglu=False, wdtype=bf16, f32=False):
        w = sd[prefix + ".weight"]
        b = sd.get(prefix + ".bias")
        self.f32, self.half = f32, wdtype
        self.geglu = geglu and not f32       # fp32 mode: plain projection, the exact-erf GEGLU is its own fp32 kernel
        self.geglu_f32 = geglu and f32
        if self.geglu:
            w, b = geglu_permute(w, b)
        self.w = _prep_w(w, 1, device, wdtype, f32)
        self.b = _dev(b, device, torch.float32) if b is not None else None

    def __call__(self, x, *, residual=None, out_f32=False):
        if self.f32:
            y = ops.linear(_split_in(x, self.half), self.w, self.b, residual=residual, out_f32=True)
            return ops.geglu_f32(y) if self.geglu_f32 else y
        return ops.linear(x, self.w, self.b, residual=residual, out_f32=out_f32, geglu=self.geglu)


class FusedLinear:
    """Several bias-free Linears on the same input, one GEMM (rows of the weights concatenated)."""

    def __init__(self, sd, prefixes, device, wdtype=bf16, f32=False):
        ws = [sd[p + ".weight"] for p in prefixes]
        self.splits = [w.shape[0] for w in ws]
        self.f32, self.half = f32, wdtype
        self.w = _prep_w(torch.cat(ws, 0), 1, device, wdtype, f32)
        bs = [sd.get(p + ".bias") for p in prefixes]
        self.b = _dev(torch.cat(bs, 0), device, torch.float32) if bs[0] is not None else None

    def __call__(self, x):
        if self.f32:
            return ops.linear(_split_in(x, self.half), self.w, self.b, out_f32=True)
        return ops.linear(x, self.w, self.b)


class GroupNorm:
    def __init__(self, sd, prefix, device, eps, groups=32, out_dtype=bf16, f32=False):
        self.g = _dev(sd[prefix + ".weight"], device, torch.float32)
        self.b = _dev(sd[prefix + ".bias"], device, torch.float32)
        self.eps, self.groups, self.out_dtype, self.f32 = eps, groups, out_dtype, f32

    def __call__(self, x, silu):
        if self.f32:
            return ops.groupnorm_f32(x, self.g, self.b, groups=self.groups, eps=self.eps, silu=silu)
        return ops.groupnorm(x, self.g, self.b, groups=self.groups, eps=self.eps, silu=silu, out_dtype=self.out_dtype)


class LayerNorm:
    def __init__(self, sd, prefix, device, eps=1e-5, out_dtype=bf16, f32=False):
        self.g = _dev(sd[prefix + ".weight"], device, torch.float32)
        self.b = _dev(sd[prefix + ".bias"], device, torch.float32)
        self.eps, self.out_dtype, self.f32 = eps, out_dtype, f32

    def __call__(self, x):
        if self.f32:
            return ops.layernorm_f32(x, self.g, self.b, self.eps)
        return ops.layernorm(x, self.g, self.b, self.eps, out_dtype=self.out_dtype)


class Resnet:
    """diffusers ResnetBlock2D: GN-SiLU-conv3x3 (+temb) - GN-SiLU-conv3x3 (+1x1 shortcut) + x."""

    def __init__(self, sd, prefix, device, eps, prec: Precision, has_temb: bool):
        self.prec = prec
        wd = nd = prec.half
        f = prec.f32
        self.norm1 = GroupNorm(sd, prefix + ".norm1", device, eps, out_dtype=nd, f32=f)
        self.conv1 = Conv(sd, prefix + ".conv1", device, wdtype=wd, f32=f)
        self.norm2 = GroupNorm(sd, prefix + ".norm2", device, eps, out_dtype=nd, f32=f)
        self.conv2 = Conv(sd, prefix + ".conv2", device, wdtype=wd, f32=f)
        self.shortcut = (Conv(sd, prefix + ".conv_shortcut", device, wdtype=wd, f32=f)
                         if (prefix + ".conv_shortcut.weight") in sd else None)
        # time-embedding projection is folded into conv1's bias per timestep (fp32, host): see UNet._temb_biases
        self.temb_w = sd[prefix + ".time_emb_proj.weight"].detach().float().cpu() if has_temb else None
        self.temb_b = sd[prefix + ".time_emb_proj.bias"].detach().float().cpu() if has_temb else None
        self.conv1_bias_host = sd[prefix + ".conv1.bias"].detach().float().cpu()

    def __call__(self, h, conv1_bias=None, gn_stats_out=False):
        """`gn_stats_out`: the block's output feeds another GroupNorm — let conv2's epilogue emit its statistics.
        conv1's output always feeds norm2, so conv1 always tries to (ops.conv2d falls back silently when the shape
        is not supported, e.g. the UNet's 10/20/40-channel groups)."""
        p = self.prec
        # GroupNorm + SiLU folded into the conv operand (no normalised tensor in HBM) where the kernel supports the
        # shape and the statistics of the input already exist; otherwise norm kernel + conv
        if _fuse_gn(h, False) and conv1_bias is None and not p.mid_f32 and self.norm1.groups == 32 \
                and ops.conv_gn_in_supported(h, self.conv1.cout, self.conv1.ksize):
            t = ops.conv2d_gn_in(h, self.norm1.g, self.norm1.b, self.norm1.eps, self.conv1.w, self.conv1.b,
                                 ksize=self.conv1.ksize, gn_stats=True)
        else:
            a = self.norm1(h, silu=True)
            t = self.conv1(a, bias=conv1_bias, out_f32=p.mid_f32, gn_stats=True)
        s = h if self.shortcut is None else self.shortcut(to_operand(h, p), out_f32=p.stream_f32)
        if _fuse_gn(t, True) and not p.stream_f32 and self.norm2.groups == 32 \
                and ops.conv_gn_in_supported(t, self.conv2.cout, self.conv2.ksize):
            return ops.conv2d_gn_in(t, self.norm2.g, self.norm2.b, self.norm2.eps, self.conv2.w, self.conv2.b,
                                    ksize=self.conv2.ksize, residual=s, gn_stats=gn_stats_out)
        c = self.norm2(t, silu=True)
        return self.conv2(c, residual=s, out_f32=p.stream_f32, gn_stats=gn_stats_out)
