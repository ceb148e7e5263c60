"""AutoencoderKL (SD-2.1 VAE) encoder / decoder on the B200 kernels.

Reference call sites: diffews/marigold_pipeline_rgb_latent_noise.py:839-862 (encode_rgb: vae.encoder -> quant_conv ->
mean * 0.18215) and :887-905 (decode_seg: z / 0.18215 -> post_quant_conv -> vae.decoder -> clip(-1,1)).
Layer schedule = diffusers-0.25 Encoder / Decoder / UNetMidBlock2D; state-dict keys unchanged.

The mid-block attention is single-head with d = 512 over (H/8 * W/8) tokens; it is only 2 x 34 GFLOP per image, so it
runs as three tcgen05 GEMMs (Q K^T, fp32 logits; row softmax; P V with V^T produced directly by a transposed
projection GEMM) instead of a dedicated d=512 flash kernel.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from .layers import Conv, GroupNorm, Linear, Precision, Resnet, SmallCinConv, UpsampleConv, _dev, _prep_w, to_operand

bf16 = torch.bfloat16


class VaeAttention:
    """diffusers Attention(heads=1, dim_head=512, bias=True, residual_connection=True, GroupNorm(32, eps 1e-6))."""

    def __init__(self, sd, prefix, device, prec: Precision):
        self.prec = prec
        wd = prec.half
        f = prec.f32
        self.norm = GroupNorm(sd, prefix + ".group_norm", device, eps=1e-6, out_dtype=prec.half, f32=f)
        self.to_q = Linear(sd, prefix + ".to_q", device, wdtype=wd, f32=f)
        self.to_k = Linear(sd, prefix + ".to_k", device, wdtype=wd, f32=f)
        # V^T = W_v X^T is produced directly ([C, L] per image, the K-major B operand of P V); its bias is added
        # after the P V product instead (softmax rows sum to 1, so P (V + 1 b^T) = P V + b^T).
        self.wv = _prep_w(sd[prefix + ".to_v.weight"], 1, device, wd, f, role=0)    # the A operand of V^T = W_v X^T
        self.bv = _dev(sd[prefix + ".to_v.bias"], device, torch.float32)
        self.to_out = Linear(sd, prefix + ".to_out.0", device, wdtype=wd, f32=f)
        self.C = self.wv.shape[0]
        self.scale = self.C ** -0.5

    def __call__(self, h):
        N, H, W, C = h.shape
        L = H * W
        xn = self.norm(h, silu=False).view(N, L, C)
        if self.prec.f32:
            return self._call_f32(h, xn)
        q = self.to_q(xn)
        k = self.to_k(xn)
        vt = ops.linear(self.wv, xn.view(N * L, C))                       # [C, N*L]: V^T of every image, side by side
        vt = vt.view(C, N, L).permute(1, 0, 2)                            # [N, C, L] view (row stride N*L)
        s = ops.bmm_nt(q, k, out_f32=True)                                # [N, L, L] fp32 logits, one launch
        p = ops.softmax_rows(s, self.scale, out_dtype=self.prec.half)
        o = ops.bmm_nt(p, vt, self.bv)                                    # P V + b_v, one launch
        y = self.to_out(o, residual=h.view(N, L, C), out_f32=self.prec.stream_f32)
        return y.view(N, H, W, C)


    def _call_f32(self, h, xn):
        """fp32 mode: the same three GEMMs on split operands (activation x activation products pair a role-0 with a role-1
        split), fp32 logits normalised in place by the exact-exp softmax."""
        N, H, W, C = h.shape
        L, half = H * W, self.prec.half
        xa = ops.split3(xn, 0, half)                                              # [N, L, 3C]
        q, k = self.to_q(xa), self.to_k(xa)                                       # fp32 [N, L, C]
        vt = ops.linear(self.wv, ops.split3(xn.view(N * L, C), 1, half), out_f32=True)       # fp32 [C, N*L]
        s = ops.bmm_nt(ops.split3(q, 0, half), ops.split3(k, 1, half), out_f32=True)         # [N, L, L]
        p = ops.softmax_rows_f32(s, self.scale)
        vts = ops.split3(vt.view(C * N, L), 1, half).view(C, N, 3 * L).permute(1, 0, 2)      # [N, C, 3L] view
        o = ops.bmm_nt(ops.split3(p, 0, half), vts, self.bv, out_f32=True)                   # P V + b_v
        return self.to_out(o, residual=h.view(N, L, C), out_f32=True).view(N, H, W, C)


class _Mid:
    def __init__(self, sd, prefix, device, prec):
        self.res0 = Resnet(sd, prefix + ".resnets.0", device, 1e-6, prec, has_temb=False)
        self.attn = VaeAttention(sd, prefix + ".attentions.0", device, prec)
        self.res1 = Resnet(sd, prefix + ".resnets.1", device, 1e-6, prec, has_temb=False)

    def __call__(self, h, gn_stats_out=True):
        # res0's output feeds the attention's GroupNorm through a .view (statistics attribute is dropped there)
        return self.res1(self.attn(self.res0(h, gn_stats_out=True)), gn_stats_out=gn_stats_out)


class AutoencoderKL:
    def __init__(self, state_dict, device="cuda", block_out_channels=(128, 256, 512, 512),
                 precision: Optional[Precision] = None):
        sd = state_dict
        self.device = dev = torch.device(device)
        # default: every VAE activation in the 16-bit format (the decoder's 512^2 x 128-channel tensors make the VAE
        # bandwidth-sensitive; with fp16 operands the decoded mask still agrees with the fp32 oracle on > 99.9 %)
        self.prec = prec = precision or Precision(stream_f32=False, mid_f32=False)
        self._sdt = torch.float32 if prec.stream_f32 else prec.half
        self.dtype = torch.float32
        c = tuple(block_out_channels)
        self.config = SimpleNamespace(block_out_channels=c, latent_channels=4, scaling_factor=0.18215)
        # ---- encoder
        f = prec.f32
        self.enc_conv_in = SmallCinConv(sd, "encoder.conv_in", dev, prec.half, f32=f)
        self.enc_down = []
        for i in range(4):
            blk = SimpleNamespace(
                resnets=[Resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", dev, 1e-6, prec, False) for j in range(2)],
                down=Conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv", dev, stride=2, pad_mode=1,
                          wdtype=prec.half, f32=f) if i < 3 else None)
            self.enc_down.append(blk)
        self.enc_mid = _Mid(sd, "encoder.mid_block", dev, prec)
        self.enc_norm_out = GroupNorm(sd, "encoder.conv_norm_out", dev, eps=1e-6, out_dtype=prec.half, f32=f)
        self.enc_conv_out = Conv(sd, "encoder.conv_out", dev, wdtype=prec.half, f32=f)
        # quant_conv (1x1, 8->8): only the 4 `mean` channels are consumed (pipeline:858-860); host-side constants
        self.quant_w = sd["quant_conv.weight"].detach().float().cpu()[:4, :, 0, 0].contiguous()
        self.quant_b = sd["quant_conv.bias"].detach().float().cpu()[:4].contiguous()
        # ---- decoder
        self.post_quant_w = sd["post_quant_conv.weight"].detach().float().cpu()[:, :, 0, 0].contiguous()
        self.post_quant_b = sd["post_quant_conv.bias"].detach().float().cpu().contiguous()
        self.dec_conv_in = SmallCinConv(sd, "decoder.conv_in", dev, prec.half, f32=f)
        self.dec_mid = _Mid(sd, "decoder.mid_block", dev, prec)
        self.dec_up = []
        for i in range(4):
            blk = SimpleNamespace(
                resnets=[Resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", dev, 1e-6, prec, False) for j in range(3)],
                up=UpsampleConv(sd, f"decoder.up_blocks.{i}.upsamplers.0.conv", dev, wdtype=prec.half, f32=f)
                if i < 3 else None)
            self.dec_up.append(blk)
        self.dec_norm_out = GroupNorm(sd, "decoder.conv_norm_out", dev, eps=1e-6, out_dtype=prec.half, f32=f)
        self.dec_conv_out = Conv(sd, "decoder.conv_out", dev, wdtype=prec.half, f32=f)
        # fused decoder head (dfw_seg_head_u8): conv_out as mma.sync B fragments, bias as a kernel parameter
        wout = sd["decoder.conv_out.weight"]
        self.seg_head_wb = ops.seg_head_prepare(wout, prec.half, dev) if tuple(wout.shape) == (3, 128, 3, 3) else None
        self.seg_head_bias = sd["decoder.conv_out.bias"].detach().float().cpu().contiguous()

    @classmethod
    def from_module(cls, module: torch.nn.Module, device="cuda", **kw):
        cfg = {}
        if hasattr(module, "encoder"):
            cfg["block_out_channels"] = tuple(b.resnets[0].conv1.out_channels for b in module.encoder.down_blocks)
        cfg.update(kw)
        return cls(module.state_dict(), device=device, **cfg)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, device="cuda", precision=None, **unused):
        """`AutoencoderKL.from_pretrained(ckpt, subfolder="vae")` (main_oss.py:347-349), local diffusers directory."""
        from . import checkpoint
        return checkpoint.load_vae(pretrained_model_name_or_path, device=device, precision=precision, subfolder=subfolder)

    def to(self, *a, **k):
        return self

    # ---- encoder: [N,3,H,W] fp32 in [-1,1] -> latent mean * scale  [N,4,H/8,W/8] fp32 (NCHW, reference layout) -----
    @torch.no_grad()
    def encode_mean(self, x_nchw: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        if not x_nchw.is_cuda:
            raise RuntimeError("AutoencoderKL (B200 engine) needs CUDA tensors: there is no CPU fallback")
        x = x_nchw.to(torch.float32).contiguous()
        N, _, H, W = x.shape
        if H % 8 or W % 8:
            raise ValueError("image height/width must be multiples of 8")
        f32 = self.prec.stream_f32
        # every tensor below is consumed by a GroupNorm next: the producing conv emits its statistics (gn_stats)
        h = self.enc_conv_in(x, out_f32=f32, gn_stats=True)
        for blk in self.enc_down:
            for j, r in enumerate(blk.resnets):
                h = r(h, gn_stats_out=not (blk.down is not None and j == len(blk.resnets) - 1))
            if blk.down is not None:
                h = blk.down(to_operand(h, self.prec), out_f32=f32, gn_stats=True)
        h = self.enc_mid(h)
        h = self.enc_norm_out(h, silu=True)
        m = self.enc_conv_out(h, out_f32=True)                                   # [N,h,w,8] fp32 moments
        hh, ww = H // 8, W // 8
        lat = torch.empty((N, 4, hh, ww), device=x.device, dtype=torch.float32)
        ops.pointwise_small(m, (hh * ww * 8, 8, 1), self.quant_w, self.quant_b, lat, (4 * hh * ww, 1, hh * ww),
                            N, hh * ww, in_scale=1.0, out_scale=scale)
        return lat

    # ---- decoder: z [N,4,h,w] fp32 (already divided by the scale factor via in_scale) -> fp32 [N, H*W, 3] rows -----
    @torch.no_grad()
    def decode_seg(self, z_nchw: torch.Tensor, in_scale: float = 1.0, want_f32: bool = True, want_u8: bool = True):
        """decoder -> clip(-1,1) -> *0.5+0.5 -> *255 (-> uint8 truncation): (seg_f32 [N,3,H,W] | None, seg_u8 | None).
        pipeline:787-795, :887-905, :534.  The head (GroupNorm + SiLU + conv_out + the whole tail) is ONE kernel when the
        decoder has its real width (128 channels before conv_out); otherwise norm kernel + conv + dfw_seg_post."""
        h = self._decode_trunk(z_nchw, in_scale)
        N, H, W, _ = h.shape
        if self.seg_head_wb is not None and ops.seg_head_supported(h):
            n = self.dec_norm_out
            return ops.seg_head_u8(h, n.g, n.b, n.eps, self.seg_head_wb, self.seg_head_bias, want_f32=want_f32, want_u8=want_u8)
        y = self.dec_conv_out(self.dec_norm_out(h, silu=True), out_f32=True)
        return ops.seg_post(y.view(N, H * W, 3), H, W, want_f32=want_f32, want_u8=want_u8)

    @torch.no_grad()
    def decode_rows(self, z_nchw: torch.Tensor, in_scale: float = 1.0) -> torch.Tensor:
        h = self._decode_trunk(z_nchw, in_scale)
        h = self.dec_norm_out(h, silu=True)
        y = self.dec_conv_out(h, out_f32=True)                                   # [N,H,W,3] fp32
        return y.view(y.shape[0], y.shape[1] * y.shape[2], 3)

    @torch.no_grad()
    def _decode_trunk(self, z_nchw: torch.Tensor, in_scale: float = 1.0) -> torch.Tensor:
        if not z_nchw.is_cuda:
            raise RuntimeError("AutoencoderKL (B200 engine) needs CUDA tensors: there is no CPU fallback")
        z = z_nchw.to(torch.float32).contiguous()
        N, Cz, hh, ww = z.shape
        f32 = self.prec.stream_f32
        zq = torch.empty_like(z)
        ops.pointwise_small(z, (Cz * hh * ww, 1, hh * ww), self.post_quant_w, self.post_quant_b, zq,
                            (Cz * hh * ww, 1, hh * ww), N, hh * ww, in_scale=in_scale, out_scale=1.0)
        h = self.dec_conv_in(zq, out_f32=f32, gn_stats=True)
        h = self.dec_mid(h)
        for blk in self.dec_up:
            for j, r in enumerate(blk.resnets):
                h = r(h, gn_stats_out=not (blk.up is not None and j == len(blk.resnets) - 1))
            if blk.up is not None:
                h = blk.up(h, out_f32=f32, gn_stats=True)
        return h                                                                 # [N,H,W,C_last], feeds conv_norm_out

    @torch.no_grad()
    def decode(self, z_nchw: torch.Tensor) -> torch.Tensor:
        """diffusers-style `vae.decoder(post_quant_conv(z))`: returns NCHW fp32 [N,3,H,W]."""
        N, _, hh, ww = z_nchw.shape
        rows = self.decode_rows(z_nchw)
        return ops.nhwc_f32_to_nchw(rows, 3, hh * 8, ww * 8)
