"""ORACLE (test infrastructure) — plain-PyTorch fp32 restatement of the SD-2.1 UNet (+DiffewS deltas) and VAE.

The arithmetic of these modules lives in the third-party dependency `diffusers==0.25.0` (reference requirements.txt:2),
which is NOT vendored in /root/reference and not installed here; it is restated from its published architecture with
diffusers' state-dict key names so real checkpoints can be dropped in.  Anchors in the reference:

  diffews/models/unet_2d_condition.py:88-664    MyUNet2DConditionModel.__init__ (copy of UNet2DConditionModel):
        :299-306 conv_in / conv_in_ref (in_channels_ref=8) ; :643-664 apply_unet_refonly_block / clear_attn_bank
  diffews/models/unet_2d_condition.py:879-1258  forward: :1008-1015 time embedding, :1118-1121 is_target switch,
        :1154-1175 down blocks, :1189-1200 mid, :1214-1243 up blocks, :1246-1249 conv_norm_out/act/conv_out
  diffews/models/attention_processor.py:39-58, 182-288 (xformers, k-shot fold), 291-383 (SDPA) — KV bank protocol
  train_tools/load_ckpt_and_modify_ref8in_tag4in.py:6-28 — base model stabilityai/stable-diffusion-2-1 and
        conv_in_ref.weight = conv_in.weight.repeat(1,2,1,1)/2
  diffews/marigold_pipeline_rgb_latent_noise.py:852-853, 898-902 — vae.encoder / quant_conv / post_quant_conv / decoder

SD-2.1 UNet config: block_out_channels (320,640,1280,1280), layers_per_block 2, attention heads (5,10,20,20)
(head_dim 64), cross_attention_dim 1024, use_linear_projection, norm_num_groups 32, norm_eps 1e-5, act silu.
VAE config: block_out_channels (128,256,512,512), layers_per_block 2, latent_channels 4, norm groups 32, eps 1e-6.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------------
# embeddings (diffusers models/embeddings.py: get_timestep_embedding, TimestepEmbedding)
# ---------------------------------------------------------------------------------------------------------------------
def timestep_embedding(timesteps: torch.Tensor, dim: int = 320, flip_sin_to_cos: bool = True,
                       downscale_freq_shift: float = 0.0, max_period: int = 10000) -> torch.Tensor:
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# ---------------------------------------------------------------------------------------------------------------------
# attention with the DiffewS KV bank
# ---------------------------------------------------------------------------------------------------------------------
class Attention(nn.Module):
    """diffusers Attention (+ MyAttention bank state, attention_processor.py:39-58).

    `bank=True` marks a self-attention re-classed to MyAttention (unet_2d_condition.py:645-654).  Protocol
    (attention_processor.py:251-267): first call after clear stores K,V; later calls concatenate
    [K_self, fold(K_bank)] where fold is the xformers processor's k-shot fold == shot-major concatenation.
    """

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 bias: bool = False, out_bias: bool = True, norm_num_groups: Optional[int] = None, eps: float = 1e-5,
                 residual_connection: bool = False, bank: bool = False):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.residual_connection = residual_connection
        self.group_norm = nn.GroupNorm(norm_num_groups, query_dim, eps=eps, affine=True) if norm_num_groups else None
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.has_bank = bank
        self.k_bank = None
        self.v_bank = None

    def clear_bank(self):  # attention_processor.py:46-50
        self.k_bank = None
        self.v_bank = None

    def _split(self, t):  # [B, L, C] -> [B, h, L, d]
        B, L, C = t.shape
        return t.view(B, L, self.heads, C // self.heads).transpose(1, 2)

    def forward(self, hidden_states, encoder_hidden_states=None):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, hh, ww = hidden_states.shape
            hidden_states = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        q = self.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        if self.has_bank:
            if self.k_bank is None:                       # support pass: store (attention_processor.py:251-252)
                self.k_bank, self.v_bank = k, v
            else:                                          # query pass: fold + concat (:253-267)
                B = q.shape[0]
                k = torch.cat([k, self.k_bank.reshape(B, -1, k.shape[-1])], dim=1)
                v = torch.cat([v, self.v_bank.reshape(B, -1, v.shape[-1])], dim=1)
        qh, kh, vh = self._split(q), self._split(k), self._split(v)
        # memory_efficient_attention / SDPA == softmax(q k^T * scale) v   (attention_processor.py:269-271, :363-365)
        out = F.scaled_dot_product_attention(qh, kh, vh, scale=self.scale)
        out = out.transpose(1, 2).reshape(q.shape[0], -1, q.shape[-1])
        out = self.to_out[0](out)
        out = self.to_out[1](out)
        if input_ndim == 4:
            out = out.transpose(-1, -2).reshape(b, c, hh, ww)
        if self.residual_connection:
            out = out + residual
        return out


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)          # exact (erf) GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, dim_head, bank=True)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h, ehs):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), ehs) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: int, groups: int = 32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, ehs):
        b, c, hh, ww = x.shape
        residual = x
        h = self.norm(x)
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, ehs)
        h = self.proj_out(h)
        h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
        return h + residual


# ---------------------------------------------------------------------------------------------------------------------
# resnet / sampling blocks
# ---------------------------------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int] = 1280, groups: int = 32,
                 eps: float = 1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, cin, cout, heads, cross_dim, add_downsample=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x, temb, ehs):
        outs = ()
        for r, a in zip(self.resnets, self.attentions):
            x = a(r(x, temb), ehs)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class DownBlock2D(nn.Module):
    def __init__(self, cin, cout, add_downsample=False):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout) for i in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_downsample else None

    def forward(self, x, temb, ehs=None):
        outs = ()
        for r in self.resnets:
            x = r(x, temb)
            outs += (x,)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs += (x,)
        return x, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, c, heads, cross_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c), ResnetBlock2D(c, c)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, cross_dim)])

    def forward(self, x, temb, ehs):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D (heads=None) / CrossAttnUpBlock2D."""

    def __init__(self, cin, cout, prev, heads, cross_dim, add_upsample):
        super().__init__()
        res = []
        for i in range(3):
            skip = cin if i == 2 else cout
            rin = prev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout))
        self.resnets = nn.ModuleList(res)
        self.attentions = (nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(3)])
                           if heads else None)
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, res_tuple, temb, ehs):
        for i, r in enumerate(self.resnets):
            skip = res_tuple[-1]
            res_tuple = res_tuple[:-1]
            x = r(torch.cat([x, skip], dim=1), temb)
            if self.attentions is not None:
                x = self.attentions[i](x, ehs)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet(nn.Module):
    """MyUNet2DConditionModel with the SD-2.1 configuration (unet_2d_condition.py:88, :879)."""

    def __init__(self, block_out_channels=(320, 640, 1280, 1280), heads=(5, 10, 20, 20), cross_attention_dim=1024,
                 in_channels=4, in_channels_ref=8, out_channels=4):
        super().__init__()
        c = block_out_channels
        self.block_out_channels = tuple(c)
        self.heads = tuple(heads)
        self.conv_in = nn.Conv2d(in_channels, c[0], 3, padding=1)
        self.conv_in_ref = nn.Conv2d(in_channels_ref, c[0], 3, padding=1)      # unet_2d_condition.py:304-306
        self.time_embedding = TimestepEmbedding(c[0], c[0] * 4)
        downs = []
        out = c[0]
        for i in range(4):
            cin, out = out, c[i]
            if i < 3:
                downs.append(CrossAttnDownBlock2D(cin, out, heads[i], cross_attention_dim, add_downsample=True))
            else:
                downs.append(DownBlock2D(cin, out, add_downsample=False))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = UNetMidBlock2DCrossAttn(c[-1], heads[-1], cross_attention_dim)
        rc = list(reversed(c))
        rh = list(reversed(heads))
        ups = []
        out = rc[0]
        for i in range(4):
            prev, out = out, rc[i]
            cin = rc[min(i + 1, 3)]
            ups.append(UpBlock(cin, out, prev, None if i == 0 else rh[i], cross_attention_dim, add_upsample=i < 3))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(32, c[0], eps=1e-5)
        self.conv_out = nn.Conv2d(c[0], out_channels, 3, padding=1)
        if c[0] * 4 != 1280:  # reduced-width test configs: rebuild time_emb_proj with the right temb width
            for m in self.modules():
                if isinstance(m, ResnetBlock2D) and m.time_emb_proj is not None:
                    m.time_emb_proj = nn.Linear(c[0] * 4, m.conv1.out_channels)

    def bank_attentions(self):
        return [m.attn1 for m in self.modules() if isinstance(m, BasicTransformerBlock)]

    def clear_attn_bank(self):                                                  # unet_2d_condition.py:656-664
        for a in self.bank_attentions():
            a.clear_bank()

    def forward(self, sample, timestep, encoder_hidden_states, is_target: bool = True):
        t = torch.as_tensor(timestep, device=sample.device)
        if t.ndim == 0:
            t = t[None]
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(timestep_embedding(t, self.block_out_channels[0]).to(sample.dtype))
        sample = self.conv_in(sample) if is_target else self.conv_in_ref(sample)   # :1118-1121
        res = (sample,)
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, encoder_hidden_states)
            res += outs
        sample = self.mid_block(sample, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            sample = blk(sample, res[-n:], emb, encoder_hidden_states)
            res = res[:-n]
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))                 # :1246-1249
        return sample


# ---------------------------------------------------------------------------------------------------------------------
# VAE (diffusers AutoencoderKL: Encoder / Decoder / UNetMidBlock2D)
# ---------------------------------------------------------------------------------------------------------------------
class VAEMidBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, None, eps=1e-6), ResnetBlock2D(c, c, None, eps=1e-6)])
        self.attentions = nn.ModuleList([Attention(c, None, heads=1, dim_head=c, bias=True, norm_num_groups=32,
                                                   eps=1e-6, residual_connection=True)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=1e-6) for i in range(2)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UpDecoderBlock2D(nn.Module):
    def __init__(self, cin, cout, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=1e-6) for i in range(3)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Encoder(nn.Module):
    def __init__(self, c=(128, 256, 512, 512), in_channels=3, latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, c[0], 3, padding=1)
        blocks, out = [], c[0]
        for i in range(4):
            cin, out = out, c[i]
            blocks.append(DownEncoderBlock2D(cin, out, add_downsample=i < 3))
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = VAEMidBlock(c[-1])
        self.conv_norm_out = nn.GroupNorm(32, c[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(c[-1], 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, c=(128, 256, 512, 512), out_channels=3, latent=4):
        super().__init__()
        rc = list(reversed(c))
        self.conv_in = nn.Conv2d(latent, rc[0], 3, padding=1)
        self.mid_block = VAEMidBlock(rc[0])
        blocks, out = [], rc[0]
        for i in range(4):
            prev, out = out, rc[i]
            blocks.append(UpDecoderBlock2D(prev, out, add_upsample=i < 3))
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, c[0], eps=1e-6)
        self.conv_out = nn.Conv2d(c[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.conv_in(z)
        x = self.mid_block(x)
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    def __init__(self, c=(128, 256, 512, 512), latent=4):
        super().__init__()
        self.encoder = Encoder(c, 3, latent)
        self.decoder = Decoder(c, 3, latent)
        self.quant_conv = nn.Conv2d(2 * latent, 2 * latent, 1)
        self.post_quant_conv = nn.Conv2d(latent, latent, 1)


# ---------------------------------------------------------------------------------------------------------------------
# deterministic random-init models (default PyTorch init under manual_seed(0); conv_in_ref from conv_in)
# ---------------------------------------------------------------------------------------------------------------------
def build_models(seed: int = 0, unet_channels=(320, 640, 1280, 1280), unet_heads=(5, 10, 20, 20),
                 vae_channels=(128, 256, 512, 512), cross_attention_dim: int = 1024):
    torch.manual_seed(seed)
    unet = UNet(unet_channels, unet_heads, cross_attention_dim)
    with torch.no_grad():   # train_tools/load_ckpt_and_modify_ref8in_tag4in.py:6-28
        unet.conv_in_ref.weight.copy_(unet.conv_in.weight.repeat(1, 2, 1, 1) / 2)
        unet.conv_in_ref.bias.copy_(unet.conv_in.bias)
    vae = AutoencoderKL(vae_channels)
    unet.eval().requires_grad_(False)
    vae.eval().requires_grad_(False)
    return unet, vae
