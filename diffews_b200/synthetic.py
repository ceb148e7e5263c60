"""Deterministic synthetic episodes with the reference's batch layout (evaluation_util/data/coco.py:49-60):
query_img [3,S,S], query_mask [S,S], support_imgs [k,3,S,S], support_masks [k,S,S], class_id.  SURVEY §8(d)."""
from __future__ import annotations

import torch


def _ellipse_mask(g: torch.Generator, size: int) -> torch.Tensor:
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32),
                            indexing="ij")
    mask = torch.zeros(size, size, dtype=torch.bool)
    n = int(torch.randint(1, 4, (1,), generator=g))
    for _ in range(n):
        cy, cx = (torch.rand(2, generator=g) * 0.6 + 0.2) * size
        ry, rx = (torch.rand(2, generator=g) * 0.25 + 0.08) * size
        mask |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return mask.to(torch.float32)


def make_episode(idx: int, size: int = 512, nshot: int = 1, nclass: int = 80) -> dict:
    g = torch.Generator().manual_seed(1234 + idx)
    ep = {
        "query_img": torch.rand(3, size, size, generator=g) * 2 - 1,
        "support_imgs": torch.rand(nshot, 3, size, size, generator=g) * 2 - 1,
        "query_mask": _ellipse_mask(g, size),
        "support_masks": torch.stack([_ellipse_mask(g, size) for _ in range(nshot)]),
        "class_id": torch.randint(0, nclass, (1,), generator=g)[0],
    }
    return ep


def make_batch(start: int, B: int, size: int = 512, nshot: int = 1, nclass: int = 80) -> dict:
    """Collated like a DataLoader batch: query_img [B,3,S,S], query_mask [B,S,S], support_imgs [B,k,3,S,S],
    support_masks [B,k,S,S], class_id [B]."""
    eps = [make_episode(start + i, size, nshot, nclass) for i in range(B)]
    return {k: torch.stack([e[k] for e in eps]) for k in eps[0]}


def pipeline_inputs(batch: dict):
    """evaluation_util/main_oss.py:99-110: masks -> 3 channels in [-1,1]; shots folded into the batch dim."""
    sm = batch["support_masks"].unsqueeze(2).repeat(1, 1, 3, 1, 1) * 2 - 1
    si = batch["support_imgs"]
    si = si.reshape(-1, *si.shape[-3:])
    sm = sm.reshape(-1, *sm.shape[-3:])
    return [si, batch["query_img"], sm]


def prompt_embedding(lctx: int = 2, dim: int = 1024) -> torch.Tensor:
    return torch.randn(1, lctx, dim, generator=torch.Generator().manual_seed(7))


# ---------------------------------------------------------------------------------------------------------------------
# Random-init SD-2.1 weights with the diffusers state-dict key names (no checkpoints offline).  Pure shape
# bookkeeping of the published architecture -- the same keys diffews_b200.unet / .vae consume -- so the product arm of
# bench.py builds its engines without importing anything from oracle/.  tests/test_host_logic.py checks keys and shapes
# against the oracle's modules.
# ---------------------------------------------------------------------------------------------------------------------
def _uniform(g, shape, fan_in):
    b = fan_in ** -0.5                                  # torch's default Linear / Conv init: U(-1/sqrt(fan_in), +)
    return (torch.rand(shape, generator=g) * 2 - 1) * b


def _conv(sd, g, name, cout, cin, k):
    sd[name + ".weight"] = _uniform(g, (cout, cin, k, k), cin * k * k)
    sd[name + ".bias"] = _uniform(g, (cout,), cin * k * k)


def _linear(sd, g, name, nout, nin, bias=True):
    sd[name + ".weight"] = _uniform(g, (nout, nin), nin)
    if bias:
        sd[name + ".bias"] = _uniform(g, (nout,), nin)


def _norm(sd, name, c):
    sd[name + ".weight"] = torch.ones(c)
    sd[name + ".bias"] = torch.zeros(c)


def _resnet(sd, g, name, cin, cout, temb=None):
    _norm(sd, name + ".norm1", cin); _conv(sd, g, name + ".conv1", cout, cin, 3)
    if temb:
        _linear(sd, g, name + ".time_emb_proj", cout, temb)
    _norm(sd, name + ".norm2", cout); _conv(sd, g, name + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(sd, g, name + ".conv_shortcut", cout, cin, 1)


def _transformer(sd, g, name, c, xdim):
    _norm(sd, name + ".norm", c); _linear(sd, g, name + ".proj_in", c, c)
    b = name + ".transformer_blocks.0"
    for i in (1, 2, 3):
        _norm(sd, f"{b}.norm{i}", c)
    for a, kv in (("attn1", c), ("attn2", xdim)):
        _linear(sd, g, f"{b}.{a}.to_q", c, c, bias=False); _linear(sd, g, f"{b}.{a}.to_k", c, kv, bias=False)
        _linear(sd, g, f"{b}.{a}.to_v", c, kv, bias=False); _linear(sd, g, f"{b}.{a}.to_out.0", c, c)
    _linear(sd, g, f"{b}.ff.net.0.proj", 8 * c, c); _linear(sd, g, f"{b}.ff.net.2", c, 4 * c)
    _linear(sd, g, name + ".proj_out", c, c)


def random_unet_state_dict(seed=0, channels=(320, 640, 1280, 1280), xdim=1024):
    g = torch.Generator().manual_seed(seed)
    sd, c, temb = {}, tuple(channels), 4 * channels[0]
    _conv(sd, g, "conv_in", c[0], 4, 3)
    _linear(sd, g, "time_embedding.linear_1", temb, c[0]); _linear(sd, g, "time_embedding.linear_2", temb, temb)
    skips, ch = [c[0]], c[0]
    for i in range(4):
        for j in range(2):
            _resnet(sd, g, f"down_blocks.{i}.resnets.{j}", ch, c[i], temb); ch = c[i]
            if i < 3:
                _transformer(sd, g, f"down_blocks.{i}.attentions.{j}", ch, xdim)
            skips.append(ch)
        if i < 3:
            _conv(sd, g, f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3); skips.append(ch)
    _resnet(sd, g, "mid_block.resnets.0", ch, ch, temb); _transformer(sd, g, "mid_block.attentions.0", ch, xdim)
    _resnet(sd, g, "mid_block.resnets.1", ch, ch, temb)
    rc = list(reversed(c))
    for i in range(4):
        for j in range(3):
            _resnet(sd, g, f"up_blocks.{i}.resnets.{j}", ch + skips.pop(), rc[i], temb); ch = rc[i]
            if i > 0:
                _transformer(sd, g, f"up_blocks.{i}.attentions.{j}", ch, xdim)
        if i < 3:
            _conv(sd, g, f"up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    _norm(sd, "conv_norm_out", ch); _conv(sd, g, "conv_out", 4, ch, 3)
    sd["conv_in_ref.weight"] = sd["conv_in.weight"].repeat(1, 2, 1, 1) / 2      # load_ckpt_and_modify_ref8in_tag4in.py:6-28
    sd["conv_in_ref.bias"] = sd["conv_in.bias"].clone()
    return sd


def _vae_attn(sd, g, name, c):
    _norm(sd, name + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        _linear(sd, g, f"{name}.{n}", c, c)


def random_vae_state_dict(seed=1, channels=(128, 256, 512, 512)):
    g = torch.Generator().manual_seed(seed)
    sd, c = {}, tuple(channels)
    _conv(sd, g, "encoder.conv_in", c[0], 3, 3)
    ch = c[0]
    for i in range(4):
        for j in range(2):
            _resnet(sd, g, f"encoder.down_blocks.{i}.resnets.{j}", ch, c[i]); ch = c[i]
        if i < 3:
            _conv(sd, g, f"encoder.down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
    for side in ("encoder", "decoder"):
        top = c[3]
        _resnet(sd, g, f"{side}.mid_block.resnets.0", top, top); _vae_attn(sd, g, f"{side}.mid_block.attentions.0", top)
        _resnet(sd, g, f"{side}.mid_block.resnets.1", top, top)
    _norm(sd, "encoder.conv_norm_out", ch); _conv(sd, g, "encoder.conv_out", 8, ch, 3)
    _conv(sd, g, "quant_conv", 8, 8, 1); _conv(sd, g, "post_quant_conv", 4, 4, 1)
    _conv(sd, g, "decoder.conv_in", c[3], 4, 3)
    rc, ch = list(reversed(c)), c[3]
    for i in range(4):
        for j in range(3):
            _resnet(sd, g, f"decoder.up_blocks.{i}.resnets.{j}", ch, rc[i]); ch = rc[i]
        if i < 3:
            _conv(sd, g, f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    _norm(sd, "decoder.conv_norm_out", ch); _conv(sd, g, "decoder.conv_out", 3, ch, 3)
    return sd
