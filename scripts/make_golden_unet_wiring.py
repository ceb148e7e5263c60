"""Golden vectors for the UNet wiring: the reference's OWN `MyUNet2DConditionModel.forward` and
`clear_attn_bank` (diffews/models/unet_2d_condition.py:879-1258, :656-664), executed UNMODIFIED.

The reference builds its blocks with diffusers factories (`get_down_block` ...), which are not installed, so its
`__init__` cannot run.  Instead the instance is assembled by hand — `object.__new__` + the attributes `forward` reads —
from the oracle's modules (oracle/sd21.py, reduced width, seed 0) behind adapters that only translate diffusers' keyword
call protocol (`hidden_states=, temb=, encoder_hidden_states=, res_hidden_states_tuple=, ...`) to the oracle blocks'
positional one.  What runs as written by the reference: timestep broadcasting and time embedding, the
`conv_in` / `conv_in_ref` switch on `is_target` (:1118-1121), the order in which skip tensors are pushed and popped, mid
block, up blocks, `conv_norm_out` -> act -> `conv_out`, the output dataclass, and `clear_attn_bank`'s module walk
(`isinstance_str(module, "BasicTransformerBlock")`).  The names the file imports from diffusers are placeholders.

Output: tests/golden/unet_wiring_reference.json.      python scripts/make_golden_unet_wiring.py
"""
import base64
import importlib
import json
import logging as pylogging
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_golden_attn as mga  # noqa: E402

REF = "/root/reference"


def install_stubs():
    mga.install_stubs()                                   # what diffews/models/attention_processor.py imports

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    ph = lambda n: type(n, (), {})                                                      # noqa: E731
    log = types.SimpleNamespace(get_logger=lambda n: pylogging.getLogger(n))
    mod("diffusers.configuration_utils", ConfigMixin=ph("ConfigMixin"), register_to_config=lambda f: f)
    mod("diffusers.loaders", UNet2DConditionLoadersMixin=ph("UNet2DConditionLoadersMixin"))
    mod("diffusers.utils", USE_PEFT_BACKEND=True, BaseOutput=type("BaseOutput", (), {}), deprecate=lambda *a, **k: None,
        logging=log, scale_lora_layers=lambda *a, **k: None, unscale_lora_layers=lambda *a, **k: None)
    mod("diffusers.models.activations", get_activation=lambda name: {"silu": nn.SiLU()}[name])
    mod("diffusers.models.attention_processor", ADDED_KV_ATTENTION_PROCESSORS=(), CROSS_ATTENTION_PROCESSORS=(),
        AttentionProcessor=object, AttnAddedKVProcessor=ph("AttnAddedKVProcessor"), AttnProcessor=ph("AttnProcessor"))
    mod("diffusers.models.embeddings", **{n: ph(n) for n in (
        "GaussianFourierProjection", "ImageHintTimeEmbedding", "ImageProjection", "ImageTimeEmbedding", "PositionNet",
        "TextImageProjection", "TextImageTimeEmbedding", "TextTimeEmbedding", "TimestepEmbedding", "Timesteps")})
    mod("diffusers.models.modeling_utils", ModelMixin=nn.Module)
    mod("diffusers.models.unet_2d_blocks", UNetMidBlock2D=ph("UNetMidBlock2D"),
        UNetMidBlock2DCrossAttn=ph("UNetMidBlock2DCrossAttn"), UNetMidBlock2DSimpleCrossAttn=ph("UNetMidBlock2DSimpleCrossAttn"),
        get_down_block=None, get_up_block=None)
    pkg = mod("diffews")
    pkg.__path__ = [os.path.join(REF, "diffews")]
    mp = mod("diffews.models")
    mp.__path__ = [os.path.join(REF, "diffews", "models")]


# ---- adapters: diffusers' keyword call protocol -> the oracle blocks' positional one (no arithmetic) --------------------
class TimeProj(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, timesteps):
        from oracle.sd21 import timestep_embedding
        return timestep_embedding(timesteps, self.dim)


class TimeEmbed(nn.Module):
    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, t_emb, timestep_cond=None):
        assert timestep_cond is None
        return self.inner(t_emb)


class Down(nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.blk = blk
        self.has_cross_attention = hasattr(blk, "attentions")
        self.resnets = blk.resnets

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None,
                encoder_attention_mask=None, scale=1.0):
        assert attention_mask is None and encoder_attention_mask is None and not cross_attention_kwargs
        return self.blk(hidden_states, temb, encoder_hidden_states)


class Mid(nn.Module):
    has_cross_attention = True

    def __init__(self, blk):
        super().__init__()
        self.blk = blk

    def forward(self, hidden_states, temb, encoder_hidden_states=None, attention_mask=None, cross_attention_kwargs=None,
                encoder_attention_mask=None):
        return self.blk(hidden_states, temb, encoder_hidden_states)


class Up(nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.blk = blk
        self.has_cross_attention = blk.attentions is not None
        self.resnets = blk.resnets

    def forward(self, hidden_states, temb, res_hidden_states_tuple, encoder_hidden_states=None, cross_attention_kwargs=None,
                upsample_size=None, attention_mask=None, encoder_attention_mask=None, scale=1.0):
        assert upsample_size is None
        return self.blk(hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states)


def assemble(ref_mod, ou):
    """A MyUNet2DConditionModel instance without running its diffusers-bound __init__."""
    m = object.__new__(ref_mod.MyUNet2DConditionModel)
    nn.Module.__init__(m)
    m.config = types.SimpleNamespace(center_input_sample=False, class_embed_type=None, class_embeddings_concat=False,
                                     addition_embed_type=None, encoder_hid_dim_type=None)
    m.num_upsamplers = 3
    m.time_proj = TimeProj(ou.block_out_channels[0])
    m.time_embedding = TimeEmbed(ou.time_embedding)
    m.class_embedding = m.time_embed_act = m.encoder_hid_proj = None
    m.conv_in, m.conv_in_ref = ou.conv_in, ou.conv_in_ref
    m.down_blocks = nn.ModuleList([Down(b) for b in ou.down_blocks])
    m.mid_block = Mid(ou.mid_block)
    m.up_blocks = nn.ModuleList([Up(b) for b in ou.up_blocks])
    m.conv_norm_out, m.conv_act, m.conv_out = ou.conv_norm_out, nn.SiLU(), ou.conv_out
    for mod_ in m.modules():                                   # attribute apply_unet_refonly_block / clear_attn_bank read
        if type(mod_).__name__ == "BasicTransformerBlock":
            mod_.only_cross_attention = False
    return m


def cases():
    """(B, k, latent side) — shared with tests/test_oracle.py."""
    return [(2, 1, 16), (1, 3, 8)]


def inputs(B, k, lat):
    g = torch.Generator().manual_seed(300 + 10 * B + k)
    return (torch.randn(B * k, 8, lat, lat, generator=g) * 0.8, torch.randn(B, 4, lat, lat, generator=g) * 0.8)


def main():
    install_stubs()
    ref = importlib.import_module("diffews.models.unet_2d_condition")
    from diffews_b200.synthetic import prompt_embedding
    from oracle.sd21 import build_models
    torch.set_num_threads(8)
    ou, _ = build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    m = assemble(ref, ou)
    ehs = prompt_embedding()
    out = {"made_by": "scripts/make_golden_unet_wiring.py: unmodified MyUNet2DConditionModel.forward / clear_attn_bank on a "
                      "hand-assembled instance over the oracle blocks", "cases": []}
    for B, k, lat in cases():
        sup, qry = inputs(B, k, lat)
        with torch.no_grad():
            m.clear_attn_bank()                                                          # pipeline:715
            s = m(sup, torch.tensor(1), encoder_hidden_states=ehs.repeat(B * k, 1, 1), is_target=False)
            q = m(qry, torch.tensor(1), encoder_hidden_states=ehs.repeat(B, 1, 1))
            m.clear_attn_bank()
            assert all(a.k_bank is None for a in ou.bank_attentions())
        y = q.sample
        out["cases"].append({"B": B, "k": k, "lat": lat, "type": type(q).__name__, "shape": list(y.shape),
                             "abs_mean": float(y.double().abs().mean()), "support_abs_mean": float(s.sample.double().abs().mean()),
                             "f32_b64": base64.b64encode(y.contiguous().numpy().tobytes()).decode()})
        print(B, k, lat, out["cases"][-1]["abs_mean"])
    path = os.path.join(ROOT, "tests", "golden", "unet_wiring_reference.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
