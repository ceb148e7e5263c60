"""Golden vectors for the KV-bank self-attention, produced by the reference's OWN processor code.

diffews/models/attention_processor.py cannot be imported as is: it imports diffusers (0.25.0, requirements.txt:2) and
xformers (0.0.20, requirements.txt:10), which are not installed.  This script puts minimal stand-ins for exactly the names
that file imports into sys.modules and then executes the UNMODIFIED reference file, so the bank protocol and the k-shot
fold (attention_processor.py:251-267, the authoritative semantics) and the SDPA variant (:291-383) run as written:

  * diffusers.models.attention_processor.Attention  -> the published diffusers-0.25 helpers the processors call:
        head_to_batch_dim / batch_to_head_dim (reshape + permute), prepare_attention_mask (None in, None out),
        to_q / to_k / to_v (Linear, no bias), to_out = [Linear, Dropout(0)], scale = dim_head ** -0.5
  * xformers.ops.memory_efficient_attention(q, k, v, attn_bias=None, op=None, scale) -> its published definition,
        softmax(q k^T * scale) v on [batch*heads, tokens, dim] tensors
  * diffusers.utils (USE_PEFT_BACKEND = True so Linear layers are called without the LoRA scale argument, `deprecate`,
    `logging.get_logger`), is_xformers_available() -> True, maybe_allow_in_graph -> identity, LoRA classes -> placeholders.

Output: tests/golden/attn_reference.json (query-pass outputs of MyXFormersAttnProcessor for k = 1, 3, 5 and of
MyAttnProcessor2_0 for k = 1).   Run in the build container:  python scripts/make_golden_attn.py
"""
import importlib.util
import json
import logging as pylogging
import os
import sys
import types

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import data_tree  # noqa: E402


class Attention(nn.Module):
    """Stand-in for diffusers.models.attention_processor.Attention (0.25.0): only what the reference processors touch."""

    def __init__(self, query_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale, self.scale_qk = heads, dim_head ** -0.5, True
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(query_dim, inner, bias=False)
        self.to_v = nn.Linear(query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor

    def head_to_batch_dim(self, tensor, out_dim=3):
        h = self.heads
        b, s, d = tensor.shape
        tensor = tensor.reshape(b, s, h, d // h).permute(0, 2, 1, 3)
        return tensor.reshape(b * h, s, d // h) if out_dim == 3 else tensor

    def batch_to_head_dim(self, tensor):
        h = self.heads
        bh, s, d = tensor.shape
        return tensor.reshape(bh // h, h, s, d).permute(0, 2, 1, 3).reshape(bh // h, s, d * h)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        assert attention_mask is None
        return None

    def forward(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)


def memory_efficient_attention(query, key, value, attn_bias=None, op=None, scale=None):
    assert attn_bias is None
    s = torch.bmm(query, key.transpose(1, 2)) * (scale if scale is not None else query.shape[-1] ** -0.5)
    return torch.bmm(torch.softmax(s, dim=-1), value)


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    log = types.SimpleNamespace(get_logger=lambda n: pylogging.getLogger(n))
    mod("diffusers")
    mod("diffusers.utils", USE_PEFT_BACKEND=True, deprecate=lambda *a, **k: None, logging=log)
    mod("diffusers.utils.import_utils", is_xformers_available=lambda: True)
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    mod("diffusers.models")
    mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear, LoRALinearLayer=nn.Linear)
    mod("diffusers.models.attention_processor", Attention=Attention, __all__=["Attention"])
    ops = mod("xformers.ops", memory_efficient_attention=memory_efficient_attention)
    mod("xformers", ops=ops)


def scheduler_golden():
    """marigold/util/scheduler_customized.py executed UNMODIFIED over stand-ins for its diffusers imports (the base
    classes as empty classes, `register_to_config` as the identity): DDIMSchedulerCustomized only overrides __init__
    and _get_variance (:107-180), so what this pins is the beta / alpha tables the reference builds from
    scheduler_1.0_1.0/scheduler_config.json — the premise (every alpha_cumprod == 0) of the z0 = -v collapse.
    set_timesteps / step are stock diffusers code (restated in diffews_b200/scheduler.py and the oracle)."""
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    base = type("DDIMScheduler", (), {})
    mod("diffusers", DDIMScheduler=base, DDPMScheduler=type("DDPMScheduler", (), {}))
    mod("diffusers.configuration_utils", ConfigMixin=object, register_to_config=lambda f: f)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_ddim", DDIMSchedulerOutput=dict)
    mod("diffusers.utils")
    mod("diffusers.utils.torch_utils", randn_tensor=torch.randn)
    spec = importlib.util.spec_from_file_location("ref_scheduler", "/root/reference/marigold/util/scheduler_customized.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cfg = json.load(open("/root/reference/scheduler_1.0_1.0/scheduler_config.json"))
    import inspect
    allowed = set(inspect.signature(ref.DDIMSchedulerCustomized.__init__).parameters) - {"self"}
    kw = {k: v for k, v in cfg.items() if k in allowed}            # what ConfigMixin.from_config passes on
    sch = ref.DDIMSchedulerCustomized(**kw)
    return {"config_used": kw, "betas_head": sch.betas[:3].tolist(), "betas_min": float(sch.betas.min()),
            "betas_max": float(sch.betas.max()), "alphas_cumprod_abs_max": float(sch.alphas_cumprod.abs().max()),
            "alphas_cumprod_len": int(sch.alphas_cumprod.numel()), "final_alpha_cumprod": float(sch.final_alpha_cumprod),
            "init_noise_sigma": float(sch.init_noise_sigma), "timesteps_head": sch.timesteps[:3].tolist(),
            "variance_t1_prev0": float(sch._get_variance(1, 0)), "variance_t1_prevneg": float(sch._get_variance(1, -999))}


def main():
    sg = scheduler_golden()
    with open(os.path.join(ROOT, "tests", "golden", "scheduler_reference.json"), "w") as f:
        json.dump({"made_by": "scripts/make_golden_attn.py: unmodified marigold/util/scheduler_customized.py "
                              "DDIMSchedulerCustomized.__init__ / _get_variance over import stand-ins", **sg}, f, indent=1)
    print("scheduler:", sg["alphas_cumprod_abs_max"], sg["final_alpha_cumprod"], sg["variance_t1_prev0"])
    install_stubs()
    spec = importlib.util.spec_from_file_location("ref_attention_processor",
                                                  "/root/reference/diffews/models/attention_processor.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {"made_by": "scripts/make_golden_attn.py: unmodified diffews/models/attention_processor.py "
                      "(MyAttention + MyXFormersAttnProcessor / MyAttnProcessor2_0) over stand-ins for the diffusers / "
                      "xformers names it imports", "cases": []}
    for c in data_tree.attn_cases():
        attn = ref.MyAttention(c["C"], c["heads"], c["C"] // c["heads"])
        with torch.no_grad():
            attn.to_q.weight.copy_(c["w"]["to_q"]); attn.to_k.weight.copy_(c["w"]["to_k"])
            attn.to_v.weight.copy_(c["w"]["to_v"]); attn.to_out[0].weight.copy_(c["w"]["to_out"])
            attn.to_out[0].bias.copy_(c["w"]["to_out_bias"])
        rec = {"B": c["B"], "k": c["k"]}
        procs = {"xformers": ref.MyXFormersAttnProcessor()}
        if c["k"] == 1:
            procs["sdpa"] = ref.MyAttnProcessor2_0()              # the SDPA variant only supports k = 1 (:354-359)
        for name, proc in procs.items():
            attn.set_bank()                                        # unet_2d_condition.py:645-654 / clear_attn_bank
            attn.set_processor(proc)
            with torch.no_grad():
                sup = attn(c["x_support"])                         # support pass: stores K, V
                qry = attn(c["x_query"])                           # query pass: attends to [self ; folded bank]
            attn.clear_bank()
            rec[name] = {"support_out_sum": float(sup.double().sum()), "query_out": qry.flatten().tolist()}
        out["cases"].append(rec)
    path = os.path.join(ROOT, "tests", "golden", "attn_reference.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
