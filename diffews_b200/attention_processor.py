"""KV-bank self-attention — drop-in for diffews/models/attention_processor.py (reference).

Reference protocol (attention_processor.py:39-58 MyAttention, :182-288 MyXFormersAttnProcessor):
  * `clear_bank()` sets k_bank = v_bank = None                                     (:46-50)
  * first processor call after a clear (support pass) stores K, V                   (:251-252, :262-263)
  * later calls (query pass) attend to cat([K_self, fold(K_bank)]) / same for V    (:253-267), where the k-shot
    fold is the shot-major concatenation of the k support samples of each episode  
  * out = memory_efficient_attention(q, k, v, scale=attn.scale) -> to_out[0] -> dropout(0) -> / rescale(1)

Here the projections run on the tcgen05 GEMM (q,k,v fused into one GEMM), the attention core on the KV-fused flash
kernel that streams K/V from two sources, so the concatenation and the fold are never materialised: the bank is just
the support pass's QKV buffer viewed as [B, k*S, 3C].
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .layers import FusedLinear, Linear

bf16 = torch.bfloat16


class MyAttention:
    """Prepared weights + bank state of one `attn1` (self-attention) module."""

    def __init__(self, sd, prefix, device, heads: int, wdtype=bf16, f32=False):
        self.heads = heads
        self.scale = 64 ** -0.5
        self.to_qkv = FusedLinear(sd, [prefix + ".to_q", prefix + ".to_k", prefix + ".to_v"], device, wdtype=wdtype,
                                  f32=f32)
        self.to_out = Linear(sd, prefix + ".to_out.0", device, wdtype=wdtype, f32=f32)
        self.inner_dim = self.to_qkv.splits[0]
        assert self.inner_dim == heads * 64, "the flash kernel is specialised for head_dim 64"
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.set_bank()
        self.set_myprocessor()

    # -- reference API (attention_processor.py:41-58) ---------------------------------------------------------------
    def set_bank(self):
        self.k_bank = None
        self.v_bank = None

    def clear_bank(self):
        self.k_bank = None
        self.v_bank = None

    def set_myprocessor(self):
        self.processor = MyXFormersAttnProcessor()

    @staticmethod
    def default_processor():
        return MyXFormersAttnProcessor()

    def set_processor(self, processor, _remove_lora: bool = False):
        self.processor = processor

    def get_processor(self, return_deprecated_lora: bool = False):
        return self.processor

    def set_use_memory_efficient_attention_xformers(self, flag: bool = True, attention_op=None):
        # the reference swaps MyAttnProcessor2_0 <-> MyXFormersAttnProcessor here (:60-101); there is one kernel.
        self.processor = MyXFormersAttnProcessor(attention_op)

    def __call__(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)


class _ModuleWeights:
    """16-bit copies of a stock attention module's projection weights, fused to one [3C, C] QKV GEMM."""

    def __init__(self, attn, half):
        dev = attn.to_q.weight.device
        if dev.type != "cuda":
            raise RuntimeError("MyXFormersAttnProcessor (B200 engine) needs the attention module on a CUDA device: "
                               "there is no CPU fallback")
        ws = [attn.to_q.weight, attn.to_k.weight, attn.to_v.weight]
        self.key = tuple((w.data_ptr(), w._version) for w in ws + [attn.to_out[0].weight]) + (half,)
        self.wqkv = torch.cat([w.detach() for w in ws], 0).to(half).contiguous()
        bs = [attn.to_q.bias, attn.to_k.bias, attn.to_v.bias]
        if any(b is not None for b in bs):
            self.bqkv = torch.cat([(b.detach() if b is not None else torch.zeros(w.shape[0], device=dev))
                                   for b, w in zip(bs, ws)]).float().contiguous()
        else:
            self.bqkv = None
        self.wout = attn.to_out[0].weight.detach().to(half).contiguous()
        self.bout = attn.to_out[0].bias.detach().float().contiguous() if attn.to_out[0].bias is not None else None
        self.inner = ws[0].shape[0]

    @staticmethod
    def current(attn, half):
        c = getattr(attn, "_dfw_weights", None)
        key = tuple((w.data_ptr(), w._version) for w in (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight,
                                                          attn.to_out[0].weight)) + (half,)
        if c is None or c.key != key:
            c = _ModuleWeights(attn, half)
            attn._dfw_weights = c        # plain attribute: re-fused whenever a weight is replaced or updated in place
        return c


class MyXFormersAttnProcessor:
    """`processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0)`
    with the reference signature (attention_processor.py:197-205).  `attn` is either the engine's prepared `MyAttention`
    or ANY module with the diffusers `Attention` attributes the reference processor reads (`to_q / to_k / to_v /
    to_out[0] / to_out[1]`, `heads`, `scale`, `k_bank / v_bank`, `residual_connection`, `rescale_output_factor`,
    `spatial_norm / group_norm / norm_cross`), e.g. a stock `diffusers.models.attention_processor.Attention` re-classed
    by `apply_unet_refonly_block` (unet_2d_condition.py:645-654): its fp32 (or 16-bit) hidden states [N,S,C] / [N,C,H,W]
    are cast to the tensor-core format, its Linear weights are fused once per module into a cached [3C, C] 16-bit
    matrix, and the result is returned in the input dtype.  `residual` / `out_f32` are extensions that let the engine
    fuse the transformer block's residual add into the to_out GEMM epilogue.

    Bank layout: K / V are kept as [N, S, heads*64] 16-bit views of the fused QKV buffer (the reference stores
    `head_to_batch_dim` tensors [N*heads, S, 64], :251-263); the k-shot fold (:256-258) is the view [B, k*S, C]."""

    def __init__(self, attention_op=None, half=torch.float16):
        self.attention_op = attention_op
        self.half = half

    # ---- attention core shared by both entry paths ------------------------------------------------------------------
    @staticmethod
    def _attend(attn, qkv, D, heads, scale):
        N = qkv.shape[0]
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        core = ops.attn_f32 if qkv.dtype == torch.float32 else ops.attn_kvfused      # fp32 mode: CUDA-core fp32 kernel
        if getattr(attn, "k_bank", None) is None:              # support pass: store (views keep qkv alive)  :251-252
            attn.k_bank, attn.v_bank = k, v
            return core(q, k, v, None, None, heads, scale)
        kb, vb = attn.k_bank, attn.v_bank                      # query pass: bank folded to [N, shots*S, D]   :253-267
        nb, sb = kb.shape[0], kb.shape[1]
        if nb % N != 0:
            raise ValueError(f"bank batch {nb} is not a multiple of the query batch {N}")
        shots = nb // N
        kbf = kb.as_strided((N, shots * sb, D), (shots * sb * kb.stride(1), kb.stride(1), 1), kb.storage_offset())
        vbf = vb.as_strided((N, shots * sb, D), (shots * sb * vb.stride(1), vb.stride(1), 1), vb.storage_offset())
        return core(q, k, v, kbf, vbf, heads, scale)

    def __call__(self, attn, hidden_states: torch.Tensor,
                 encoder_hidden_states: Optional[torch.Tensor] = None, attention_mask=None, temb=None,
                 scale: float = 1.0, residual: Optional[torch.Tensor] = None, out_f32: bool = False):
        if attention_mask is not None:
            raise NotImplementedError("attention masks are not part of the DiffewS hot path (reference: None)")
        if encoder_hidden_states is not None and encoder_hidden_states is not hidden_states:
            raise NotImplementedError("the KV-bank processor is self-attention only (reference: attn1)")
        if scale != 1.0:
            raise NotImplementedError("LoRA scale is not part of the hot path")
        if isinstance(attn, MyAttention):                      # engine path: prepared weights, 16-bit tokens
            assert hidden_states.ndim == 3 and hidden_states.dtype in (bf16, torch.float16, torch.float32)
            qkv = attn.to_qkv(hidden_states)                   # [N, S, 3C]  one GEMM
            o = self._attend(attn, qkv, attn.inner_dim, attn.heads, attn.scale)
            return attn.to_out(o, residual=residual, out_f32=out_f32)
        return self._call_module(attn, hidden_states)

    def _call_module(self, attn, hidden_states):
        """Reference semantics on a stock attention module (attention_processor.py:206-288)."""
        if getattr(attn, "spatial_norm", None) is not None or getattr(attn, "group_norm", None) is not None \
                or getattr(attn, "norm_cross", None):
            raise NotImplementedError("spatial_norm / group_norm / norm_cross are None on every DiffewS attn1 module")
        drop = attn.to_out[1] if len(attn.to_out) > 1 else None
        if drop is not None and getattr(drop, "p", 0.0) > 0.0 and drop.training:
            raise NotImplementedError("dropout > 0 in training mode is not part of the hot path")
        if not hidden_states.is_cuda:
            raise RuntimeError("MyXFormersAttnProcessor (B200 engine) needs CUDA tensors: there is no CPU fallback")
        in_dtype = hidden_states.dtype
        res_in = hidden_states
        ndim = hidden_states.ndim
        if ndim == 4:                                          # :213-215
            N, C, H, W = hidden_states.shape
            hidden_states = hidden_states.view(N, C, H * W).transpose(1, 2)
        w = _ModuleWeights.current(attn, self.half)
        heads = attn.heads
        if w.inner != heads * 64:
            raise NotImplementedError("the flash kernel is specialised for head_dim 64 (every SD-2.x attention layer)")
        x = hidden_states.contiguous()
        x16 = x if x.dtype == self.half else ops.cast16(x.float().contiguous(), self.half)
        qkv = ops.linear(x16, w.wqkv, w.bqkv)
        o = self._attend(attn, qkv, w.inner, heads, float(attn.scale))
        y = ops.linear(o, w.wout, w.bout, out_f32=(in_dtype == torch.float32))   # to_out[0]; to_out[1] = Dropout(0)
        if y.dtype != in_dtype:
            y = y.to(in_dtype)
        if ndim == 4:                                          # :280-281
            y = y.transpose(-1, -2).reshape(N, C, H, W)
        if getattr(attn, "residual_connection", False):        # :283-284
            y = y + res_in
        rf = getattr(attn, "rescale_output_factor", 1.0)
        return y / rf if rf != 1.0 else y                      # :286


# The reference also ships an SDPA and an unfused variant with the same bank logic (:104-180, :291-383); on the B200
# path they are the same kernel.
MyAttnProcessor2_0 = MyXFormersAttnProcessor
MyAttnProcessor = MyXFormersAttnProcessor
