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

    def __init__(self, sd, prefix, device, heads: int, wdtype=bf16):
        self.heads = heads
        self.scale = 64 ** -0.5
        self.to_qkv = FusedLinear(sd, [prefix + ".to_q", prefix + ".to_k", prefix + ".to_v"], device, wdtype=wdtype)
        self.to_out = Linear(sd, prefix + ".to_out.0", device, wdtype=wdtype)
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

    def set_processor(self, processor):
        self.processor = processor

    def set_use_memory_efficient_attention_xformers(self, flag: bool = True, attention_op=None):
        # the reference swaps MyAttnProcessor2_0 <-> MyXFormersAttnProcessor here (:60-101); there is one kernel.
        self.processor = MyXFormersAttnProcessor(attention_op)

    def __call__(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)


class MyXFormersAttnProcessor:
    """`processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0)`
    with the reference signature (attention_processor.py:197-205); `residual` / `out_f32` are extensions that let the
    caller fuse the transformer block's residual add into the to_out GEMM epilogue."""

    def __init__(self, attention_op=None):
        self.attention_op = attention_op

    def __call__(self, attn: MyAttention, hidden_states: torch.Tensor,
                 encoder_hidden_states: Optional[torch.Tensor] = None, attention_mask=None, temb=None,
                 scale: float = 1.0, residual: Optional[torch.Tensor] = None, out_f32: bool = False):
        if encoder_hidden_states is not None or attention_mask is not None:
            raise NotImplementedError("the KV-bank processor is self-attention only (reference: attn1)")
        if scale != 1.0:
            raise NotImplementedError("LoRA scale is not part of the hot path")
        assert hidden_states.ndim == 3 and hidden_states.dtype in (bf16, torch.float16)
        N, S, C = hidden_states.shape
        qkv = attn.to_qkv(hidden_states)                       # [N, S, 3C]  one GEMM
        D = attn.inner_dim
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        if attn.k_bank is None:                                # support pass: store (views keep qkv alive)
            attn.k_bank, attn.v_bank = k, v
            o = ops.attn_kvfused(q, k, v, None, None, attn.heads, attn.scale)
        else:                                                  # query pass: bank folded to [N, shots*S, D]
            kb, vb = attn.k_bank, attn.v_bank
            nb, sb = kb.shape[0], kb.shape[1]
            if nb % N != 0:
                raise ValueError(f"bank batch {nb} is not a multiple of the query batch {N}")
            shots = nb // N
            kbf = kb.as_strided((N, shots * sb, D), (shots * sb * kb.stride(1), kb.stride(1), 1), kb.storage_offset())
            vbf = vb.as_strided((N, shots * sb, D), (shots * sb * vb.stride(1), vb.stride(1), 1), vb.storage_offset())
            o = ops.attn_kvfused(q, k, v, kbf, vbf, attn.heads, attn.scale)
        return attn.to_out(o, residual=residual, out_f32=out_f32)


# The reference also ships an SDPA and an unfused variant with the same bank logic (:104-180, :291-383); on the B200
# path they are the same kernel.
MyAttnProcessor2_0 = MyXFormersAttnProcessor
MyAttnProcessor = MyXFormersAttnProcessor
