"""Training step of the DiffewS UNet on the B200 kernels.

Reference: train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py
  :1374-1375  model_pred_cond_ref = unet(latents_rgb_cond_ref, t, ehs_nshot, is_target=False)   (fills the K/V banks)
              model_pred          = unet(latents_tag,          t, ehs,       is_target=True)    (attends to them)
  :1376-1379  unet.clear_attn_bank()
  :1381-1384  loss = F.mse_loss((model_pred.float() + model_pred_cond_ref.float() * 0.).float(), target.float())
  :1386       accelerator.backward(loss)              -> every op below has a hand-written backward kernel
  :1226       accelerator.prepare(unet, ...)          -> DDP: gradient all-reduce overlapped with the backward
  :1393-1396  clip_grad_norm_, optimizer.step(), optimizer.zero_grad()

Design.  torch.autograd is only the TAPE for activation gradients: every node is a `torch.autograd.Function` whose
forward and backward call the C ABI (conv / linear forward kernels reused for the data gradients with permuted weights,
`dfw_conv_wgrad` for the weight gradients, the norm / GEGLU / attention backward kernels).  Parameters are NOT autograd
leaves: each `Param` owns an fp32 master tensor (GEMM layout — AdamW is elementwise, so the layout is free), an fp32
gradient that the backward kernels write in place (first use of a step overwrites, later uses accumulate: every weight
is used by the support pass AND the query pass), and a 16-bit operand copy emitted by the optimizer step.  All
gradients live in ONE flat buffer laid out in reverse execution order, so a DDP bucket is a contiguous slice of it and
is all-reduced (NCCL, its own stream) as soon as its last parameter has received its last contribution, while the rest
of the backward is still running.  Gradients flow from the query pass into the support pass through the K/V banks
(the reference does not detach them, attention_processor.py:251-267).

Numerics: 16-bit activations and operands (fp16 by default, like the reference's fp16 autocast), fp32 accumulation,
fp32 master weights / moments, static loss scale (the reference runs a GradScaler).  CUDA only — no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from . import ops
from .optim import AdamW, mse_loss
from .weights import conv_weight_to_gemm

f16 = torch.float16


# ---------------------------------------------------------------------------------------------------------------------
# parameters
# ---------------------------------------------------------------------------------------------------------------------
class Param:
    """One trainable tensor.  `w` fp32 master, `grad` fp32 (views into the store's flat buffers), `h` 16-bit copy of `w`;
    `kind`: "conv" (GEMM layout [Cout, T*Cin]), "linear" ([Nout, K]), "vec" (bias / norm affine)."""

    def __init__(self, name: str, init: torch.Tensor, kind: str, meta: Optional[dict] = None):
        self.name, self.kind, self.meta = name, kind, meta or {}
        self.init = init.detach().float().contiguous().cpu()
        self.shape = tuple(self.init.shape)
        self.numel = self.init.numel()
        self.w = self.grad = self.h = None
        self.bt = None            # derived 16-bit operand of the data-gradient kernel (refreshed after every step)
        self.count = 0            # gradient contributions received in the current step
        self.expected = None      # contributions per step (learned on the first step; the graph is static)
        self.bucket = None

    # gradient contribution from a small fp32 tensor produced by a backward kernel (norm affine, biases)
    def add_grad(self, g: torch.Tensor, scale: float, store: "ParamStore"):
        g = g.reshape(self.shape)
        if self.count == 0:
            torch.mul(g, scale, out=self.grad)
        else:
            self.grad.add_(g, alpha=scale)
        store.contributed(self)


class ParamStore:
    def __init__(self, device, half=f16):
        self.device, self.half = torch.device(device), half
        self.params: List[Param] = []
        self.by_name: Dict[str, Param] = {}
        self.reducer: Optional["GradReducer"] = None
        self.grad_scale = 1.0          # 1 / (loss_scale * world_size), applied where a gradient is written

    def add(self, name, init, kind, meta=None) -> Param:
        p = Param(name, init, kind, meta)
        self.params.append(p)
        self.by_name[name] = p
        return p

    def finalize(self):
        """Flat buffers in REVERSE registration (= execution) order: the parameters whose gradients complete first come
        first, so DDP buckets are contiguous slices that fill front to back during the backward."""
        order = list(reversed(self.params))
        offs, total = [], 0
        for p in order:
            offs.append(total)
            total += (p.numel + 7) // 8 * 8            # 32-byte aligned fp32 / 16-byte aligned 16-bit slices
        self.total = total
        self.flat_w = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=self.device, dtype=torch.float32)
        self.flat_h = torch.zeros(total, device=self.device, dtype=self.half)
        for p, o in zip(order, offs):
            p.offset = o
            p.w = self.flat_w[o:o + p.numel].view(p.shape)
            p.grad = self.flat_g[o:o + p.numel].view(p.shape)
            p.h = self.flat_h[o:o + p.numel].view(p.shape)
            p.w.copy_(p.init)
            p.h.copy_(p.w)
            p.w.grad = p.grad                          # what optim.AdamW reads
            p.init = None
        self.order = order

    def begin_step(self):
        for p in self.params:
            p.count = 0
        if self.reducer is not None:
            self.reducer.begin_step()

    def contributed(self, p: Param):
        p.count += 1
        if self.reducer is not None:
            self.reducer.contributed(p)

    def end_step(self):
        missing = [p.name for p in self.params if p.count == 0]
        if missing:
            raise RuntimeError(f"no gradient reached {len(missing)} parameters, e.g. {missing[:4]}")
        for p in self.params:
            if p.expected is None:
                p.expected = p.count
        if self.reducer is not None:
            self.reducer.finish()


class GradReducer:
    """DDP-style bucketed gradient all-reduce (SUM; the 1 / world_size factor is folded into `ParamStore.grad_scale`),
    overlapped with the backward: bucket b = a contiguous slice of the flat gradient buffer; it is reduced on the
    process group's stream as soon as every parameter in it has received its `expected` number of contributions.
    The first step learns `expected` and reduces everything at the end.  Works on any backend (NCCL on the GPUs; the
    gloo CPU test drives it with CPU tensors)."""

    def __init__(self, store: ParamStore, group=None, bucket_bytes: int = 64 << 20):
        import torch.distributed as dist
        self.dist, self.group, self.store = dist, group, store
        self.world = dist.get_world_size(group)
        self.buckets = []                       # (start, end, [params])
        cur, start, size = [], 0, 0
        for p in store.order:
            cur.append(p)
            size += p.numel * 4
            if size >= bucket_bytes:
                end = p.offset + (p.numel + 7) // 8 * 8
                self.buckets.append((start, end, cur))
                cur, start, size = [], end, 0
        if cur:
            self.buckets.append((start, store.total, cur))
        for b, (_, _, ps) in enumerate(self.buckets):
            for p in ps:
                p.bucket = b
        self.works, self.pending, self.launched = [], [], []
        self.overlapped = 0                    # buckets launched before the end of the backward (last step)

    def begin_step(self):
        self.works = []
        # -1: some parameter of the bucket has no learned contribution count yet (first step) -> reduced in finish()
        self.pending = [len(ps) if all(p.expected is not None for p in ps) else -1 for (_, _, ps) in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.overlapped = 0

    def _launch(self, b):
        s, e, _ = self.buckets[b]
        self.works.append(self.dist.all_reduce(self.store.flat_g[s:e], group=self.group, async_op=True))
        self.launched[b] = True

    def contributed(self, p: Param):
        b = p.bucket
        if self.pending[b] < 0 or p.expected is None or p.count != p.expected:
            return
        self.pending[b] -= 1
        if self.pending[b] == 0 and not self.launched[b]:
            self._launch(b)
            self.overlapped += 1

    def finish(self):
        for b in range(len(self.buckets)):
            if not self.launched[b]:
                self._launch(b)
        for w in self.works:
            w.wait()
        self.works = []


# ---------------------------------------------------------------------------------------------------------------------
# autograd nodes (activation gradients only; parameter gradients are written in place by the kernels)
# ---------------------------------------------------------------------------------------------------------------------
def _wgrad(store: ParamStore, p: Param, x, dy, ksize, stride=1, cout_store=None):
    ops.conv_wgrad(x, dy, p.grad.view(p.grad.shape[0], -1), ksize=ksize, stride=stride, scale=store.grad_scale,
                   accumulate=p.count > 0, cout_store=cout_store)
    store.contributed(p)


def _bias_grad(store: ParamStore, b: Optional[Param], dy):
    """Column sums of dy: returns the (loss-scaled) fp32 [C] vector, and adds it (unscaled) to the bias gradient."""
    s = ops.colsum(dy)
    if b is not None:
        b.add_grad(s, store.grad_scale, store)
    return s


class ConvFn(torch.autograd.Function):
    """y = conv(x, w) + bias (+ extra per-channel bias, e.g. the time-embedding projection) (+ residual)."""

    @staticmethod
    def forward(ctx, x, residual, extra_bias, layer):
        bias = layer.b.w if layer.b is not None else None
        if extra_bias is not None:
            bias = extra_bias.float().reshape(-1) + (bias if bias is not None else 0.0)
        y = ops.conv2d(x, layer.w.h.view(layer.cout, -1), bias, ksize=layer.ksize, stride=layer.stride, residual=residual)
        ctx.layer = layer
        ctx.save_for_backward(x)
        ctx.has_res, ctx.has_extra = residual is not None, extra_bias is not None
        ctx.extra_shape = extra_bias.shape if extra_bias is not None else None
        ctx.extra_dtype = extra_bias.dtype if extra_bias is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        layer, st = ctx.layer, ctx.layer.store
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            if layer.stride == 1:
                dx = ops.conv2d(dy, layer.w.bt, ksize=layer.ksize)
            else:
                dx = ops.upconv2x(dy, layer.w.bt)
        _wgrad(st, layer.w, x, dy, layer.ksize, layer.stride)
        s = None
        if layer.b is not None or ctx.has_extra:
            s = _bias_grad(st, layer.b, dy)
        d_extra = s.reshape(ctx.extra_shape).to(ctx.extra_dtype) if ctx.has_extra else None
        return dx, (dy if ctx.has_res else None), d_extra, None


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, layer):
        y = ops.linear(x, layer.w.h, layer.b.w if layer.b is not None else None, residual=residual)
        ctx.layer = layer
        ctx.save_for_backward(x)
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        layer, st = ctx.layer, ctx.layer.store
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.linear(dy, layer.w.bt) if ctx.needs_input_grad[0] else None
        K, Nout = x.shape[-1], dy.shape[-1]
        M = x.numel() // K
        _wgrad(st, layer.w, x.view(1, 1, M, K), dy.view(1, 1, M, Nout), 1)
        if layer.b is not None:
            _bias_grad(st, layer.b, dy)
        return dx, (dy if ctx.has_res else None), None


class GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer, silu):
        ctx.layer, ctx.silu = layer, silu
        ctx.save_for_backward(x)
        return ops.groupnorm(x, layer.g.w, layer.b.w, groups=32, eps=layer.eps, silu=silu, out_dtype=x.dtype)

    @staticmethod
    def backward(ctx, dy):
        layer, st = ctx.layer, ctx.layer.store
        (x,) = ctx.saved_tensors
        dx, dg, db = ops.groupnorm_backward(x, dy.contiguous(), layer.g.w, layer.b.w, groups=32, eps=layer.eps, silu=ctx.silu)
        layer.g.add_grad(dg, st.grad_scale, st)
        layer.b.add_grad(db, st.grad_scale, st)
        return dx, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, layer):
        ctx.layer = layer
        ctx.save_for_backward(x)
        return ops.layernorm(x, layer.g.w, layer.b.w, layer.eps, out_dtype=x.dtype)

    @staticmethod
    def backward(ctx, dy):
        layer, st = ctx.layer, ctx.layer.store
        (x,) = ctx.saved_tensors
        dx, dg, db = ops.layernorm_backward(x, dy.contiguous(), layer.g.w, layer.eps)
        layer.g.add_grad(dg, st.grad_scale, st)
        layer.b.add_grad(db, st.grad_scale, st)
        return dx, None


class GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        ctx.save_for_backward(h)
        return ops.geglu(h)

    @staticmethod
    def backward(ctx, dy):
        (h,) = ctx.saved_tensors
        return ops.geglu_backward(h, dy.contiguous())


class AttnFn(torch.autograd.Function):
    """softmax(q [k_self ; k_bank]^T * scale) [v_self ; v_bank] on the KV-fused flash kernel; flash-style backward."""

    @staticmethod
    def forward(ctx, q, k, v, kb, vb, heads, scale):
        o, lse = ops.attn_kvfused(q, k, v, kb, vb, heads, scale, return_lse=True)
        ctx.save_for_backward(q, k, v, kb, vb, o, lse)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, kb, vb, o, lse = ctx.saved_tensors
        dq, dk, dv, dkb, dvb = ops.attn_kvfused_backward(q, k, v, kb, vb, o, do.contiguous(), ctx.heads, ctx.scale, lse=lse)
        return dq, dk, dv, dkb, dvb, None, None


class UpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2x(x)

    @staticmethod
    def backward(ctx, dy):
        return ops.downsum2x(dy.contiguous())


class ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca = a.shape[-1]
        return ops.concat_channels(a, b)

    @staticmethod
    def backward(ctx, dy):
        return ops.split_channels(dy.contiguous(), ctx.ca)


class ConvInFn(torch.autograd.Function):
    """conv_in / conv_in_ref on the NCHW fp32 latents (im2col + GEMM); the latents need no gradient."""

    @staticmethod
    def forward(ctx, x_nchw, anchor, layer):
        cols = ops.im2col3x3_small(x_nchw, layer.kpad, layer.store.half)
        ctx.layer = layer
        ctx.save_for_backward(cols)
        return ops.conv2d(cols, layer.w.h, layer.b.w, ksize=1)

    @staticmethod
    def backward(ctx, dy):
        layer, st = ctx.layer, ctx.layer.store
        (cols,) = ctx.saved_tensors
        dy = dy.contiguous()
        _wgrad(st, layer.w, cols, dy, 1)
        _bias_grad(st, layer.b, dy)
        return None, None, None


class ConvOutFn(torch.autograd.Function):
    """conv_out (C -> 4) with the reference's NCHW fp32 output; the 4-channel gradient is zero-padded to 64 channels so
    that the data / weight gradients run on the tensor-core kernels."""

    @staticmethod
    def forward(ctx, x, layer):
        N, H, W, _ = x.shape
        y = ops.conv2d(x, layer.w.h.view(layer.cout, -1), layer.b.w, ksize=3, out_f32=True)
        ctx.layer = layer
        ctx.save_for_backward(x)
        return ops.nhwc_f32_to_nchw(y.view(N, H * W, layer.cout), layer.cout, H, W)

    @staticmethod
    def backward(ctx, dpred):
        layer, st = ctx.layer, ctx.layer.store
        (x,) = ctx.saved_tensors
        dyp = ops.nchw_to_nhwc_pad(dpred.contiguous().float(), 64, st.half)
        dx = ops.conv2d(dyp, layer.w.bt, ksize=3)
        _wgrad(st, layer.w, x, dyp, 3, cout_store=layer.cout)
        s = ops.colsum(dyp)
        layer.b.add_grad(s.view(-1)[:layer.cout], st.grad_scale, st)
        return dx, None


# ---------------------------------------------------------------------------------------------------------------------
# layers (forward composition; diffusers state-dict names)
# ---------------------------------------------------------------------------------------------------------------------
_ROT9 = [8 - t for t in range(9)]
_K_OF = {(0, 0): -1, (0, 1): 1, (1, 0): 2, (1, 1): 0}     # stride-2 dgrad: (output phase, tap a) -> forward kernel index


class TConv:
    def __init__(self, store, sd, prefix, stride=1):
        w = sd[prefix + ".weight"].detach().float()
        self.store, self.stride = store, stride
        self.cout, self.cin, self.ksize, _ = w.shape
        self.w = store.add(prefix + ".weight", conv_weight_to_gemm(w), "conv", {"oihw": tuple(w.shape)})
        self.b = store.add(prefix + ".bias", sd[prefix + ".bias"], "vec")

    def refresh(self):
        T = self.ksize * self.ksize
        if self.stride == 1:
            tm = _ROT9 if T == 9 else [0]
            self.w.bt = ops.weight_permute(self.w.h, self.cout, T, self.cin, tm, out=self.w.bt).view(self.cin, T * self.cout)
        else:
            if self.w.bt is None:
                self.w.bt = torch.empty(4, self.cin, 4 * self.cout, device=self.w.h.device, dtype=self.w.h.dtype)
            for ph in range(2):
                for pw in range(2):
                    tm = []
                    for a in range(2):
                        for b in range(2):
                            kh, kw = _K_OF[(ph, a)], _K_OF[(pw, b)]
                            tm.append(-1 if kh < 0 or kw < 0 else kh * 3 + kw)
                    ops.weight_permute(self.w.h, self.cout, 9, self.cin, tm, out=self.w.bt[ph * 2 + pw])

    def __call__(self, x, residual=None, extra_bias=None):
        return ConvFn.apply(x, residual, extra_bias, self)


class TLinear:
    def __init__(self, store, sd, prefix):
        w = sd[prefix + ".weight"].detach().float()
        self.store = store
        self.nout, self.k = w.shape
        self.w = store.add(prefix + ".weight", w, "linear")
        b = sd.get(prefix + ".bias")
        self.b = store.add(prefix + ".bias", b, "vec") if b is not None else None

    def refresh(self):
        self.w.bt = ops.weight_permute(self.w.h, self.nout, 1, self.k, [0], out=self.w.bt).view(self.k, self.nout)

    def __call__(self, x, residual=None):
        return LinearFn.apply(x, residual, self)


class TNorm:
    def __init__(self, store, sd, prefix, eps):
        self.store, self.eps = store, eps
        self.g = store.add(prefix + ".weight", sd[prefix + ".weight"], "vec")
        self.b = store.add(prefix + ".bias", sd[prefix + ".bias"], "vec")


class TConvIn:
    def __init__(self, store, sd, prefix):
        w = sd[prefix + ".weight"].detach().float()
        self.store = store
        self.cout, self.cin = w.shape[0], w.shape[1]
        k = 9 * self.cin
        self.kpad = (k + 63) // 64 * 64
        wg = torch.zeros(self.cout, self.kpad)
        wg[:, :k] = conv_weight_to_gemm(w.cpu())
        self.w = store.add(prefix + ".weight", wg, "conv_in", {"oihw": tuple(w.shape)})
        self.b = store.add(prefix + ".bias", sd[prefix + ".bias"], "vec")

    def refresh(self):
        pass

    def __call__(self, x_nchw, anchor):
        return ConvInFn.apply(x_nchw, anchor, self)


class TConvOut:
    def __init__(self, store, sd, prefix):
        w = sd[prefix + ".weight"].detach().float()
        self.store = store
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.w = store.add(prefix + ".weight", conv_weight_to_gemm(w), "conv", {"oihw": tuple(w.shape)})
        self.b = store.add(prefix + ".bias", sd[prefix + ".bias"], "vec")
        self._pad = None

    def refresh(self):
        if self._pad is None:
            self._pad = torch.zeros(64, 9 * self.cin, device=self.w.h.device, dtype=self.w.h.dtype)
        self._pad[:self.cout].copy_(self.w.h.view(self.cout, -1))
        self.w.bt = ops.weight_permute(self._pad, 64, 9, self.cin, _ROT9, out=self.w.bt).view(self.cin, 9 * 64)

    def __call__(self, x):
        return ConvOutFn.apply(x, self)


class TResnet:
    def __init__(self, store, sd, prefix):
        self.norm1 = TNorm(store, sd, prefix + ".norm1", 1e-5)
        self.conv1 = TConv(store, sd, prefix + ".conv1")
        self.tproj = TLinear(store, sd, prefix + ".time_emb_proj")
        self.norm2 = TNorm(store, sd, prefix + ".norm2", 1e-5)
        self.conv2 = TConv(store, sd, prefix + ".conv2")
        self.shortcut = TConv(store, sd, prefix + ".conv_shortcut") if (prefix + ".conv_shortcut.weight") in sd else None
        self.layers = [self.conv1, self.tproj, self.conv2] + ([self.shortcut] if self.shortcut else [])

    def __call__(self, x, temb_bias):
        t = self.conv1(GroupNormFn.apply(x, self.norm1, True), extra_bias=temb_bias)
        s = x if self.shortcut is None else self.shortcut(x)
        return self.conv2(GroupNormFn.apply(t, self.norm2, True), residual=s)


class TAttention:
    def __init__(self, store, sd, prefix, heads, cross):
        self.heads, self.scale, self.cross = heads, 64 ** -0.5, cross
        self.to_q = TLinear(store, sd, prefix + ".to_q")
        self.to_k = TLinear(store, sd, prefix + ".to_k")
        self.to_v = TLinear(store, sd, prefix + ".to_v")
        self.to_out = TLinear(store, sd, prefix + ".to_out.0")
        self.k_bank = self.v_bank = None
        self.layers = [self.to_q, self.to_k, self.to_v, self.to_out]

    def __call__(self, x, residual, ehs=None, bank_only=False):
        ctx = x if ehs is None else ehs
        k, v = self.to_k(ctx), self.to_v(ctx)
        kb = vb = None
        if not self.cross:
            if self.k_bank is None:                         # support pass: store (attention_processor.py:251-252)
                self.k_bank, self.v_bank = k, v
                if bank_only:
                    return None
            else:                                           # query pass: k-shot fold == shot-major concat (:253-267)
                B = x.shape[0]
                kb = self.k_bank.reshape(B, -1, k.shape[-1])
                vb = self.v_bank.reshape(B, -1, v.shape[-1])
        q = self.to_q(x)
        o = AttnFn.apply(q, k, v, kb, vb, self.heads, self.scale)
        return self.to_out(o, residual=residual)


class TTransformer:
    def __init__(self, store, sd, prefix, heads):
        b = prefix + ".transformer_blocks.0"
        self.norm = TNorm(store, sd, prefix + ".norm", 1e-6)
        self.proj_in = TLinear(store, sd, prefix + ".proj_in")
        self.norm1 = TNorm(store, sd, b + ".norm1", 1e-5)
        self.attn1 = TAttention(store, sd, b + ".attn1", heads, cross=False)
        self.norm2 = TNorm(store, sd, b + ".norm2", 1e-5)
        self.attn2 = TAttention(store, sd, b + ".attn2", heads, cross=True)
        self.norm3 = TNorm(store, sd, b + ".norm3", 1e-5)
        self.ff1 = TLinear(store, sd, b + ".ff.net.0.proj")
        self.ff2 = TLinear(store, sd, b + ".ff.net.2")
        self.proj_out = TLinear(store, sd, prefix + ".proj_out")
        self.layers = [self.proj_in] + self.attn1.layers + self.attn2.layers + [self.ff1, self.ff2, self.proj_out]

    def __call__(self, h, ehs, bank_only=False):
        N, H, W, C = h.shape
        x = self.proj_in(GroupNormFn.apply(h, self.norm, False).view(N, H * W, C))
        x1 = self.attn1(LayerNormFn.apply(x, self.norm1), residual=x, bank_only=bank_only)
        if x1 is None:
            return None
        x2 = self.attn2(LayerNormFn.apply(x1, self.norm2), residual=x1, ehs=ehs)
        g = GegluFn.apply(self.ff1(LayerNormFn.apply(x2, self.norm3)))
        x3 = self.ff2(g, residual=x2)
        return self.proj_out(x3, residual=h.view(N, H * W, C)).view(N, H, W, C)


class TrainableUNet:
    """MyUNet2DConditionModel (unet_2d_condition.py:879-1258, SD-2.1 config + conv_in_ref) with a backward."""

    def __init__(self, state_dict, device="cuda", block_out_channels=(320, 640, 1280, 1280), heads=(5, 10, 20, 20),
                 half=f16):
        sd = state_dict
        self.store = st = ParamStore(device, half)
        self.c = c = tuple(block_out_channels)
        self.device, self.half = torch.device(device), half
        self.layers = []           # every object with a refresh() (derived dgrad operands)

        def reg(obj):
            self.layers.extend(getattr(obj, "layers", [obj]))
            return obj

        self.conv_in = reg(TConvIn(st, sd, "conv_in"))
        self.conv_in_ref = reg(TConvIn(st, sd, "conv_in_ref"))
        self.te1 = reg(TLinear(st, sd, "time_embedding.linear_1"))
        self.te2 = reg(TLinear(st, sd, "time_embedding.linear_2"))
        self.down = []
        for i in range(4):
            res = [reg(TResnet(st, sd, f"down_blocks.{i}.resnets.{j}")) for j in range(2)]
            att = [reg(TTransformer(st, sd, f"down_blocks.{i}.attentions.{j}", heads[i])) for j in range(2)] if i < 3 else []
            dn = reg(TConv(st, sd, f"down_blocks.{i}.downsamplers.0.conv", stride=2)) if i < 3 else None
            self.down.append((res, att, dn))
        self.mid = (reg(TResnet(st, sd, "mid_block.resnets.0")), reg(TTransformer(st, sd, "mid_block.attentions.0", heads[3])),
                    reg(TResnet(st, sd, "mid_block.resnets.1")))
        rh = list(reversed(heads))
        self.up = []
        for i in range(4):
            res = [reg(TResnet(st, sd, f"up_blocks.{i}.resnets.{j}")) for j in range(3)]
            att = [reg(TTransformer(st, sd, f"up_blocks.{i}.attentions.{j}", rh[i])) for j in range(3)] if i > 0 else []
            up = reg(TConv(st, sd, f"up_blocks.{i}.upsamplers.0.conv")) if i < 3 else None
            self.up.append((res, att, up))
        self.conv_norm_out = TNorm(st, sd, "conv_norm_out", 1e-5)
        self.conv_out = reg(TConvOut(st, sd, "conv_out"))
        st.finalize()
        self.refresh_operands()
        self.transformers = [t for (_, att, _) in self.down for t in att] + [self.mid[1]] + \
                            [t for (_, att, _) in self.up for t in att]
        self.resnets = [r for (res, _, _) in self.down for r in res] + [self.mid[0], self.mid[2]] + \
                       [r for (res, _, _) in self.up for r in res]

    @classmethod
    def from_module(cls, module, device="cuda", **kw):
        cfg = {}
        if hasattr(module, "block_out_channels"):
            cfg = dict(block_out_channels=module.block_out_channels, heads=module.heads)
        cfg.update(kw)
        return cls(module.state_dict(), device=device, **cfg)

    def refresh_operands(self):
        """Rebuild the permuted 16-bit operands of the data-gradient kernels from the 16-bit weight copies (after a step)."""
        for l in self.layers:
            l.refresh()

    def clear_attn_bank(self):                                   # unet_2d_condition.py:656-664
        for t in self.transformers:
            t.attn1.k_bank = t.attn1.v_bank = None

    def parameters(self):
        return [p.w for p in self.store.params]

    # -- torch-layout views of the parameters / gradients (checkpointing, parity tests) ----------------------------------
    def _to_torch_layout(self, p: Param, t: torch.Tensor) -> torch.Tensor:
        if p.kind == "conv":
            co, ci, kh, kw = p.meta["oihw"]
            return t.view(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        if p.kind == "conv_in":
            co, ci, kh, kw = p.meta["oihw"]
            return t[:, :kh * kw * ci].reshape(co, kh, kw, ci).permute(0, 3, 1, 2).contiguous()
        return t.clone()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {p.name: self._to_torch_layout(p, p.w) for p in self.store.params}

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return {p.name: self._to_torch_layout(p, p.grad) for p in self.store.params}

    # -- forward -----------------------------------------------------------------------------------------------------
    def _time_act(self, t_value: float):
        """silu(time_embedding(t)) [1, 4*c0] (unet_2d_condition.py:1008-1015): the timestep is one scalar per step
        (train...v3.py:1365), so the embedding is computed once and every ResnetBlock2D adds the same projection."""
        dim = self.c[0]
        half = dim // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=self.device) / half
        ang = torch.full((1, 1), float(t_value), device=self.device) * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1).to(self.half)             # flip_sin_to_cos=True
        emb.requires_grad_()       # tape anchor: parameters are not autograd leaves, so a graph input has to be one
        e = self.te2(torch.nn.functional.silu(self.te1(emb)))
        return torch.nn.functional.silu(e)

    def time_biases(self, timestep):
        """time_emb_proj(silu(time_embedding(t))) of every ResnetBlock2D, in execution order ([1, Cout] each).  The support
        and the query pass of a step use the same timestep (train...v3.py:1365), so a step evaluates this chain once and
        both passes add the same vectors; their gradients meet at these tensors."""
        t_value = float(timestep.reshape(-1)[0]) if torch.is_tensor(timestep) else float(timestep)
        act = self._time_act(t_value)
        return [r.tproj(act) for r in self.resnets]

    def forward(self, sample, timestep, encoder_hidden_states, is_target: bool = True, temb_biases=None):
        """sample fp32 NCHW latents ([N,4,h,w] target / [N,8,h,w] support), encoder_hidden_states [N or 1, L, 1024].
        Returns the fp32 NCHW prediction; the support pass (is_target=False) stops once the last K/V bank is filled and
        returns None (its output only ever enters the loss multiplied by 0, train...v3.py:1381)."""
        if not sample.is_cuda:
            raise RuntimeError("TrainableUNet (B200 engine) needs CUDA tensors: there is no CPU fallback")
        x = sample.to(torch.float32).contiguous()
        N = x.shape[0]
        tb = iter(temb_biases if temb_biases is not None else self.time_biases(timestep))
        ehs = encoder_hidden_states.detach().to(device=self.device, dtype=self.half)
        if ehs.shape[0] == 1 and N > 1:
            ehs = ehs.expand(N, -1, -1)               # encoder_hidden_states.repeat(temp_nshot, 1, 1), train...v3.py:1370
        ehs = ehs.contiguous().clone().requires_grad_()          # tape anchor for attn2.to_k / to_v (see _time_act)
        anchor = torch.zeros((), device=self.device, requires_grad=True)
        h = (self.conv_in if is_target else self.conv_in_ref)(x, anchor)
        skips = [h]
        for res, att, dn in self.down:
            for j, r in enumerate(res):
                h = r(h, next(tb))
                if att:
                    h = att[j](h, ehs)
                skips.append(h)
            if dn is not None:
                h = dn(h)
                skips.append(h)
        h = self.mid[0](h, next(tb))
        h = self.mid[1](h, ehs)
        h = self.mid[2](h, next(tb))
        for bi, (res, att, up) in enumerate(self.up):
            for j, r in enumerate(res):
                h = r(ConcatFn.apply(h, skips.pop()), next(tb))
                if att:
                    last = (not is_target) and bi == 3 and j == 2
                    h = att[j](h, ehs, bank_only=last)
                    if h is None:
                        return None
            if up is not None:
                h = up(UpsampleFn.apply(h))
        return self.conv_out(GroupNormFn.apply(h, self.conv_norm_out, True))

    __call__ = forward


# ---------------------------------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------------------------------
class Trainer:
    """One optimisation step of the reference loop (train...v3.py:1337-1396) on latents.

        trainer = Trainer(TrainableUNet(sd), lr=..., max_grad_norm=1.0)            # + process_group for DDP
        loss = trainer.step(latents_rgb_cond_ref [k,8,h,w], latents_tag [1,4,h,w], target [1,4,h,w], ehs [1,77,1024], t)

    `loss_scale`: static scale of d loss / d prediction (activation gradients are 16-bit); removed where the fp32
    parameter gradients are written.  Non-finite gradients skip the update (GradScaler semantics, optim.AdamW)."""

    def __init__(self, unet: TrainableUNet, lr=1e-5, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8, max_grad_norm=1.0,
                 loss_scale=1024.0, process_group=None, bucket_bytes=64 << 20):
        self.unet, self.store = unet, unet.store
        self.max_grad_norm, self.loss_scale = max_grad_norm, float(loss_scale)
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                         and torch.distributed.get_world_size() > 1):
            self.store.reducer = GradReducer(self.store, process_group, bucket_bytes)
            self.world = self.store.reducer.world
        self.store.grad_scale = 1.0 / (self.loss_scale * self.world)
        self.opt = AdamW(unet.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                         half_copies=[p.h for p in self.store.params])
        self.last_norm = None
        self._graph = None

    def enable_cuda_graph(self, latents_ref, latents_tag, target, ehs, timestep):
        """Capture forward + loss + backward (and, separately, the operand refresh) for these shapes into CUDA graphs:
        the ~4800 launches of a step become two graph launches plus the three optimizer launches (the AdamW step number
        is a kernel argument, so the optimizer stays eager).  Later `step()` calls with the same shapes copy their inputs
        into the captured buffers and replay; other shapes run eagerly."""
        dev = self.unet.device
        self._static = [t.detach().to(dev).clone() for t in (latents_ref, latents_tag, target, ehs)]
        self._static_t = float(timestep)
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up on the side stream: learns the contribution counts, sizes the workspaces
                self.forward_backward(*self._static, self._static_t)
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._static_loss = self.forward_backward(*self._static, self._static_t)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            self.unet.refresh_operands()
        self._graph, self._refresh_graph = g, g2

    def _graph_matches(self, tensors, timestep):
        return (self._graph is not None and float(timestep) == self._static_t and
                all(tuple(a.shape) == tuple(b.shape) for a, b in zip(tensors, self._static)))

    def forward_backward(self, latents_ref, latents_tag, target, ehs, timestep):
        u = self.unet
        k = latents_ref.shape[0] // latents_tag.shape[0]
        self.store.begin_step()
        u.clear_attn_bank()
        with torch.enable_grad():
            tb = u.time_biases(timestep)
            u(latents_ref, timestep, ehs if ehs.shape[0] == 1 else ehs.repeat_interleave(k, 0), is_target=False, temb_biases=tb)
            pred = u(latents_tag, timestep, ehs, is_target=True, temb_biases=tb)
        u.clear_attn_bank()
        loss, dpred = mse_loss(pred.detach(), target.float().contiguous(), upstream=self.loss_scale)
        pred.backward(dpred)
        self.store.end_step()
        return loss

    def step(self, latents_ref, latents_tag, target, ehs, timestep, skip_nonfinite=False):
        tensors = (latents_ref, latents_tag, target, ehs)
        graphed = not torch.is_tensor(timestep) and self._graph_matches(tensors, timestep)
        if graphed:
            for dst, src in zip(self._static, tensors):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
            loss = self._static_loss
        else:
            loss = self.forward_backward(latents_ref, latents_tag, target, ehs, timestep)
        if self.max_grad_norm is not None:
            self.last_norm = self.opt.clip_grad_norm_(self.max_grad_norm)
        self.opt.step(skip_nonfinite=skip_nonfinite)
        if graphed:
            self._refresh_graph.replay()
        else:
            self.unet.refresh_operands()
        return loss
