"""GPU bring-up: every kernel against a torch fp32 reference. Each case runs in its own subprocess with a timeout so a
trap / hang in one kernel cannot take the others (or the box) down.  Usage: python scripts/gpu_bringup.py [case ...]"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _mk(shape, scale=1.0, seed=0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale)


@case
def linear_basic():
    import torch
    from diffews_b200 import ops
    for (M, K, N) in [(128, 64, 128), (300, 320, 320), (4096, 1280, 1280), (1000, 640, 640), (64, 1024, 2560), (513, 320, 960)]:
        x = _mk((M, K), 1.0, 1).bfloat16(); w = _mk((N, K), K ** -0.5, 2).bfloat16(); b = _mk((N,), 1.0, 3)
        y = ops.linear(x, w, b)
        ref = x.float() @ w.float().t() + b
        print(f"linear M{M} K{K} N{N}: rel {rel(y, ref):.3e}")
        assert rel(y, ref) < 6e-3
    x = _mk((300, 320), 1, 1).bfloat16(); w = _mk((320, 320), 320 ** -0.5, 2).bfloat16(); b = _mk((320,), 1, 3)
    r = _mk((300, 320), 1, 4)
    y = ops.linear(x, w, b, residual=r, out_f32=True, out_scale=0.5)
    ref = (x.float() @ w.float().t() + b) * 0.5 + r
    print("linear f32 res/scale:", rel(y, ref)); assert rel(y, ref) < 1e-5 + 1e-3
    y = ops.linear(x, w, b, residual=r.bfloat16(), silu=True)
    ref = torch.nn.functional.silu(x.float() @ w.float().t() + b) + r.bfloat16().float()
    print("linear silu+res bf16:", rel(y, ref)); assert rel(y, ref) < 6e-3


@case
def linear_fp16_operands():
    """The 16-bit format is a per-call choice (bf16 or fp16); A and B must share it (a mix is rejected on the host:
    tcgen05 kind::f16 traps with an illegal instruction when a_format != b_format)."""
    import torch
    from diffews_b200 import ops
    M, K, N = 777, 640, 640
    xf = _mk((M, K), 1.0, 1); wf = _mk((N, K), K ** -0.5, 2); b = _mk((N,), 1.0, 3)
    for dt in (torch.bfloat16, torch.float16):
        x, w = xf.to(dt), wf.to(dt)
        y = ops.linear(x, w, b, out_f32=True)
        ref = x.float() @ w.float().t() + b
        e = rel(y, ref)
        print(f"linear {dt} f32 out: rel {e:.3e}")
        assert e < 1e-4
        r = _mk((M, N), 1, 5).to(dt)
        y = ops.linear(x, w, b, residual=r)
        assert y.dtype == dt
        e = rel(y, ref + r.float()); print(f"linear {dt} 16-bit out + res: rel {e:.3e}")
        assert e < (6e-3 if dt == torch.bfloat16 else 8e-4)
    try:
        ops.linear(xf.bfloat16(), wf.half())
        raise SystemExit("mixed operand formats must be rejected")
    except AssertionError:
        pass
    from diffews_b200.weights import conv_weight_to_gemm
    x = _mk((2, 16, 16, 128), 1, 1).half(); w = _mk((256, 128, 3, 3), (128 * 9) ** -0.5, 2).half(); bb = _mk((256,), 1, 3)
    y = ops.conv2d(x, conv_weight_to_gemm(w), bb, ksize=3, out_f32=True)
    e = rel(y, _conv_ref(x, w, bb)); print("conv fp16:", e); assert e < 1e-3
    g = _mk((128,), 1, 2); be = _mk((128,), 1, 3); xx = _mk((2, 256, 128), 2, 1)
    import torch.nn.functional as F
    for xin in (xx, xx.half(), xx.bfloat16()):
        yn = ops.groupnorm(xin, g, be, eps=1e-6, silu=True, out_dtype=torch.float16)
        refn = F.silu(F.group_norm(xin.float().transpose(1, 2), 32, g, be, 1e-6).transpose(1, 2))
        assert yn.dtype == torch.float16 and rel(yn, refn) < 6e-4, rel(yn, refn)
        yl = ops.layernorm(xin, g, be, out_dtype=torch.float16)
        assert yl.dtype == torch.float16 and rel(yl, F.layer_norm(xin.float(), (128,), g, be, 1e-5)) < 6e-4
    print("gn/ln fp16 ok")
    # fp16 attention + cross attention
    B, h, L = 2, 5, 256; C = 320
    q = _mk((B, L, C), 1.5, 1).half(); k = _mk((B, L, C), 1.5, 2).half(); v = _mk((B, L, C), 1, 3).half()
    o = ops.attn_kvfused(q, k, v, k, v, h, 0.125)
    ref = _attn_ref(q, torch.cat([k, k], 1), torch.cat([v, v], 1), h, 0.125)
    assert o.dtype == torch.float16
    e = rel(o, ref); print("attn fp16:", e); assert e < 1.5e-3
    kk = _mk((1, 2, C), 1, 2).half(); vv = _mk((1, 2, C), 1, 3).half()
    o = ops.cross_attn(q, kk, vv, h, 0.125)
    e = rel(o, _attn_ref(q, kk.expand(B, -1, -1), vv.expand(B, -1, -1), h, 0.125)); print("cross fp16:", e); assert e < 1e-3
    xs = _mk((2, 4, 4, 64), 1, 1)
    assert torch.equal(ops.upsample2x(xs, torch.float16).float(),
                       F.interpolate(xs.half().float().permute(0, 3, 1, 2), scale_factor=2.0).permute(0, 2, 3, 1))
    assert torch.equal(ops.cast16(xs, torch.float16), xs.half())
    s_ = _mk((64, 256), 10, 1)
    e = rel(ops.softmax_rows(s_, 0.1, out_dtype=torch.float16), torch.softmax(s_ * 0.1, -1)); assert e < 6e-4, e
    print("fp16 misc ok")


@case
def linear_geglu():
    import torch
    from diffews_b200 import ops
    from diffews_b200.weights import geglu_permute
    M, C = 500, 320
    x = _mk((M, C), 1, 1).bfloat16(); w = _mk((8 * C, C), C ** -0.5, 2).bfloat16(); b = _mk((8 * C,), 0.5, 3)
    wp, bp = geglu_permute(w, b)
    y = ops.linear(x, wp, bp, geglu=True)
    h = x.float() @ w.float().t() + b
    ref = h[:, :4 * C] * torch.nn.functional.gelu(h[:, 4 * C:])
    print("geglu:", rel(y, ref)); assert rel(y, ref) < 6e-3


def _conv_ref(x_nhwc, w_oihw, b, stride=1, pad=1, asym=False):
    import torch
    import torch.nn.functional as F
    x = x_nhwc.float().permute(0, 3, 1, 2)
    if asym:
        x = F.pad(x, (0, 1, 0, 1)); pad = 0
    y = F.conv2d(x, w_oihw.float(), b, stride=stride, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


@case
def conv_s1():
    import torch
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    for (N, H, W, Ci, Co, ks) in [(2, 16, 16, 128, 128, 3), (3, 8, 8, 64, 320, 3), (1, 64, 64, 320, 320, 3),
                                  (2, 32, 32, 640, 1280, 1), (1, 24, 40, 128, 256, 3), (2, 12, 12, 64, 128, 3),
                                  (1, 128, 128, 128, 128, 3), (5, 8, 8, 1280, 1280, 3),
                                  (4, 256, 256, 128, 128, 3), (1, 593, 128, 64, 128, 3), (3, 200, 128, 128, 64, 1)]:
        x = _mk((N, H, W, Ci), 1, 1).bfloat16(); w = _mk((Co, Ci, ks, ks), (Ci * ks * ks) ** -0.5, 2).bfloat16()
        b = _mk((Co,), 1, 3)
        y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=ks)
        ref = _conv_ref(x, w, b, 1, (ks - 1) // 2)
        e = rel(y, ref)
        print(f"conv N{N} {H}x{W} {Ci}->{Co} k{ks}: rel {e:.3e}")
        assert e < 6e-3
    # per-sample bias + residual
    N, H, W, Ci, Co = 3, 16, 16, 128, 320
    x = _mk((N, H, W, Ci), 1, 1).bfloat16(); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).bfloat16()
    b = _mk((N, Co), 1, 3); r = _mk((N, H, W, Co), 1, 4).bfloat16()
    y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, residual=r, bias_per_sample=True)
    ref = _conv_ref(x, w, None) + b[:, None, None, :] + r.float()
    print("conv per-sample bias + res:", rel(y, ref)); assert rel(y, ref) < 6e-3
    # paired-tile path (single Cout tile, many pixel tiles) with residual, 16-bit and fp32
    N, H, W, Ci, Co = 2, 300, 256, 128, 128
    x = _mk((N, H, W, Ci), 1, 1).bfloat16(); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).bfloat16(); b = _mk((Co,), 1, 3)
    r = _mk((N, H, W, Co), 1, 4)
    y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, residual=r.bfloat16())
    e = rel(y, _conv_ref(x, w, b) + r.bfloat16().float()); print("paired conv + bf16 res:", e); assert e < 6e-3
    y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, residual=r, out_f32=True)
    e = rel(y, _conv_ref(x, w, b) + r); print("paired conv + f32 res:", e); assert e < 1e-3


@case
def conv_s2():
    import torch
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    for (N, H, W, Ci, Co, pm) in [(2, 16, 16, 128, 128, 0), (2, 16, 16, 128, 128, 1), (1, 64, 64, 320, 320, 0),
                                  (1, 32, 48, 64, 256, 1)]:
        x = _mk((N, H, W, Ci), 1, 1).bfloat16(); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).bfloat16()
        b = _mk((Co,), 1, 3)
        y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, stride=2, pad_mode=pm)
        ref = _conv_ref(x, w, b, 2, 1, asym=(pm == 1))
        e = rel(y, ref)
        print(f"conv s2 N{N} {H}x{W} {Ci}->{Co} pad_mode{pm}: rel {e:.3e}")
        assert e < 6e-3


@case
def conv_small_cout():
    import torch
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    for (N, H, W, Ci, Co) in [(2, 32, 32, 128, 3), (2, 16, 16, 320, 4), (1, 16, 16, 512, 8)]:
        x = _mk((N, H, W, Ci), 1, 1).bfloat16(); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).bfloat16()
        b = _mk((Co,), 1, 3)
        y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, out_f32=True)
        ref = _conv_ref(x, w, b)
        e = rel(y, ref)
        print(f"conv small cout {Ci}->{Co}: rel {e:.3e}")
        assert e < 1e-3


def _attn_ref(q, k, v, heads, scale):
    import torch
    B, Lq, C = q.shape
    qh = q.float().view(B, Lq, heads, 64).transpose(1, 2)
    kh = k.float().view(B, -1, heads, 64).transpose(1, 2)
    vh = v.float().view(B, -1, heads, 64).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    o = torch.softmax(s, dim=-1) @ vh
    return o.transpose(1, 2).reshape(B, Lq, C)


@case
def attn():
    import torch
    from diffews_b200 import ops
    for (B, h, Lq, Ls, Lb) in [(1, 1, 128, 128, 0), (2, 5, 256, 256, 256), (1, 20, 64, 64, 128), (2, 10, 1024, 1024, 2048),
                               (1, 5, 144, 144, 288), (1, 5, 4096, 4096, 4096)]:
        C = h * 64
        q = _mk((B, Lq, C), 1.5, 1).bfloat16(); ks = _mk((B, Ls, C), 1.5, 2).bfloat16(); vs = _mk((B, Ls, C), 1, 3).bfloat16()
        kb = _mk((B, Lb, C), 1.5, 4).bfloat16() if Lb else None; vb = _mk((B, Lb, C), 1, 5).bfloat16() if Lb else None
        o = ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125)
        kk = torch.cat([ks, kb], 1) if Lb else ks; vv = torch.cat([vs, vb], 1) if Lb else vs
        ref = _attn_ref(q, kk, vv, h, 0.125)
        e = rel(o, ref)
        print(f"attn B{B} h{h} Lq{Lq} Ls{Ls} Lb{Lb}: rel {e:.3e}")
        assert e < 1e-2
    # strided (fused qkv buffer) inputs
    B, h, L = 2, 5, 256; C = 320
    qkv = _mk((B, L, 3 * C), 1.2, 7).bfloat16()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    o = ops.attn_kvfused(q, k, v, k, v, h, 0.125)
    ref = _attn_ref(q, torch.cat([k, k], 1), torch.cat([v, v], 1), h, 0.125)
    print("attn strided:", rel(o, ref)); assert rel(o, ref) < 1e-2


@case
def cross_attn():
    import torch
    from diffews_b200 import ops
    B, L, h = 3, 200, 5; C = 320
    for Lctx, Bk in [(2, 1), (2, 3), (77, 3)]:
        q = _mk((B, L, C), 1, 1).bfloat16(); k = _mk((Bk, Lctx, C), 1, 2).bfloat16(); v = _mk((Bk, Lctx, C), 1, 3).bfloat16()
        o = ops.cross_attn(q, k, v, h, 0.125)
        ref = _attn_ref(q, k.expand(B, -1, -1), v.expand(B, -1, -1), h, 0.125)
        print(f"cross Lctx{Lctx} Bk{Bk}:", rel(o, ref)); assert rel(o, ref) < 6e-3


@case
def norms():
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    for (N, HW, C) in [(2, 64, 1280), (3, 4096, 320), (1, 1024, 960), (2, 256, 2560), (1, 16384, 128), (2, 256, 1920)]:
        x = (_mk((N, HW, C), 2, 1) + 0.5)
        g = _mk((C,), 1, 2); b = _mk((C,), 1, 3)
        for xin in (x.bfloat16(), x):
            for silu in (False, True):
                y = ops.groupnorm(xin, g, b, eps=1e-5, silu=silu)
                ref = F.group_norm(xin.float().transpose(1, 2), 32, g, b, 1e-5).transpose(1, 2)
                if silu: ref = F.silu(ref)
                e = rel(y, ref)
                assert e < 5e-3, (N, HW, C, xin.dtype, silu, e)
        print(f"gn N{N} HW{HW} C{C}: ok ({e:.2e})")
    for (M, C) in [(1000, 320), (77, 640), (4096, 1280)]:
        x = _mk((M, C), 2, 1) + 0.3; g = _mk((C,), 1, 2); b = _mk((C,), 1, 3)
        for xin in (x.bfloat16(), x):
            y = ops.layernorm(xin, g, b)
            ref = F.layer_norm(xin.float(), (C,), g, b, 1e-5)
            e = rel(y, ref); assert e < 5e-3, e
        print(f"ln M{M} C{C}: ok ({e:.2e})")
    s = _mk((300, 4096), 20, 1)
    p = ops.softmax_rows(s, 0.044)
    ref = torch.softmax(s * 0.044, -1)
    print("softmax:", rel(p, ref)); assert rel(p, ref) < 5e-3


@case
def misc():
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    x = _mk((2, 5, 7, 64), 1, 1).bfloat16()
    y = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref); print("upsample ok")
    a = _mk((3, 10, 320), 1, 1).bfloat16(); b = _mk((3, 10, 640), 1, 2).bfloat16()
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], -1)); print("concat ok")
    for (N, Ci, H, W, Co) in [(2, 3, 40, 72, 128), (2, 4, 16, 16, 320), (1, 8, 64, 64, 320), (2, 4, 24, 24, 512)]:
        x = _mk((N, Ci, H, W), 1, 1); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2); bb = _mk((Co,), 1, 3)
        y = ops.conv3x3_small_cin(x, w.permute(0, 2, 3, 1).contiguous(), bb)
        ref = F.conv2d(x, w, bb, padding=1).permute(0, 2, 3, 1)
        e = rel(y, ref); print(f"small cin {Ci}->{Co}: {e:.2e}"); assert e < 4e-3
    # pointwise: NCHW -> NCHW 4->4 with scales
    x = _mk((2, 4, 8, 8), 1, 1); w = torch.randn(4, 4); b = torch.randn(4)
    y = torch.empty_like(x)
    ops.pointwise_small(x, (4 * 64, 1, 64), w, b, y, (4 * 64, 1, 64), 2, 64, in_scale=-1 / 0.18215, out_scale=1.0)
    ref = F.conv2d(x.cpu() * (-1 / 0.18215), w[:, :, None, None], b)      # CPU fp32 (GPU conv2d defaults to TF32)
    print("pointwise:", rel(y.cpu(), ref)); assert rel(y.cpu(), ref) < 1e-5
    # seg_post
    dec = _mk((2, 64, 16), 1.5, 3)
    f, u = ops.seg_post(dec, 8, 8)
    ref = ((dec[..., :3].clip(-1, 1) * 0.5 + 0.5) * 255).permute(0, 2, 1).reshape(2, 3, 8, 8)
    assert torch.equal(f, ref), (f - ref).abs().max()
    assert torch.equal(u, ref.clip(0, 255).to(torch.uint8)); print("seg_post ok")
    t = ops.nhwc_f32_to_nchw(dec, 4, 8, 8, scale=2.0)
    assert torch.equal(t, (dec[..., :4] * 2.0).permute(0, 2, 1).reshape(2, 4, 8, 8)); print("nhwc->nchw ok")


@case
def bmm_and_im2col():
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    B, M, K, N = 3, 512, 256, 384
    x = _mk((B, M, K), 1, 1).half(); w = _mk((B, N, K), K ** -0.5, 2).half(); b = _mk((N,), 1, 3)
    y = ops.bmm_nt(x, w, b, out_f32=True)
    ref = torch.bmm(x.float(), w.float().transpose(1, 2)) + b
    e = rel(y, ref); print("bmm_nt:", e); assert e < 1e-4
    # strided per-batch weights: w[b] = big[:, b*N:(b+1)*N]^T slices of a [K2, B*L] buffer (the V^T layout of the VAE attention)
    C, L = 128, 256
    big = _mk((C, B * L), 1, 4).half()                      # [C, B*L]
    vt = big.view(C, B, L).permute(1, 0, 2)                 # [B, C, L], row stride B*L
    p = _mk((B, 384, L), 1, 5).half()
    y = ops.bmm_nt(p, vt)
    ref = torch.bmm(p.float(), vt.float().transpose(1, 2))
    e = rel(y, ref); print("bmm_nt strided w:", e); assert e < 1.5e-3
    from diffews_b200.layers import SmallCinConv
    for (Nn, Ci, H, W, Co) in [(2, 3, 40, 72, 128), (2, 4, 16, 16, 320), (1, 8, 64, 64, 320), (2, 4, 24, 24, 512)]:
        conv = torch.nn.Conv2d(Ci, Co, 3, padding=1)
        xx = _mk((Nn, Ci, H, W), 1, 1)
        layer = SmallCinConv({"c.weight": conv.weight.data, "c.bias": conv.bias.data}, "c", "cuda", torch.float16)
        y = layer(xx, out_f32=True)
        ref = F.conv2d(xx.cpu(), conv.weight.data, conv.bias.data, padding=1).permute(0, 2, 3, 1)
        e = rel(y.cpu(), ref); print(f"im2col conv {Ci}->{Co}: {e:.2e}"); assert e < 1.5e-3


@case
def upconv():
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    from diffews_b200.weights import upconv_phase_weights
    for (N, H, W, Ci, Co, dt, f32) in [(2, 16, 16, 128, 128, torch.float16, False), (1, 24, 40, 256, 256, torch.bfloat16, False),
                                       (3, 8, 8, 1280, 1280, torch.float16, True), (2, 64, 64, 512, 512, torch.float16, False)]:
        x = _mk((N, H, W, Ci), 1, 1).to(dt); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).to(dt); b = _mk((Co,), 1, 3)
        y = ops.upconv2x(x, upconv_phase_weights(w).cuda().to(dt), b, out_f32=f32)
        ref = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest"), w.float(), b,
                       padding=1).permute(0, 2, 3, 1)
        e = rel(y, ref); print(f"upconv N{N} {H}x{W} {Ci}->{Co} {dt}: rel {e:.3e}")
        assert y.shape == ref.shape and e < (6e-3 if dt == torch.bfloat16 else 1.5e-3)


@case
def gn_fused_stats():
    """conv epilogue emits the GroupNorm statistics of its output; groupnorm() then skips its statistics pass."""
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm, upconv_phase_weights
    for (N, H, W, Ci, Co, res, f32) in [(3, 32, 32, 128, 128, False, False), (2, 40, 24, 128, 256, True, False),
                                         (2, 16, 16, 256, 512, True, True), (16, 64, 64, 128, 128, True, False),
                                         (5, 300, 256, 64, 128, False, False)]:
        x = _mk((N, H, W, Ci), 1, 1).half(); w = _mk((Co, Ci, 3, 3), (Ci * 9) ** -0.5, 2).half(); b = _mk((Co,), 1, 3)
        r = (_mk((N, H, W, Co), 1, 4) if f32 else _mk((N, H, W, Co), 1, 4).half()) if res else None
        g = _mk((Co,), 1, 5); be = _mk((Co,), 1, 6)
        y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=3, residual=r, out_f32=f32, gn_stats=True)
        fused = hasattr(y, "_gn_partial")
        yn = ops.groupnorm(y, g, be, eps=1e-6, silu=True, out_dtype=torch.float16)
        y_plain = y.clone()                                  # clone drops the attribute -> two-pass path
        yn_ref_kernel = ops.groupnorm(y_plain, g, be, eps=1e-6, silu=True, out_dtype=torch.float16)
        ref = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, g, be, 1e-6)).permute(0, 2, 3, 1)
        e1, e2 = rel(yn, ref), rel(yn_ref_kernel, ref)
        print(f"gn fused N{N} {H}x{W} {Ci}->{Co} res={res} f32={f32}: fused={fused} rel {e1:.2e} (two-pass {e2:.2e})")
        assert e1 < 1e-3 and e2 < 1e-3
        assert fused == bool(ops.lib.dfw_conv_gnstats_supported(N, H, W, Co))
    x = _mk((2, 32, 32, 256), 1, 1).half(); w = _mk((256, 256, 3, 3), (256 * 9) ** -0.5, 2).half(); b = _mk((256,), 1, 3)
    g = _mk((256,), 1, 5); be = _mk((256,), 1, 6)
    y = ops.upconv2x(x, upconv_phase_weights(w).cuda().half(), b, gn_stats=True)
    assert hasattr(y, "_gn_partial")
    yn = ops.groupnorm(y, g, be, eps=1e-6, silu=False, out_dtype=torch.float16)
    ref = F.group_norm(y.float().permute(0, 3, 1, 2), 32, g, be, 1e-6).permute(0, 2, 3, 1)
    e = rel(yn, ref); print("gn fused after upconv:", e); assert e < 1e-3


@case
def conv_t128():
    """Cout = 128 3x3 stride-1 layers on large images run the channel-major (T128) kernel: bias / residual / fused
    GroupNorm statistics / fp16 and bf16, tile grids that are not powers of two."""
    import torch
    import torch.nn.functional as F
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    for (N, H, W, Ci, res, dt, ks, Co) in [
            (3, 256, 256, 128, True, torch.float16, 3, 128), (4, 128, 256, 256, False, torch.float16, 3, 128),
            (5, 272, 144, 64, True, torch.bfloat16, 3, 128), (16, 112, 96, 128, False, torch.float16, 3, 128),
            (4, 256, 256, 64, False, torch.float16, 1, 128), (3, 144, 272, 256, True, torch.float16, 1, 128),
            (3, 128, 128, 128, True, torch.float16, 3, 256), (2, 144, 112, 256, False, torch.float16, 3, 512),
            (5, 96, 80, 128, True, torch.float16, 1, 256), (16, 64, 64, 512, True, torch.float16, 3, 512)]:
        x = _mk((N, H, W, Ci), 1, 1).to(dt); w = _mk((Co, Ci, ks, ks), (Ci * ks * ks) ** -0.5, 2).to(dt); b = _mk((Co,), 1, 3)
        r = _mk((N, H, W, Co), 1, 4).to(dt) if res else None
        g = _mk((Co,), 1, 5); be = _mk((Co,), 1, 6)
        y = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=ks, residual=r, gn_stats=True)
        assert hasattr(y, "_gn_partial")
        ref = _conv_ref(x, w, b, 1, (ks - 1) // 2) + (r.float() if res else 0)
        e = rel(y, ref)
        yn = ops.groupnorm(y, g, be, eps=1e-6, silu=True, out_dtype=dt)
        refn = F.silu(F.group_norm(y.float().permute(0, 3, 1, 2), 32, g, be, 1e-6)).permute(0, 2, 3, 1)
        e2 = rel(yn, refn)
        y2 = ops.conv2d(x, conv_weight_to_gemm(w), b, ksize=ks, residual=r)           # without statistics
        same = torch.equal(y, y2)
        print(f"t128 N{N} {H}x{W} {Ci}->{Co} k{ks} res={res} {dt}: conv rel {e:.2e}  gn(fused stats) rel {e2:.2e}  equal w/o stats {same}")
        assert e < (6e-3 if dt == torch.bfloat16 else 1e-3) and e2 < (8e-3 if dt == torch.bfloat16 else 1e-3) and same


@case
def conv_gn_in():
    """GroupNorm + SiLU applied to the conv operand on the fly == groupnorm kernel followed by the conv (same fp32
    formula, same fp16 rounding of the normalised value, same MMA order -> expected bit-identical)."""
    import torch
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    for (N, H, W, Ci, Co, ks, res, dt) in [(3, 256, 256, 128, 128, 3, True, torch.float16), (16, 64, 64, 512, 512, 3, True, torch.float16),
                                           (4, 144, 176, 128, 256, 3, False, torch.float16), (4, 128, 128, 256, 128, 1, False, torch.float16),
                                           (3, 160, 128, 256, 256, 3, True, torch.bfloat16)]:
        x0 = _mk((N, H, W, 128), 1, 1).to(dt); w0 = _mk((Ci, 128, 3, 3), (128 * 9) ** -0.5, 7).to(dt); b0 = _mk((Ci,), 1, 8)
        x = ops.conv2d(x0, conv_weight_to_gemm(w0), b0, ksize=3, gn_stats=True)         # producer: emits the statistics
        assert hasattr(x, "_gn_partial")
        w = _mk((Co, Ci, ks, ks), (Ci * ks * ks) ** -0.5, 2).to(dt); b = _mk((Co,), 1, 3)
        g = _mk((Ci,), 1, 5) + 1.0; be = _mk((Ci,), 1, 6)
        r = _mk((N, H, W, Co), 1, 4).to(dt) if res else None
        assert ops.conv_gn_in_supported(x, Co, ks)
        y = ops.conv2d_gn_in(x, g, be, 1e-6, conv_weight_to_gemm(w), b, ksize=ks, residual=r, gn_stats=True)
        xn = ops.groupnorm(x, g, be, eps=1e-6, silu=True, out_dtype=dt)
        y_ref = ops.conv2d(xn, conv_weight_to_gemm(w), b, ksize=ks, residual=r, gn_stats=True)
        e = rel(y, y_ref)
        same = torch.equal(y, y_ref)
        p1, p2 = y._gn_partial[0], y_ref._gn_partial[0]
        pe = float((p1 - p2).abs().max() / p2.abs().max())
        print(f"gn-in N{N} {H}x{W} {Ci}->{Co} k{ks} res={res} {dt}: rel {e:.2e} identical {same}  stats diff {pe:.1e}")
        assert e < 1e-5 and pe < 1e-5


@case
def rthres():
    import torch
    from diffews_b200 import ops
    from oracle.metric import rthres_mask, classify_prediction
    g = torch.Generator().manual_seed(0)
    B, H, W = 3, 64, 96
    pred = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    pred[1] = (pred[1] // 64) * 64          # many ties
    pred[2, :, :, :] = 0                     # all-zero episode
    gt = (torch.rand(B, H, W, generator=g) > 0.6).to(torch.uint8)
    ign = ((torch.rand(B, H, W, generator=g) > 0.9) & (gt == 0)).to(torch.uint8)
    for ignore in (None, ign):
        inter, union, mask = ops.rthres_iou_hist(pred.cuda(), gt.cuda(), None if ignore is None else ignore.cuda(), 0.25)
        for b in range(B):
            m = rthres_mask(pred[b:b + 1], 0.25)                      # [1,H,W] float
            batch = {"query_mask": gt[b:b + 1].float()}
            if ignore is not None: batch["query_ignore_idx"] = ignore[b:b + 1].float()
            ai, au = classify_prediction(m.clone(), batch)
            assert torch.equal(inter[b].cpu(), ai[:, 0].long()), (b, inter[b], ai)
            assert torch.equal(union[b].cpu(), au[:, 0].long()), (b, union[b], au)
            if ignore is None:
                assert torch.equal(mask[b].cpu().float(), m[0])
    print("rthres/hist bit-exact ok")


def main():
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1 and os.environ.get("DFW_BRINGUP_CHILD") == "1":
        CASES[names[0]]()
        import torch
        torch.cuda.synchronize()
        print("CASE_OK", names[0])
        return 0
    failed = []
    for n in names:
        t0 = time.time()
        env = dict(os.environ, DFW_BRINGUP_CHILD="1")
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), n], env=env, capture_output=True, text=True,
                               timeout=int(os.environ.get("DFW_CASE_TIMEOUT", "180")))
            out = r.stdout + r.stderr
            ok = r.returncode == 0 and "CASE_OK" in r.stdout
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace") + "\nTIMEOUT"
            ok = False
        print(f"===== {n}: {'PASS' if ok else 'FAIL'} ({time.time() - t0:.1f}s)")
        print(out[-6000:])
        if not ok:
            failed.append(n)
    print("FAILED:", failed)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
