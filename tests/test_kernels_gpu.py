"""Kernel-level GPU parity (round 2): every test calls one C-ABI kernel through `diffews_b200.ops` and compares it with the
oracle (`oracle/metric.py`, bit-exact) or with the same operator in torch fp32 on the CPU (floating point, tolerance in
the test).  VERDICT r01 "What's weak" 1a / 1e.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


# ---------------------------------------------------------------------------------------------------------------------
# a11 / a12: dfw_rthres_iou_hist, bit-exact against oracle.metric (main_oss.py:128-134 + evaluation.py:12-39)
# ---------------------------------------------------------------------------------------------------------------------
def _check_rthres(pred, gt, ign):
    from diffews_b200 import ops
    from oracle.metric import classify_prediction, rthres_mask
    B = pred.shape[0]
    inter, union, mask = ops.rthres_iou_hist(pred.cuda(), gt.cuda(), None if ign is None else ign.cuda(), 0.25)
    torch.cuda.synchronize()
    for b in range(B):
        m = rthres_mask(pred[b:b + 1], 0.25)                               # the reference's fp32 CPU expression
        batch = {"query_mask": gt[b:b + 1].float()}
        if ign is not None:
            batch["query_ignore_idx"] = ign[b:b + 1].float()
        ai, au = classify_prediction(m.clone(), batch)
        assert torch.equal(inter[b].cpu(), ai[:, 0].long()), (b, inter[b].tolist(), ai[:, 0].tolist())
        assert torch.equal(union[b].cpu(), au[:, 0].long()), (b, union[b].tolist(), au[:, 0].tolist())
        if ign is None:
            assert torch.equal(mask[b].cpu().float(), m[0]), f"episode {b}: mask differs from the oracle"
        else:   # ignored pixels carry 255 in the kernel's mask (evaluation.py:20 pred[gt==255] = 255)
            keep = ign[b] == 0
            assert torch.equal(mask[b].cpu().float()[keep], m[0][keep])
            assert bool((mask[b].cpu()[~keep] == 255).all())


@pytest.mark.parametrize("B,H,W", [(3, 64, 96), (2, 512, 512), (2, 768, 768), (2, 10, 10), (3, 7, 9), (1, 1, 5)])
@pytest.mark.parametrize("with_ignore", [False, True])
def test_rthres_kernel_bit_exact(B, H, W, with_ignore):
    """Random uint8 predictions, a //64*64 image (many exact ties of mean vs max/4), an all-zero episode; optional PASCAL
    ignore mask; sizes incl. the config-2 / config-5 images and shapes whose planes are not 16- / 4-byte aligned
    (3*H*W % 16 != 0 for b > 0: the scalar path)."""
    g = torch.Generator().manual_seed(H * 1000 + W)
    pred = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    if B > 1:
        pred[1] = (pred[1] // 64) * 64
    if B > 2:
        pred[2] = 0
    gt = (torch.rand(B, H, W, generator=g) > 0.6).to(torch.uint8)
    ign = ((torch.rand(B, H, W, generator=g) > 0.9) & (gt == 0)).to(torch.uint8) if with_ignore else None
    _check_rthres(pred, gt, ign)


def test_rthres_kernel_golden_tie_table():
    """tests/golden/rthres_cases.json ((R,G,B,max) -> mean(dim=1) > max*0.25 as torch CPU fp32 evaluates it, incl. the
    exact-tie combinations where the integer rule 4(R+G+B) > 3*max is wrong) expanded to images: one episode per
    distinct max, every (R,G,B) of that max as a pixel, the max planted in a spare pixel.  The kernel's mask must equal
    the golden verdict pixel by pixel."""
    from diffews_b200 import ops
    with open(os.path.join(GOLDEN, "rthres_cases.json")) as f:
        cases = json.load(f)["cases"]
    by_max = {}
    for r, g_, b_, mx, want in cases:
        by_max.setdefault(mx, []).append((r, g_, b_, want))
    W = 16
    n_checked = 0
    for mx, rows in sorted(by_max.items()):
        n = len(rows) + 1
        H = (n + W - 1) // W
        pred = torch.zeros(1, 3, H, W, dtype=torch.uint8)
        flat = pred.view(1, 3, H * W)
        for i, (r, g_, b_, _) in enumerate(rows):
            assert max(r, g_, b_) <= mx
            flat[0, 0, i], flat[0, 1, i], flat[0, 2, i] = r, g_, b_
        flat[0, 0, len(rows)] = mx                                # plants the episode max
        gt = torch.zeros(1, H, W, dtype=torch.uint8)
        _, _, mask = ops.rthres_iou_hist(pred.cuda(), gt.cuda(), None, 0.25)
        got = mask.cpu().view(-1)
        for i, (r, g_, b_, want) in enumerate(rows):
            assert int(got[i]) == int(want), f"(R,G,B,max)=({r},{g_},{b_},{mx}): kernel {int(got[i])}, torch CPU fp32 {want}"
            n_checked += 1
        _check_rthres(pred, gt, None)
    assert n_checked == len(cases)


def test_iou_accumulate_matches_index_add():
    """logger.py:35-37 index_add_ by class id == dfw_iou_accumulate (int64, duplicates in one batch included)."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(4)
    B, nclass = 64, 20
    inter = torch.randint(0, 1 << 20, (B, 2), generator=g)
    union = inter + torch.randint(0, 1 << 20, (B, 2), generator=g)
    cls = torch.randint(0, nclass, (B,), generator=g)
    ib = torch.zeros(2, nclass, dtype=torch.int64, device="cuda"); ub = torch.zeros_like(ib)
    ops.iou_accumulate(inter.cuda(), union.cuda(), cls.cuda(), ib, ub)
    ri = torch.zeros(2, nclass, dtype=torch.int64).index_add_(1, cls, inter.t().contiguous())
    ru = torch.zeros(2, nclass, dtype=torch.int64).index_add_(1, cls, union.t().contiguous())
    assert torch.equal(ib.cpu(), ri) and torch.equal(ub.cpu(), ru)


# ---------------------------------------------------------------------------------------------------------------------
# a8 / a6: LayerNorm, GEGLU epilogue, residual epilogue, cross-attention, row softmax vs torch fp32 (CPU)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,C,dt", [(4096, 320, torch.float16), (1024, 640, torch.float16), (300, 1280, torch.float32),
                                    (64, 1280, torch.bfloat16), (77, 64, torch.float16)])
def test_layernorm_kernel(M, C, dt):
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 2.0 + 0.5).to(dt)
    gam = torch.randn(C, generator=g) * 0.3 + 1.0; bet = torch.randn(C, generator=g) * 0.2
    for odt in (torch.float16, torch.bfloat16):
        y = ops.layernorm(x.cuda(), gam.cuda(), bet.cuda(), 1e-5, out_dtype=odt)
        ref = F.layer_norm(x.float(), (C,), gam, bet, 1e-5)
        tol = 1e-3 if odt == torch.float16 else 6e-3            # output rounding: fp16 2^-11, bf16 2^-8 (rms over elements)
        assert rel_l2(y, ref) <= tol, (odt, rel_l2(y, ref))


@pytest.mark.parametrize("M,C", [(4096, 320), (1024, 640), (256, 1280), (70, 1280)])
def test_linear_geglu_epilogue(M, C):
    """ff.net.0.proj + GEGLU (value * gelu_erf(gate)) fused in the GEMM epilogue (rows interleaved per 256-row tile by
    weights.geglu_permute) vs Linear -> chunk -> exact-erf GELU in fp32."""
    from diffews_b200 import ops
    from diffews_b200.weights import geglu_permute
    g = torch.Generator().manual_seed(C)
    x = torch.randn(M, C, generator=g).half()
    w = (torch.randn(8 * C, C, generator=g) * C ** -0.5).half()
    b = torch.randn(8 * C, generator=g) * 0.1
    wp, bp = geglu_permute(w, b)
    y = ops.linear(x.cuda(), wp.cuda(), bp.cuda(), geglu=True)
    h = x.float() @ w.float().t() + b
    val, gate = h.chunk(2, dim=-1)
    ref = val * F.gelu(gate)                                          # diffusers GEGLU: exact (erf) GELU
    assert y.shape == (M, 4 * C)
    assert rel_l2(y, ref) <= 1.5e-3, rel_l2(y, ref)


@pytest.mark.parametrize("M,K,N,res_dt,out_f32", [(4096, 320, 320, torch.float16, False), (1000, 1280, 640, torch.float32, True),
                                                   (256, 5120, 1280, torch.float16, False), (64, 640, 1920, None, False)])
def test_linear_bias_residual_epilogue(M, K, N, res_dt, out_f32):
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(K + N)
    x = torch.randn(M, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half()
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g).to(res_dt) if res_dt is not None else None
    y = ops.linear(x.cuda(), w.cuda(), b.cuda(), residual=None if r is None else r.cuda(), out_f32=out_f32)
    ref = x.float() @ w.float().t() + b
    if r is not None:
        ref = ref + r.float()
    assert rel_l2(y, ref) <= (2e-5 if out_f32 else 6e-4), rel_l2(y, ref)


@pytest.mark.parametrize("B,L,heads,Lctx,shared", [(3, 1024, 5, 2, True), (2, 256, 20, 77, False), (1, 4096, 5, 2, True),
                                                    (2, 64, 10, 128, True)])
def test_cross_attention_kernel(B, L, heads, Lctx, shared):
    """attn2 core: softmax(q k^T / 8) v against Lctx prompt tokens (2 at eval, 77 in training), K/V shared by all
    samples or per sample."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(L + Lctx)
    C = heads * 64
    q = torch.randn(B, L, C, generator=g).half()
    k = torch.randn(1 if shared else B, Lctx, C, generator=g).half()
    v = torch.randn(1 if shared else B, Lctx, C, generator=g).half()
    o = ops.cross_attn(q.cuda(), k.cuda(), v.cuda(), heads, 0.125)
    hd = lambda t, n: t.float().expand(B, -1, -1).reshape(B, n, heads, 64).transpose(1, 2)
    ref = torch.softmax(hd(q, L) @ hd(k, Lctx).transpose(-1, -2) * 0.125, -1) @ hd(v, Lctx)
    ref = ref.transpose(1, 2).reshape(B, L, C)
    assert rel_l2(o, ref) <= 1e-3, rel_l2(o, ref)


@pytest.mark.parametrize("M,L", [(512, 4096), (300, 1024), (64, 9216), (17, 100)])
def test_softmax_rows_kernel(M, L):
    """VAE mid-block attention: row softmax of fp32 logits * scale -> 16-bit probabilities."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(L)
    s = torch.randn(M, L, generator=g) * 30.0
    scale = 512 ** -0.5
    p = ops.softmax_rows(s.cuda(), scale, out_dtype=torch.float16)
    ref = torch.softmax(s * scale, -1)
    assert rel_l2(p, ref) <= 1e-3, rel_l2(p, ref)
    assert float((p.float().sum(-1) - 1).abs().max()) <= 3e-3


@pytest.mark.parametrize("B,h,Lq,Ls,Lb,dt", [(2, 5, 4096, 4096, 4096, torch.float16),     # config-2 level-0 shape
                                             (1, 10, 1024, 1024, 5120, torch.float16),    # config-3 (5-shot) level 1
                                             (1, 20, 144, 144, 144, torch.float16),       # config-5 (768^2) coarsest: ragged
                                             (2, 5, 576, 576, 576, torch.float16),
                                             (1, 20, 64, 64, 0, torch.float16),           # support pass, 8x8 level
                                             (1, 5, 2304, 2304, 2304, torch.bfloat16),
                                             (3, 2, 200, 200, 333, torch.float16)])       # ragged everywhere
@pytest.mark.parametrize("variant", ["v3", "v4"])
def test_attention_forward_kernel(B, h, Lq, Ls, Lb, dt, variant):
    """dfw_attn_kvfused_fwd vs softmax(q [k_self; k_bank]^T / 8) [v_self; v_bank] in fp32 (the reference's concat
    attention, attention_processor.py:251-271), q/k/v as strided column slices of a fused QKV buffer.  v3 = the default
    kernel; v4 = the 96-key / two-issuer variant (DFW_OPT_ATTN_V4, off by default)."""
    from diffews_b200 import _lib, ops
    old = ops.set_option(_lib.OPT_ATTN_V4, int(variant == "v4"))
    try:
        _attention_forward_case(B, h, Lq, Ls, Lb, dt)
    finally:
        ops.set_option(_lib.OPT_ATTN_V4, old)


def _attention_forward_case(B, h, Lq, Ls, Lb, dt):
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(Lq + Lb)
    C = h * 64
    qkv = torch.randn(B, Ls, 3 * C, generator=g).to(dt)
    bank = torch.randn(B, max(Lb, 1), 3 * C, generator=g).to(dt)
    qc, bc = qkv.cuda(), bank.cuda()
    q, ks, vs = qc[..., :C], qc[..., C:2 * C], qc[..., 2 * C:]
    kb, vb = (bc[..., C:2 * C], bc[..., 2 * C:]) if Lb else (None, None)
    o = ops.attn_kvfused(q[:, :Lq], ks, vs, kb, vb, h, 0.125)
    K = qkv[..., C:2 * C].float(); V = qkv[..., 2 * C:].float()
    if Lb:
        K = torch.cat([K, bank[..., C:2 * C].float()], 1); V = torch.cat([V, bank[..., 2 * C:].float()], 1)
    hd = lambda t: t.view(B, -1, h, 64).transpose(1, 2)
    ref = torch.softmax(hd(qkv[:, :Lq, :C].float()) @ hd(K).transpose(-1, -2) * 0.125, -1) @ hd(V)
    ref = ref.transpose(1, 2).reshape(B, Lq, C)
    e = rel_l2(o, ref)
    print(f"attn B{B} h{h} Lq{Lq} Ls{Ls} Lb{Lb} {dt}: rel-L2 {e:.2e}")
    assert e <= (2e-2 if dt == torch.bfloat16 else 3e-3), e


def test_attention_forward_large_logits():
    """Peaked softmax (logits up to +-60 in the log2 domain, what trained weights produce): exercises the lazy O rescale
    and the clamped polynomial exponential far from the row maximum."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, h, L = 1, 3, 512
    C = h * 64
    q = (torch.randn(B, L, C, generator=g) * 3.0).half()
    k = (torch.randn(B, L, C, generator=g) * 3.0).half()
    v = torch.randn(B, L, C, generator=g).half()
    kb = (torch.randn(B, 2 * L, C, generator=g) * 3.0).half()
    vb = torch.randn(B, 2 * L, C, generator=g).half()
    o = ops.attn_kvfused(q.cuda(), k.cuda(), v.cuda(), kb.cuda(), vb.cuda(), h, 0.125)
    hd = lambda t: t.float().view(B, -1, h, 64).transpose(1, 2)
    K = torch.cat([k, kb], 1); V = torch.cat([v, vb], 1)
    ref = (torch.softmax(hd(q) @ hd(K).transpose(-1, -2) * 0.125, -1) @ hd(V)).transpose(1, 2).reshape(B, L, C)
    assert rel_l2(o, ref) <= 3e-3, rel_l2(o, ref)


# ---------------------------------------------------------------------------------------------------------------------
# a9: dfw_seg_head_u8 -- GroupNorm-apply + SiLU + conv 128 -> 3 + clip + *0.5+0.5 + *255 + uint8 in one pass
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,W", [(2, 64, 64), (1, 96, 160), (3, 40, 48), (1, 37, 80), (16, 8, 16), (2, 512, 512), (1, 768, 768)])
def test_seg_head_kernel(N, H, W):
    """The decoder tail of pipeline:887-905 / :787-795 / :534 against the same operators in torch fp32 on the CPU
    (F.group_norm -> F.silu -> F.conv2d -> clip -> *0.5+0.5 -> *255 -> clip(0,255) -> uint8 truncation).  The input is
    produced by a convolution that emits its GroupNorm statistics, exactly as the decoder's last resnet does.  Float image
    within 2e-3 of the range (fp16 operands, fp32 accumulate); uint8 image equal except where the float value sits within
    truncation distance of an integer (<= 1 level, on <= 3 % of the values)."""
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    g = torch.Generator().manual_seed(H * 7 + W)
    x0 = torch.randn(N, H, W, 64, generator=g).half()
    w0 = (torch.randn(128, 64, 1, 1, generator=g) * 0.2).half()
    b0 = torch.randn(128, generator=g) * 0.5
    x = ops.conv2d(x0.cuda(), conv_weight_to_gemm(w0).cuda().half(), b0.cuda(), ksize=1, gn_stats=True)
    if getattr(x, "_gn_partial", None) is None:
        pytest.skip("producer conv cannot emit GroupNorm statistics for this shape (the pipeline then uses the 3-launch path)")
    gam = torch.randn(128, generator=g) * 0.3 + 1.0; bet = torch.randn(128, generator=g) * 0.2
    w = torch.randn(3, 128, 3, 3, generator=g) * 0.03
    b = torch.randn(3, generator=g) * 0.1
    assert ops.seg_head_supported(x)
    wb = ops.seg_head_prepare(w, torch.float16, "cuda")
    f, u = ops.seg_head_u8(x, gam.cuda(), bet.cuda(), 1e-6, wb, b.contiguous(), want_f32=True, want_u8=True)
    xr = x.float().cpu().permute(0, 3, 1, 2)
    a = F.silu(F.group_norm(xr, 32, gam, bet, 1e-6))
    ref = F.conv2d(a, w.half().float(), b, padding=1).clip(-1, 1)
    ref = (ref * 0.5 + 0.5) * 255
    err = (f.cpu() - ref).abs().max().item()
    print(f"seg_head N{N} {H}x{W}: max |f32 - ref| = {err:.3f} grey levels")
    assert err <= 0.5, err                                   # 2e-3 of the 255 range
    ru8 = ref.clip(0, 255).to(torch.uint8)
    d = (u.cpu().int() - ru8.int()).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) <= 3e-2, (int(d.max()), float((d > 0).float().mean()))
    assert torch.equal(u.cpu(), f.cpu().clip(0, 255).to(torch.uint8))     # the kernel's own two outputs are consistent


def test_seg_head_equals_three_launch_path():
    """Full-width decoder: the fused head and the gn-apply + conv + seg_post path give the same uint8 image up to fp16
    rounding of the normalised tensor the unfused path materialises."""
    from diffews_b200 import _lib, ops
    from diffews_b200.vae import AutoencoderKL
    from oracle.sd21 import build_models
    _, vae_o = build_models(0)
    eng = AutoencoderKL.from_module(vae_o)
    z = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(2)).cuda()
    f1, u1 = eng.decode_seg(z, in_scale=1.0 / 0.18215)
    old = ops.set_option(_lib.OPT_SEG_HEAD, 0)
    try:
        f0, u0 = eng.decode_seg(z, in_scale=1.0 / 0.18215)
    finally:
        ops.set_option(_lib.OPT_SEG_HEAD, old)
    d = (u1.int() - u0.int()).abs()
    print("fused head vs 3-launch path: max float diff", float((f1 - f0).abs().max()), "uint8 mismatches", float((d > 0).float().mean()))
    assert float((f1 - f0).abs().max()) <= 1.0 and int(d.max()) <= 1 and float((d > 0).float().mean()) <= 2e-2


@pytest.mark.parametrize("M,K,N,res", [(65536, 320, 320, True), (40000, 320, 960, False), (65536, 512, 512, True)])
def test_linear_weight_stationary_mode_is_bit_identical(M, K, N, res):
    """DFW_OPT_B_RESIDENT (weights of a Cout tile resident in smem, Cout-tile-major unit order) computes the same bits as the
    default ring mainloop, and both match x @ w^T + b in fp32."""
    from diffews_b200 import _lib, ops
    g = torch.Generator().manual_seed(M + N)
    x = torch.randn(M, K, generator=g).half().cuda()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half().cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).half().cuda() if res else None
    y0 = ops.linear(x, w, b, residual=r)
    old = ops.set_option(_lib.OPT_B_RESIDENT, 1)
    try:
        y1 = ops.linear(x, w, b, residual=r)
    finally:
        ops.set_option(_lib.OPT_B_RESIDENT, old)
    assert torch.equal(y0, y1)
    ref = x[:4096].float() @ w.float().t() + b + (r[:4096].float() if res else 0.0)
    assert rel_l2(y1[:4096], ref) <= 2e-3
