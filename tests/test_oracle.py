"""CPU tests that pin the ORACLE (tests/ may import oracle/; the product never does).

The reference ships no tests / golden vectors for this path, so the pins are of two kinds:
  (1) against the reference's OWN code run in the build container (scripts/make_golden_*.py execute its files unmodified
      over stand-ins for the un-installable imports; fixtures in tests/golden/): the whole test_diffusion loop, the
      pipeline's __call__ / single_infer, the KV-bank attention processors, the scheduler tables, Evaluator / AverageMeter
      (the tests named *_unmodified_reference_* at the end of this file);
  (2) structural checks that need no reference run:
  - parameter counts of the restated SD-2.1 UNet / VAE equal the published totals (state-dict compatibility),
  - the k-shot bank fold of attention_processor.py:253-267 restated literally == the oracle's shot-major concat,
  - the DDIM step with scheduler_1.0_1.0/scheduler_config.json == exact negation,
  - rthres fp32 tie semantics (golden table produced by torch CPU, the very code path main_oss.py:128-134 executes),
  - torch.histc bins=2 drops the 255 ignore value (evaluation.py:16-33),
  - committed golden fixtures of the oracle's own outputs (drift guard).
"""
import json
import os

import pytest
import torch

from oracle import metric as om
from oracle import pipeline as op
from oracle import sd21

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.timeout(300)
def test_parameter_counts_match_sd21():
    unet, vae = sd21.build_models(0)
    n_ref = sum(p.numel() for p in unet.conv_in_ref.parameters())
    assert sum(p.numel() for p in unet.parameters()) - n_ref == 865_910_724
    assert n_ref == 23_360
    assert sum(p.numel() for p in vae.parameters()) == 83_653_863
    assert sum(p.numel() for p in vae.encoder.parameters()) == 34_163_592
    assert sum(p.numel() for p in vae.decoder.parameters()) == 49_490_179
    assert len(unet.bank_attentions()) == 16
    keys = unet.state_dict().keys()
    for k in ["conv_in_ref.weight", "down_blocks.0.attentions.1.transformer_blocks.0.attn1.to_q.weight",
              "down_blocks.2.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_in.weight",
              "up_blocks.3.attentions.2.transformer_blocks.0.ff.net.0.proj.weight", "up_blocks.2.upsamplers.0.conv.bias",
              "up_blocks.1.resnets.2.conv_shortcut.weight", "conv_norm_out.weight", "time_embedding.linear_2.bias"]:
        assert k in keys, k
    assert unet.up_blocks[1].resnets[2].conv1.in_channels == 1920
    assert unet.up_blocks[3].resnets[0].conv1.in_channels == 960
    assert torch.equal(unet.conv_in_ref.weight, unet.conv_in.weight.repeat(1, 2, 1, 1) / 2)


def _head_to_batch_dim(t, heads):      # diffusers Attention.head_to_batch_dim (upstream)
    B, S, C = t.shape
    return t.reshape(B, S, heads, C // heads).permute(0, 2, 1, 3).reshape(B * heads, S, C // heads)


def _batch_to_head_dim(t, heads):      # diffusers Attention.batch_to_head_dim (upstream)
    Bh, S, d = t.shape
    return t.reshape(Bh // heads, heads, S, d).permute(0, 2, 1, 3).reshape(Bh // heads, S, heads * d)


@pytest.mark.parametrize("B,k,S,heads", [(1, 1, 16, 5), (2, 5, 8, 5), (3, 7, 4, 10)])
def test_kshot_fold_is_shot_major_concat(B, k, S, heads):
    """attention_processor.py:253-267 restated literally vs the oracle's `bank.reshape(B, -1, C)`."""
    C = heads * 64
    g = torch.Generator().manual_seed(0)
    key_support = torch.randn(B * k, S, C, generator=g)              # to_k output of the support pass
    bank = _head_to_batch_dim(key_support, heads)                    # what the reference stores (:247-252)
    folded = _batch_to_head_dim(bank, heads)                         # :256
    folded = _head_to_batch_dim(folded.view(B, -1, folded.shape[-1]), heads)   # :257  -> [B*h, k*S, d]
    mine = _head_to_batch_dim(key_support.reshape(B, -1, C), heads)
    assert torch.equal(folded, mine)
    for b in range(B):                                               # shot-major order of one episode / head
        for s in range(k):
            assert torch.equal(mine[b * heads + 1, s * S:(s + 1) * S], key_support[b * k + s, :, 64:128])


def test_bank_attention_matches_explicit_concat():
    torch.manual_seed(0)
    att = sd21.Attention(128, None, heads=2, dim_head=64, bank=True)
    B, k, S = 2, 3, 8
    sup, qry = torch.randn(B * k, S, 128), torch.randn(B, S, 128)
    att.clear_bank()
    att(sup)
    out = att(qry)
    ks, vs = att.to_k(sup), att.to_v(sup)
    for b in range(B):
        kk = torch.cat([att.to_k(qry[b:b + 1]), ks[b * k:(b + 1) * k].reshape(1, k * S, 128)], 1)
        vv = torch.cat([att.to_v(qry[b:b + 1]), vs[b * k:(b + 1) * k].reshape(1, k * S, 128)], 1)
        q = att.to_q(qry[b:b + 1])
        o = torch.zeros(1, S, 128)
        for h in range(2):
            sl = slice(h * 64, (h + 1) * 64)
            p = torch.softmax(q[..., sl] @ kk[..., sl].transpose(1, 2) * 0.125, -1)
            o[..., sl] = p @ vv[..., sl]
        assert torch.allclose(att.to_out[0](o), out[b:b + 1], atol=1e-5)
    att.clear_bank()
    assert att.k_bank is None and att.v_bank is None


def test_scheduler_collapses_to_negation():
    """scheduler_1.0_1.0/scheduler_config.json: beta_start = beta_end = 1 -> alphas_cumprod == 0 -> x0 = -v exactly."""
    ac = op.ddim_alphas_cumprod()
    assert float(ac.abs().max()) == 0.0
    g = torch.Generator().manual_seed(1)
    v, x = torch.randn(2, 4, 8, 8, generator=g) * 3, torch.randn(2, 4, 8, 8, generator=g)
    assert torch.equal(op.ddim_step_pred_original(v, 1, x), -v)
    from diffews_b200.scheduler import DDIMSchedulerCustomized
    s = DDIMSchedulerCustomized()
    s.set_timesteps(1)
    assert s.timesteps.tolist() == [1] and s.is_pure_negation(1)
    assert torch.equal(s.step(v, 1, x).pred_original_sample, -v)
    s.set_timesteps(20)
    assert s.timesteps.tolist()[:3] == [951, 901, 851] and s.timesteps.tolist()[-1] == 1   # pipeline:646-647 comment
    cfg = os.path.join("/root/reference", "scheduler_1.0_1.0", "scheduler_config.json")
    if os.path.exists(cfg):           # only in the build container; the GPU box has no /root/reference
        with open(cfg) as f:
            ref = json.load(f)
        from diffews_b200.scheduler import DEFAULT_CONFIG
        for k_, v_ in DEFAULT_CONFIG.items():
            assert ref[k_] == v_, k_


def test_rthres_golden_table():
    """tests/golden/rthres_cases.json: (R,G,B,max) -> mean(dim=1) > max*0.25 evaluated by torch CPU fp32 exactly as
    main_oss.py:128-134 does; includes exact ties where the integer rule 4(R+G+B) > 3 max disagrees."""
    with open(os.path.join(GOLDEN, "rthres_cases.json")) as f:
        cases = json.load(f)
    n_tie_disagree = 0
    for r, g, b, mx, expect in cases["cases"]:
        img = torch.zeros(1, 3, 1, 2, dtype=torch.uint8)
        img[0, :, 0, 0] = torch.tensor([r, g, b], dtype=torch.uint8)
        img[0, 0, 0, 1] = mx                                  # a second pixel that sets the episode max
        got = bool(om.rthres_mask(img, 0.25)[0, 0, 0])
        assert got == bool(expect), (r, g, b, mx)
        if 4 * (r + g + b) == 3 * mx and got:
            n_tie_disagree += 1
    assert n_tie_disagree >= 5          # the table must contain ties on which the integer rule is wrong


def test_histc_drops_ignore_value():
    pred = torch.tensor([[[0., 1., 1., 0.], [1., 1., 0., 0.]]])
    gt = torch.tensor([[[0., 1., 0., 0.], [1., 0., 0., 0.]]])
    ign = torch.tensor([[[0., 0., 1., 0.], [0., 0., 0., 1.]]])
    inter, union = om.classify_prediction(pred.clone(), {"query_mask": gt.clone(), "query_ignore_idx": ign.clone()})
    # ignored pixels (2 of 8) vanish from every histogram
    assert inter[:, 0].tolist() == [3.0, 2.0]                      # 6 live pixels: 3 agree on 0, 2 agree on 1
    assert union[:, 0].tolist() == [3.0 + 4.0 - 3.0, 3.0 + 2.0 - 2.0]
    inter2, union2 = om.classify_prediction(pred.clone(), {"query_mask": gt.clone()})
    assert inter2[:, 0].tolist() == [4.0, 2.0] and union2[:, 0].tolist() == [4.0 + 6.0 - 4.0, 4.0 + 2.0 - 2.0]


def test_average_meter_formulas():
    m = om.AverageMeter("coco", range(80), exact=True)
    inter = torch.tensor([[10., 20.], [5., 7.]])      # [2, B=2]
    union = torch.tensor([[20., 40.], [10., 7.]])
    m.update(inter, union, torch.tensor([3, 3]))
    m.update(inter, union, torch.tensor([3, 9]))
    miou, fb, _ = m.compute_iou()
    iou3 = (5 + 7 + 5) / (10 + 7 + 10)
    iou9 = 7 / 7
    assert abs(float(miou) - (iou3 + iou9) / 80 * 100) < 1e-4
    fg = (5 + 7 + 5 + 7) / (10 + 7 + 10 + 7)
    bg = (10 + 20 + 10 + 20) / (20 + 40 + 20 + 40)
    assert abs(float(fb) - (fg + bg) / 2 * 100) < 1e-4


def test_oracle_pipeline_golden_fixture():
    """Drift guard: the oracle's own output on a seeded reduced-width episode (tests/golden/oracle_small_episode.json,
    produced by scripts/make_golden.py)."""
    with open(os.path.join(GOLDEN, "oracle_small_episode.json")) as f:
        gold = json.load(f)
    from diffews_b200.synthetic import make_batch, prompt_embedding
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    unet, vae = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    inter, union, mask, seg_u8, lat = op.evaluate_episode(unet, vae, prompt_embedding(), make_batch(0, 1, 64, 1))
    assert abs(float(lat.double().abs().mean()) - gold["latent_abs_mean"]) < 1e-4 * gold["latent_abs_mean"]
    assert abs(float(seg_u8.double().mean()) - gold["seg_u8_mean"]) < 0.05
    # counts depend on threshold ties; allow a handful of pixels of slack across BLAS builds
    for a, b in zip(inter[:, 0].tolist() + union[:, 0].tolist(), gold["inter"] + gold["union"]):
        assert abs(a - b) <= 8


def test_ddim_multi_step_schedule_and_identity():
    """set_timesteps (leading spacing, offset 1) and the full DDIM step under the DiffewS scheduler (beta == 1 =>
    alpha_cumprod == 0): x0 = -v at every step and x_{t-1} == x_t, i.e. extra steps re-run the same prediction at other
    timesteps (pipeline:706-767 with scheduler_1.0_1.0/scheduler_config.json)."""
    import torch
    from oracle import pipeline as op
    from diffews_b200.scheduler import DDIMSchedulerCustomized
    assert op.ddim_timesteps(1) == [1]
    assert op.ddim_timesteps(4) == [751, 501, 251, 1]
    sch = DDIMSchedulerCustomized()
    sch.set_timesteps(4)
    assert [int(t) for t in sch.timesteps] == op.ddim_timesteps(4)
    g = torch.Generator().manual_seed(0)
    v, x = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for t in op.ddim_timesteps(4):
        assert torch.equal(op.ddim_step_pred_original(v, t, x), -v)
        assert torch.allclose(op.ddim_step_prev_sample(v, t, x, 4), x)
        out = sch.step(v, t, x)
        assert torch.allclose(out.pred_original_sample, -v) and torch.allclose(out.prev_sample, x)


def test_metric_oracle_matches_unmodified_reference_evaluator():
    """a12 pinned: tests/golden/metric_reference.json holds what the UNMODIFIED reference
    evaluation_util/common/evaluation.py Evaluator.classify_prediction returned on CPU for the seeded masks of
    tests/data_tree.py::metric_cases (scripts/make_golden_data.py); the oracle restatement must reproduce it exactly."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import data_tree
    from oracle.metric import classify_prediction
    gold = json.load(open(os.path.join(here, "golden", "metric_reference.json")))["cases"]
    cases = data_tree.metric_cases()
    assert len(gold) == len(cases)
    for c, g in zip(cases, gold):
        assert (c["seed"], c["B"], c["H"], c["W"]) == (g["seed"], g["B"], g["H"], g["W"])
        batch = {"query_mask": c["gt"].clone()}
        if c["ign"] is not None:
            batch["query_ignore_idx"] = c["ign"].clone()
        inter, union = classify_prediction(c["pred"].clone(), batch)
        assert inter.dtype == torch.float32 and inter.tolist() == g["area_inter"], g["kind"]
        assert union.tolist() == g["area_union"], g["kind"]


def test_meter_oracle_matches_unmodified_reference_average_meter():
    """a13 pinned: the `meter` block of tests/golden/metric_reference.json is what the UNMODIFIED
    evaluation_util/common/logger.py AverageMeter (update + compute_iou) produced when fed the reference Evaluator's own
    per-episode outputs (scripts/make_golden_data.py; Tensor.cuda patched to the identity, the source untouched).  The
    oracle meter must reproduce it in its float32-faithful mode exactly, and in the int64 mode the B200 path uses to
    float32 rounding (the counts here are far below 2^24, so the buffers are equal too)."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "metric_reference.json")))
    cases, mg = gold["cases"], gold["meter"]
    for exact in (False, True):
        meter = om.AverageMeter(mg["benchmark"], mg["class_ids_interest"], exact=exact)
        it = iter(mg["update_class_ids"])
        for rep in range(3):
            for c in cases:
                meter.update(torch.tensor(c["area_inter"]), torch.tensor(c["area_union"]), torch.tensor(next(it)))
        miou, fb_iou, head = meter.compute_iou()
        assert meter.intersection_buf.float().tolist() == mg["intersection_buf"]
        assert meter.union_buf.float().tolist() == mg["union_buf"]
        assert float(miou) == mg["miou"] and float(fb_iou) == mg["fb_iou"]
        assert head.tolist() == mg["iou_head"]


def test_attention_oracle_matches_unmodified_reference_processor():
    """a5 pinned: tests/golden/attn_reference.json holds the query-pass outputs of the reference's OWN
    diffews/models/attention_processor.py (MyAttention + MyXFormersAttnProcessor for k = 1, 3, 5; MyAttnProcessor2_0 for
    k = 1), executed unmodified over stand-ins for the diffusers / xformers names it imports
    (scripts/make_golden_attn.py).  The oracle's KV-bank attention — the k-shot fold restated as a shot-major
    concatenation — must reproduce them (fp32, different but equivalent op order: 2e-6)."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import data_tree
    gold = json.load(open(os.path.join(here, "golden", "attn_reference.json")))["cases"]
    cases = data_tree.attn_cases()
    assert len(gold) == len(cases)
    for c, g in zip(cases, gold):
        assert (c["B"], c["k"]) == (g["B"], g["k"])
        attn = sd21.Attention(c["C"], heads=c["heads"], dim_head=c["C"] // c["heads"], bank=True)
        with torch.no_grad():
            attn.to_q.weight.copy_(c["w"]["to_q"]); attn.to_k.weight.copy_(c["w"]["to_k"])
            attn.to_v.weight.copy_(c["w"]["to_v"]); attn.to_out[0].weight.copy_(c["w"]["to_out"])
            attn.to_out[0].bias.copy_(c["w"]["to_out_bias"])
            attn.clear_bank()
            sup = attn(c["x_support"])
            qry = attn(c["x_query"])
        for name in ("xformers", "sdpa"):
            if name not in g:
                continue
            ref = torch.tensor(g[name]["query_out"]).view_as(qry)
            assert (qry - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 2e-6, (c["B"], c["k"], name)
            assert abs(float(sup.double().sum()) - g[name]["support_out_sum"]) <= 1e-4 * abs(g[name]["support_out_sum"]) + 1e-3
        if "sdpa" in g:     # the reference's two processors agree with each other at k = 1
            a, b = torch.tensor(g["xformers"]["query_out"]), torch.tensor(g["sdpa"]["query_out"])
            assert (a - b).abs().max().item() <= 2e-6 * a.abs().max().item() + 2e-6


def test_scheduler_tables_match_unmodified_reference_scheduler():
    """a10 premise pinned: tests/golden/scheduler_reference.json = the beta / alpha tables that the UNMODIFIED
    marigold/util/scheduler_customized.py DDIMSchedulerCustomized.__init__ builds from scheduler_1.0_1.0/
    scheduler_config.json (scripts/make_golden_attn.py).  Every alpha_cumprod is exactly 0 there, which is what makes the
    v-prediction DDIM step collapse to z0 = -v; the product scheduler must build the same tables and detect the collapse."""
    import json
    import os
    from diffews_b200.scheduler import DDIMSchedulerCustomized
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "scheduler_reference.json")))
    sch = DDIMSchedulerCustomized(**g["config_used"])
    assert sch.betas[:3].tolist() == g["betas_head"]
    assert float(sch.betas.min()) == g["betas_min"] and float(sch.betas.max()) == g["betas_max"]
    assert float(sch.alphas_cumprod.abs().max()) == g["alphas_cumprod_abs_max"] == 0.0
    assert sch.alphas_cumprod.numel() == g["alphas_cumprod_len"]
    assert float(sch.final_alpha_cumprod) == g["final_alpha_cumprod"] == 0.0
    assert sch.timesteps[:3].tolist() == g["timesteps_head"]
    sch.set_timesteps(1)
    t = int(sch.timesteps[0])
    assert sch.is_pure_negation(t)
    v, x = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    assert torch.equal(sch.step(v, t, x).pred_original_sample, -v)


def test_pipeline_oracle_matches_unmodified_reference_single_infer():
    """a2 / a3 / a9 / a10 wiring pinned: tests/golden/pipeline_reference.json = the output of the reference's OWN
    MarigoldPipelineRGBLatentNoise.single_infer (diffews/marigold_pipeline_rgb_latent_noise.py executed unmodified over
    import stand-ins, scripts/make_golden_pipeline.py) when it drives the oracle UNet / VAE and the restated scheduler.
    The oracle's restatement of single_infer on the same sub-modules and inputs must give the same image (same fp32 ops in
    the same order; 2e-3 on the 0..255 scale leaves room for another CPU's conv kernels)."""
    import base64
    import json
    import os
    import numpy as np
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "pipeline_reference.json")))["cases"]
    unet, vae = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    for g in gold:
        ref_imgs, tag, gt = pipeline_inputs(make_batch(0, g["B"], g["size"], g["k"]))
        seg = op.single_infer(unet, vae, prompt_embedding(), ref_imgs, tag, gt)
        lat = op.encode_rgb(vae, tag)
        want = torch.from_numpy(np.frombuffer(base64.b64decode(g["seg"]["sub_f32_b64"]), dtype=np.float32).copy())
        want = want.view(g["seg"]["sub_shape"])
        assert list(seg.shape) == g["seg"]["shape"]
        assert (seg[:, :, ::4, ::4] - want).abs().max().item() <= 2e-3, (g["B"], g["k"])
        assert abs(float(seg.double().mean()) - g["seg"]["mean"]) <= 1e-3
        assert float(seg.min()) == g["seg"]["min"] and float(seg.max()) == g["seg"]["max"]
        assert abs(float(lat.double().abs().mean()) - g["encode_rgb_tag_abs_mean"]) <= 1e-6


def test_pipeline_oracle_matches_unmodified_reference_call():
    """a1 pinned: the `call` block of tests/golden/pipeline_reference.json is what the reference's OWN
    MarigoldPipelineRGBLatentNoise.__call__ returned (unmodified file; tensor inputs, a real file in rgb_paths, bsz 1 —
    exactly how test_diffusion calls it, main_oss.py:113-123): a PIL RGB image, `uncertainty=None`.  The oracle's
    single_infer + the uint8 truncation of pipeline:534 must give the same bytes (<= 1 LSB on <= 0.1 % of the pixels is
    tolerated for another CPU's conv kernels)."""
    import base64
    import json
    import os
    import numpy as np
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "pipeline_reference.json")))["call"]
    assert g["type"] == "Image" and g["dtype"] == "uint8" and g["uncertainty_is_none"]
    unet, vae = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    ref_imgs, tag, gt = pipeline_inputs(make_batch(0, g["B"], g["size"], g["k"]))
    seg_u8 = op.to_uint8(op.single_infer(unet, vae, prompt_embedding(), ref_imgs, tag, gt))     # [1,3,H,W]
    img = seg_u8[0].permute(1, 2, 0).numpy()                                                      # chw2hwc
    assert list(img.shape) == g["shape"]
    want = np.frombuffer(base64.b64decode(g["u8_b64"]), dtype=np.uint8).reshape(g["size"] // 2, g["size"] // 2, 3)
    diff = np.abs(img[::2, ::2].astype(np.int32) - want.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3
    assert abs(int(img.astype(np.int64).sum()) - g["sum"]) <= 16


def test_oracle_eval_loop_matches_unmodified_reference_test_diffusion(tmp_path):
    """The whole path pinned at loop level: tests/golden/eval_loop_reference.json = per-episode intersection / union and
    the final mIoU / FB-IoU of the reference's OWN `test_diffusion` (evaluation_util/main_oss.py:84-171, unmodified, with
    its own dataset, transform, pipeline __call__, inline rthres, Evaluator and AverageMeter; the oracle UNet / VAE and the
    restated scheduler plugged in — scripts/make_golden_eval_loop.py).  The oracle's restatement of the same loop (data
    layer, input folding, single_infer, uint8 truncation, rthres, counts, meter) must reproduce it: identical episodes, counts
    within 2 pixels (exactly equal on the build machine; the slack is for another CPU's conv kernels), mIoU within 0.05."""
    import json
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import data_tree
    from diffews_b200 import data as pdata
    from diffews_b200.synthetic import prompt_embedding
    from oracle import data as od
    gold = json.load(open(os.path.join(here, "golden", "eval_loop_reference.json")))
    S = gold["img_size"]
    tree = str(tmp_path)
    data_tree.build_coco_tree(tree)
    ds = pdata.DatasetCOCO(tree, fold=0, transform=pdata.EpisodeTransform(S), split="test", shot=1,
                           use_original_imgsize=False)
    unet, vae = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    meter = om.AverageMeter("coco", ds.class_ids, exact=False)
    np.random.seed(0)
    for idx, g in enumerate(gold["episodes"]):
        raw = ds.raw_episode(idx)
        assert raw["query_name"] == g["query_name"] and raw["support_names"] == g["support_names"]
        assert raw["class_sample"] == g["class_id"]
        batch = {"query_img": od.transform_image(raw["query_img"], S)[None],
                 "query_mask": od.coco_mask(raw["query_label"], raw["class_sample"], S)[None],
                 "support_imgs": torch.stack([od.transform_image(a, S) for a in raw["support_imgs"]])[None],
                 "support_masks": torch.stack([od.coco_mask(l, raw["class_sample"], S) for l in raw["support_labels"]])[None],
                 "class_id": torch.tensor([raw["class_sample"]])}
        inter, union, pred_mask, _, _ = op.evaluate_episode(unet, vae, prompt_embedding(), batch)
        assert abs(int(pred_mask.sum()) - g["pred_fg"]) <= 2
        assert (inter - torch.tensor(g["area_inter"])).abs().max().item() <= 2, (idx, inter.tolist(), g["area_inter"])
        assert (union - torch.tensor(g["area_union"])).abs().max().item() <= 2
        meter.update(inter, union, batch["class_id"])
    miou, fb_iou, _ = meter.compute_iou()
    assert abs(float(miou) - gold["miou"]) <= 0.05 and abs(float(fb_iou) - gold["fb_iou"]) <= 0.05


def test_unet_oracle_matches_unmodified_reference_forward_wiring():
    """a4 wiring pinned: tests/golden/unet_wiring_reference.json = the query-pass output of the reference's OWN
    MyUNet2DConditionModel.forward (+ clear_attn_bank), executed unmodified on an instance hand-assembled from the oracle's
    blocks behind call-protocol adapters (scripts/make_golden_unet_wiring.py).  The oracle UNet's own forward — the
    restatement of that wiring — must give the same tensor on the same inputs."""
    import base64
    import json
    import os
    import numpy as np
    from diffews_b200.synthetic import prompt_embedding
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "unet_wiring_reference.json")))["cases"]
    unet, _ = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    ehs = prompt_embedding()
    for g in gold:
        B, k, lat = g["B"], g["k"], g["lat"]
        gen = torch.Generator().manual_seed(300 + 10 * B + k)
        sup = torch.randn(B * k, 8, lat, lat, generator=gen) * 0.8
        qry = torch.randn(B, 4, lat, lat, generator=gen) * 0.8
        with torch.no_grad():
            unet.clear_attn_bank()
            s = unet(sup, 1, ehs.repeat(B * k, 1, 1), is_target=False)
            y = unet(qry, 1, ehs.repeat(B, 1, 1))
            unet.clear_attn_bank()
        want = torch.from_numpy(np.frombuffer(base64.b64decode(g["f32_b64"]), dtype=np.float32).copy()).view(g["shape"])
        assert g["type"] == "UNet2DConditionOutput" and list(y.shape) == g["shape"]
        assert (y - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item()), (B, k)
        assert abs(float(s.double().abs().mean()) - g["support_abs_mean"]) <= 1e-6
