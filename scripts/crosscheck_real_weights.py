"""One-command cross-validation against the real thing — for a machine that HAS what this sandbox
lacks: the released DiffewS weights (ModelScope `zzzmmz/Diffews`, reference README.md:71), diffusers 0.25, xformers and
the reference checkout.  It runs the same episodes through

  (A) the UNMODIFIED reference: CustomUNet2DConditionModel / AutoencoderKL / MarigoldPipelineRGBLatentNoise loaded exactly
      as evaluation_util/main_oss.py:338-379 does (fp32, xformers processors, scheduler_1.0_1.0), and
  (B) this repo's B200 engine loaded from the same checkpoint directory (MarigoldPipelineRGBLatentNoise.from_pretrained),

and reports per episode: UNet-latent rel-L2 (bar 1e-2 for the fp16 engine, 1e-4 with --operands f32), uint8-image
agreement, rthres-mask agreement (bar 99.5 %) and the intersection / union counts.  Episodes are synthetic
(diffews_b200.synthetic, seeded) unless --datapath / --benchmark point at a real dataset tree, in which case the reference's
own FSSDataset feeds (A) and diffews_b200.data feeds (B) with the same np.random seed.

    python scripts/crosscheck_real_weights.py --reference /path/to/DiffewS --checkpoint /path/to/weight/stable-diffusion-2-1-ref8inchannels-tag4inchannels \\
        --unet-ckpt /path/to/weight/coco_fold0 [--episodes 8] [--size 512] [--nshot 1] [--operands f16|f32]

Nothing here runs in the build sandbox (no diffusers / xformers / weights / network): the script is shipped so that a
maintainer can close the last parity gap — the insides of the diffusers blocks, which DESIGN.md section 7 lists as restated
without a reference run behind them.  It exits non-zero when a bar is missed.
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def reference_pipeline(args, device):
    """main_oss.py:338-379, verbatim in structure."""
    sys.path.insert(0, args.reference)
    from diffusers import AutoencoderKL, DDIMScheduler
    from transformers import CLIPTokenizer  # noqa: F401  (loaded by DiffusionPipeline.from_pretrained)
    from diffews.marigold_pipeline_rgb_latent_noise import MarigoldPipelineRGBLatentNoise as RefPipe
    from diffews.models.unet_2d_condition import MyUNet2DConditionModel as RefUNet
    unet = RefUNet.from_pretrained(args.unet_ckpt or args.checkpoint, subfolder="unet")
    vae = AutoencoderKL.from_pretrained(args.checkpoint, subfolder="vae")
    params = dict(torch_dtype=torch.float32, unet=unet, vae=vae, controlnet=None, text_embeds=None, image_projector=None,
                  customized_head=None, image_encoder=None)
    sched = args.scheduler or os.path.join(args.reference, "scheduler_1.0_1.0")
    params["scheduler"] = DDIMScheduler.from_pretrained(sched, subfolder="scheduler" if os.path.isdir(os.path.join(sched, "scheduler")) else None)
    pipe = RefPipe.from_pretrained(args.checkpoint, **params).to(device)
    pipe.test_timestep = 1
    pipe.enable_xformers_memory_efficient_attention()
    return pipe


def engine_pipeline(args, device, text_embeds):
    from diffews_b200 import checkpoint
    from diffews_b200.layers import Precision
    from diffews_b200.pipeline import MarigoldPipelineRGBLatentNoise
    from diffews_b200.scheduler import DDIMSchedulerCustomized
    prec = Precision(f32=True) if args.operands == "f32" else None
    unet = checkpoint.load_unet(args.unet_ckpt or args.checkpoint, device=device, precision=prec)
    vae = checkpoint.load_vae(args.checkpoint, device=device, precision=prec)
    sched = args.scheduler or os.path.join(args.reference, "scheduler_1.0_1.0")
    cfg = os.path.join(sched, "scheduler", "scheduler_config.json")
    if not os.path.exists(cfg):
        cfg = os.path.join(sched, "scheduler_config.json")
    pipe = MarigoldPipelineRGBLatentNoise(unet, vae, scheduler=DDIMSchedulerCustomized.from_config_file(cfg),
                                          text_embeds=text_embeds)
    pipe.test_timestep = 1
    return pipe


def rthres(pil_or_u8, r=0.25):
    """evaluation_util/main_oss.py:125-137 on the CPU in fp32."""
    import numpy as np
    a = torch.from_numpy(np.asarray(pil_or_u8)).permute(2, 0, 1).float().div(255)[None] if not torch.is_tensor(pil_or_u8) \
        else pil_or_u8.float().div(255)[None]
    thr = a.max() * r
    return (a.mean(dim=1) > thr).float()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of ga1i13o/DiffewS")
    ap.add_argument("--checkpoint", required=True, help="diffusers directory with unet/ vae/ text_encoder/ tokenizer/")
    ap.add_argument("--unet-ckpt", default=None, help="directory whose unet/ subfolder holds the fine-tuned DiffewS UNet")
    ap.add_argument("--scheduler", default=None, help="scheduler directory (default: <reference>/scheduler_1.0_1.0)")
    ap.add_argument("--episodes", type=int, default=8)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--nshot", type=int, default=1)
    ap.add_argument("--operands", default="f16", choices=["f16", "f32"])
    args = ap.parse_args()
    device = torch.device("cuda")
    from diffews_b200.evaluation import Evaluator
    from diffews_b200.synthetic import make_batch, pipeline_inputs
    ref = reference_pipeline(args, device)
    emb = ref.encode_clip_feature(None).detach().float().cpu()             # the reference's own empty-prompt embedding
    eng = engine_pipeline(args, device, emb)
    lat_bar = 1e-4 if args.operands == "f32" else 1e-2
    figs = os.path.join(args.reference, "figs")
    any_image = os.path.join(figs, sorted(f for f in os.listdir(figs) if f.lower().endswith((".png", ".jpg")))[0])
    worst_lat, worst_mask, ok = 0.0, 1.0, True
    for e in range(args.episodes):
        batch = make_batch(e, 1, args.size, args.nshot)
        sup, qry, gt = pipeline_inputs(batch)
        with torch.no_grad():
            out_ref = ref([sup.to(device), qry.to(device), gt.to(device)], denoising_steps=1, ensemble_size=1,
                          processing_res=args.size, match_input_res=True, batch_size=1, show_progress_bar=False, mode="seg",
                          rgb_paths=[any_image], seed=0)     # the reference opens it for unused CLIP image features
        out_eng = eng([sup.to(device), qry.to(device), gt.to(device)], denoising_steps=1, ensemble_size=1,
                      processing_res=args.size, batch_size=1, show_progress_bar=False, mode="seg", rgb_paths=[], seed=0,
                      output_type="pt")
        m_ref = rthres(out_ref.seg_colored)
        inter, union, m_eng = Evaluator.rthres_classify(out_eng.seg_u8, {"query_mask": batch["query_mask"].to(device)}, 0.25,
                                                        want_mask=True)
        agree = (m_eng[0].cpu().float() == m_ref[0]).float().mean().item()
        # UNet latent: the reference does not expose it; re-run its UNet on the engine's own UNet inputs (identical inputs)
        s_lat, q_lat = (t.to(device) for t in eng._last_unet_inputs)
        with torch.no_grad():
            ref.unet.clear_attn_bank()
            ref.unet(s_lat, 1, emb.to(device).repeat(s_lat.shape[0], 1, 1), is_target=False)
            want = ref.unet(q_lat, 1, emb.to(device)).sample
            ref.unet.clear_attn_bank()
        lat = ((eng._last_noise_pred.float() - want.float()).norm() / want.float().norm()).item()
        worst_lat, worst_mask = max(worst_lat, lat), min(worst_mask, agree)
        print(f"episode {e}: UNet latent rel-L2 {lat:.3e} (bar {lat_bar:g}), mask agreement {agree:.5f} (bar 0.995), "
              f"engine inter {inter[0].tolist()} union {union[0].tolist()}")
        ok &= lat <= lat_bar and agree >= 0.995
    print(f"worst latent rel-L2 {worst_lat:.3e}, worst mask agreement {worst_mask:.5f}: {'PASS' if ok else 'FAIL'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
