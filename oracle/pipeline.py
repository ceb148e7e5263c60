"""ORACLE (test infrastructure) — CPU fp32 restatement of the DiffewS inference path end to end.

Follows diffews/marigold_pipeline_rgb_latent_noise.py:
  :616-836 single_infer  (3 VAE encodes :649-651, cond concat :674, embed repeat :690-692, clear/support/query/clear
                          :715-725, scheduler.step().pred_original_sample :764-769, decode_seg + clip + *0.5+0.5 + *255
                          :787-795)
  :839-862 encode_rgb, :887-905 decode_seg, :534 uint8 truncation,
and evaluation_util/main_oss.py:99-137 (input folding, rthres) + evaluation.py:12-39 via oracle.metric.
The scheduler (marigold/util/scheduler_customized.py:107-180 with scheduler_1.0_1.0/scheduler_config.json) is restated
in full in `ddim_step` and shown by tests to be exactly z0 = -v.
"""
from __future__ import annotations

import torch

from .metric import classify_prediction, rthres_mask

SCALE = 0.18215          # pipeline:120-124


def ddim_alphas_cumprod(beta_start=1.0, beta_end=1.0, n=1000):
    """scheduler_customized.py:133-158, `scaled_linear`."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_step_pred_original(model_output, timestep: int, sample, alphas_cumprod=None):
    """diffusers DDIMScheduler.step, prediction_type='v_prediction':  x0 = sqrt(a_t) x - sqrt(1-a_t) v."""
    ac = ddim_alphas_cumprod() if alphas_cumprod is None else alphas_cumprod
    a_t = ac[timestep]
    return (a_t ** 0.5) * sample - ((1 - a_t) ** 0.5) * model_output


def ddim_timesteps(num_inference_steps: int, n_train=1000, steps_offset=1):
    """DDIMScheduler.set_timesteps, timestep_spacing='leading' (scheduler_1.0_1.0/scheduler_config.json), pipeline:644."""
    ratio = n_train // num_inference_steps
    return [int(i * ratio) + steps_offset for i in reversed(range(num_inference_steps))]


def ddim_step_prev_sample(model_output, timestep: int, sample, num_inference_steps: int, alphas_cumprod=None,
                          set_alpha_to_one=False):
    """diffusers DDIMScheduler.step (eta = 0, v_prediction): x_{t-1} = sqrt(a_prev) x0 + sqrt(1 - a_prev) eps."""
    ac = ddim_alphas_cumprod() if alphas_cumprod is None else alphas_cumprod
    a_t = ac[timestep]
    prev_t = timestep - 1000 // num_inference_steps
    a_prev = ac[prev_t] if prev_t >= 0 else (torch.tensor(1.0) if set_alpha_to_one else ac[0])
    x0 = (a_t ** 0.5) * sample - ((1 - a_t) ** 0.5) * model_output
    eps = (a_t ** 0.5) * model_output + ((1 - a_t) ** 0.5) * sample
    return (a_prev ** 0.5) * x0 + ((1 - a_prev) ** 0.5) * eps


def encode_rgb(vae, rgb_in):                                   # pipeline:839-862
    moments = vae.quant_conv(vae.encoder(rgb_in))
    mean, _logvar = torch.chunk(moments, 2, dim=1)
    return mean * SCALE


def decode_seg(vae, seg_latent):                               # pipeline:887-905
    z = vae.post_quant_conv(seg_latent / SCALE)
    return vae.decoder(z).clip(-1, 1)


@torch.no_grad()
def single_infer(unet, vae, text_embed, rgb_in_ref, rgb_in_tag, gt_in_ref, test_timestep=1, return_latent=False,
                 num_inference_steps=1):
    timesteps = ddim_timesteps(num_inference_steps)            # 1 step: [1] (leading spacing, offset 1, :644)
    rgb_latent_ref = encode_rgb(vae, rgb_in_ref)
    rgb_latent_tag = encode_rgb(vae, rgb_in_tag)
    gt_latent_ref = encode_rgb(vae, gt_in_ref)
    latents_rgb_cond_ref = torch.cat([rgb_latent_ref, gt_latent_ref], dim=1)
    depth_latent = rgb_latent_tag.clone()
    batch_embed = text_embed.repeat((rgb_latent_tag.shape[0], 1, 1))
    batch_embed_ref = batch_embed.repeat((rgb_latent_ref.shape[0] // rgb_latent_tag.shape[0], 1, 1))
    for t in timesteps:
        unet.clear_attn_bank()
        unet(latents_rgb_cond_ref, t * test_timestep, batch_embed_ref, is_target=False)
        noise_pred = unet(depth_latent, t * test_timestep, batch_embed)
        unet.clear_attn_bank()
        z0 = ddim_step_pred_original(noise_pred, t, depth_latent)                     # :764-769
        depth_latent = ddim_step_prev_sample(noise_pred, t, depth_latent, num_inference_steps)
    seg = decode_seg(vae, z0)
    seg = torch.clip(seg, -1.0, 1.0)
    seg = (seg * 0.5) + 0.5
    seg = seg * 255
    if return_latent:
        return seg, noise_pred
    return seg


def to_uint8(seg):                                             # pipeline:534  clip(0,255) -> numpy astype(uint8)
    return torch.from_numpy(seg.clip(0, 255).cpu().numpy().astype("uint8"))


@torch.no_grad()
def evaluate_episode(unet, vae, text_embed, batch1, r_threshold=0.25):
    """One bsz=1 iteration of test_diffusion (main_oss.py:92-155): returns (area_inter [2,1], area_union [2,1],
    pred_mask [1,H,W], seg_u8 [1,3,H,W], unet latent)."""
    from diffews_b200.synthetic import pipeline_inputs
    ref, tag, gt = pipeline_inputs(batch1)
    seg, lat = single_infer(unet, vae, text_embed, ref, tag, gt, return_latent=True)
    seg_u8 = to_uint8(seg)
    pred_mask = rthres_mask(seg_u8, r_threshold)
    inter, union = classify_prediction(pred_mask.clone(), {"query_mask": batch1["query_mask"]})
    return inter, union, pred_mask, seg_u8, lat
