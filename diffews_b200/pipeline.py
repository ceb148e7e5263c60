"""MarigoldPipelineRGBLatentNoise — drop-in for diffews/marigold_pipeline_rgb_latent_noise.py (reference).

Same constructor keywords, `__call__` keywords, `single_infer` / `encode_rgb` / `decode_seg` methods, class-level
latent scale factors and the externally set `test_timestep` attribute (evaluation_util/main_oss.py:373) as the
reference; inputs and outputs keep the reference's NCHW fp32 tensor layouts.  What runs underneath is the B200 engine
(`diffews_b200.unet`, `diffews_b200.vae`).  Differences, all extensions or reference-defect tolerances:

  * `rgb_paths` is accepted and ignored: the reference opens those files only to compute CLIP *image* features that
    are unused in text-embedding mode (pipeline:311-325, :590-601), and with empty `rgb_paths` it raises NameError
    (`transforms` is never imported, :321).
  * `output_type="pt"` returns the uint8 segmentation as a CUDA tensor [B,3,H,W] (batched evaluation, which the
    reference cannot do: its eval loop only works for bsz=1).  Default "pil" matches the reference.
  * only `mode='seg'` (any non-'depth' mode is treated as seg by the reference, :280, :532) and
    `denoising_steps == 1` ("nosample") is the hot path; more steps run the same two UNet passes per step (:706-767).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import torch

from . import ops
from .scheduler import DDIMSchedulerCustomized


@dataclass
class MarigoldSegOutput:
    seg_colored: object                 # PIL.Image | list[PIL.Image] | torch.Tensor (output_type="pt")
    uncertainty: Optional[object] = None
    seg_u8: Optional[torch.Tensor] = None   # always: uint8 CUDA tensor [B,3,H,W]


class MarigoldPipelineRGBLatentNoise:
    rgb_latent_scale_factor = 0.18215       # pipeline:120-124
    depth_latent_scale_factor = 0.18215
    seg_latent_scale_factor = 0.18215
    sr_latent_scale_factor = 0.18215
    normal_latent_scale_factor = 0.18215

    def __init__(self, unet, vae, scheduler=None, tokenizer=None, text_embeds: Optional[torch.Tensor] = None,
                 text_encoder=None, image_encoder=None, image_projector=None, controlnet=None, customized_head=None):
        if image_encoder is not None or image_projector is not None:
            raise NotImplementedError("vision-embedding conditioning is not on the DiffewS hot path")
        if controlnet is not None or customized_head is not None:
            raise NotImplementedError("controlnet / customized_head are not on the DiffewS hot path "
                                      "(reference: customized_head=None, main_oss.py:362)")
        if text_embeds is None and text_encoder is None:
            raise ValueError("need text_embeds (the empty-prompt embedding) or a text_encoder")   # pipeline:158-159
        self.unet, self.vae = unet, vae
        self.scheduler = scheduler if scheduler is not None else DDIMSchedulerCustomized()
        self.tokenizer, self.text_encoder = tokenizer, text_encoder
        self.empty_text_embed = text_embeds
        self.text_embed_flag, self.vision_embed_flag = True, False
        self.test_timestep = 1
        self.device = unet.device
        self.dtype = torch.float32
        self._embed_cache = {}
        self.validate_inputs = True     # the reference asserts the [-1,1] input range (pipeline:309)
        if hasattr(unet, "skip_support_tail"):
            unet.skip_support_tail = True   # single_infer discards the support pass's output (pipeline:719-720)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, torch_dtype=None, unet=None, vae=None, scheduler=None,
                        tokenizer=None, text_encoder=None, text_embeds=None, controlnet=None, image_projector=None,
                        customized_head=None, image_encoder=None, device="cuda", unet_precision=None, vae_precision=None,
                        **unused):
        """`MarigoldPipeline.from_pretrained(args.checkpoint, torch_dtype=..., unet=unet, vae=vae, controlnet=None,
        text_embeds=None, ...)` (main_oss.py:351-369): components that are not passed are loaded from the diffusers
        directory -- `unet/`, `vae/` through diffews_b200.checkpoint, `scheduler/scheduler_config.json`, and, when no
        `text_embeds` is given, `text_encoder/` + `tokenizer/` through transformers' CLIPTextModel / CLIPTokenizer (library
        code, run ONCE for the empty prompt: pipeline:585-601).  Local files only; `torch_dtype` is accepted and ignored
        like the reference ignores it for pre-built sub-models."""
        import os
        from . import checkpoint
        root = str(pretrained_model_name_or_path)
        if unet is None:
            unet = checkpoint.load_unet(root, device=device, precision=unet_precision)
        if vae is None:
            vae = checkpoint.load_vae(root, device=device, precision=vae_precision)
        if scheduler is None and os.path.exists(os.path.join(root, "scheduler", "scheduler_config.json")):
            scheduler = DDIMSchedulerCustomized.from_config_file(os.path.join(root, "scheduler", "scheduler_config.json"))
        if text_embeds is None:
            if text_encoder is None:
                from transformers import CLIPTextModel
                text_encoder = CLIPTextModel.from_pretrained(os.path.join(root, "text_encoder"), local_files_only=True).eval()
            if tokenizer is None:
                from transformers import CLIPTokenizer
                tokenizer = CLIPTokenizer.from_pretrained(os.path.join(root, "tokenizer"), local_files_only=True)
        return cls(unet=unet, vae=vae, scheduler=scheduler, tokenizer=tokenizer, text_embeds=text_embeds,
                   text_encoder=text_encoder, image_encoder=image_encoder, image_projector=image_projector,
                   controlnet=controlnet, customized_head=customized_head)

    def to(self, *a, **k):
        return self

    def enable_xformers_memory_efficient_attention(self, attention_op=None):   # main_oss.py:374-379
        self.unet.enable_xformers_memory_efficient_attention(attention_op)

    # ------------------------------------------------------------------------------------------------------------
    def encode_clip_feature(self, clip_rgb_in=None):                            # pipeline:585-601
        if self.empty_text_embed is not None:
            return self.empty_text_embed
        prompt = ""
        ids = self.tokenizer(prompt, padding="do_not_pad", max_length=self.tokenizer.model_max_length,
                             truncation=True, return_tensors="pt").input_ids.to(self.text_encoder.device)
        with torch.no_grad():
            self.empty_text_embed = self.text_encoder(ids)[0].to(self.dtype)        # [1, 2, 1024]: <bos>, <eos>
        return self.empty_text_embed

    def _batch_embed(self, n: int) -> torch.Tensor:
        hit = self._embed_cache.get(n)
        if hit is None:
            e = self.encode_clip_feature().to(self.device, torch.float32)
            hit = e.repeat((n, 1, 1)).contiguous()                              # pipeline:690-692
            self._embed_cache[n] = hit
        return hit

    def encode_rgb(self, rgb_in: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pipeline:839-862: mean(quant_conv(encoder(x))) * rgb_latent_scale_factor -> [N,4,H/8,W/8] fp32."""
        lat = self.vae.encode_mean(rgb_in, scale=self.rgb_latent_scale_factor)
        if out is not None:
            out.copy_(lat)
            return out
        return lat

    def decode_seg(self, seg_latent: torch.Tensor) -> torch.Tensor:
        """pipeline:887-905: decoder(post_quant_conv(z / scale)).clip(-1, 1) -> [N,3,H,W] fp32."""
        N, _, hh, ww = seg_latent.shape
        rows = self.vae.decode_rows(seg_latent, in_scale=1.0 / self.seg_latent_scale_factor)
        return ops.nhwc_f32_to_nchw(rows, 3, hh * 8, ww * 8, lo=-1.0, hi=1.0)

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def single_infer(self, rgb_in_ref: torch.Tensor, rgb_in_tag: torch.Tensor, gt_in_ref: torch.Tensor,
                     clip_rgb_in=None, num_inference_steps: int = 1, show_pbar: bool = False, mode: str = "seg",
                     seed=None, _want_u8: bool = False, _want_f32: bool = True):
        """pipeline:616-836.  rgb_in_ref / gt_in_ref [B*k,3,H,W], rgb_in_tag [B,3,H,W] in [-1,1] ->
        seg [B,3,H,W] float in [0,255]."""
        if mode == "depth":
            raise NotImplementedError("mode='depth' is not on the DiffewS hot path")
        device = self.device
        self.scheduler.set_timesteps(num_inference_steps, device="cpu")          # :644
        timesteps = [int(x) for x in self.scheduler.timesteps]
        B, Bk = rgb_in_tag.shape[0], rgb_in_ref.shape[0]
        # :649-651, :674  — three VAE encodes; the support image / mask latents land in the two channel halves
        rgb_latent_ref = self.encode_rgb(rgb_in_ref.to(device))
        rgb_latent_tag = self.encode_rgb(rgb_in_tag.to(device))
        gt_latent_ref = self.encode_rgb(gt_in_ref.to(device))
        latents_rgb_cond_ref = torch.cat([rgb_latent_ref, gt_latent_ref], dim=1)
        depth_latent = rgb_latent_tag
        batch_embed = self._batch_embed(B)                                       # :680-692
        batch_embed_ref = self._batch_embed(Bk)
        unet_in = depth_latent
        for i, t in enumerate(timesteps):                                        # :706-767 (one iteration on the hot path)
            self.unet.clear_attn_bank()                                          # :715
            self.unet(latents_rgb_cond_ref, t * self.test_timestep, encoder_hidden_states=batch_embed_ref,
                      is_target=False)                                           # :719-720 support pass (fills banks)
            noise_pred = self.unet(depth_latent, t * self.test_timestep, encoder_hidden_states=batch_embed).sample
            self.unet.clear_attn_bank()                                          # :725
            last = i == len(timesteps) - 1
            if last and self.scheduler.is_pure_negation(t):
                z0, z_scale = noise_pred, -1.0 / self.seg_latent_scale_factor                          # z0 = -v, fused
                break
            # :764-769 scheduler.step(...): x_{t-1} for the next iteration, pred_original_sample after the last
            step_out = self.scheduler.step(noise_pred, t, depth_latent)
            if last:
                z0, z_scale = step_out.pred_original_sample, 1.0 / self.seg_latent_scale_factor
            else:
                unet_in = depth_latent = step_out.prev_sample.contiguous()
        # :787-795 decode_seg -> clip -> *0.5+0.5 -> *255 (and the uint8 truncation of :534): one fused head kernel
        seg_f32, seg_u8 = self.vae.decode_seg(z0, in_scale=z_scale, want_f32=_want_f32, want_u8=_want_u8)
        self._last_noise_pred = noise_pred
        self._last_unet_inputs = (latents_rgb_cond_ref, unet_in)
        return (seg_f32, seg_u8) if _want_u8 else seg_f32

    # ------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, input_images, denoising_steps: int = 10, ensemble_size: int = 10, processing_res: int = 768,
                 match_input_res: bool = True, batch_size: int = 0, color_map: str = "Spectral",
                 show_progress_bar: bool = True, ensemble_kwargs=None, mode: str = "depth", rgb_paths=(),
                 seed=None, output_type: str = "pil") -> MarigoldSegOutput:
        if mode == "depth":
            raise NotImplementedError("mode='depth' is not on the DiffewS hot path")
        assert processing_res >= 0 and denoising_steps >= 1 and ensemble_size >= 1       # :294-297
        if len(input_images) != 3:
            raise ValueError("input_images = [support_imgs, query_img, support_masks]  (main_oss.py:106-110)")
        ins = []
        for im in input_images:
            if not torch.is_tensor(im):
                im = self._pil_to_tensor(im, processing_res)
            im = im.to(self.device, torch.float32)
            if self.validate_inputs:
                assert float(im.min()) >= -1.0 and float(im.max()) <= 1.0                # :309
            ins.append(im)
        ref, tag, gt = ins
        input_size = tuple(tag.shape[-2:])
        if ensemble_size > 1:   # :376-383  stack x ensemble, fold into the batch
            ref, tag, gt = (x.repeat((ensemble_size, 1, 1, 1)) for x in (ref, tag, gt))
        need_f32 = ensemble_size > 1 or match_input_res      # the float image only feeds the ensemble mean / the resize
        seg_f32, seg_u8 = self.single_infer(ref, tag, gt, None, denoising_steps, False, mode, seed, _want_u8=True,
                                            _want_f32=need_f32)
        if ensemble_size > 1:   # :444-468 mean over the ensemble, then the uint8 truncation of :534
            bs = seg_f32.shape[0] // ensemble_size
            seg_f32 = seg_f32.view(ensemble_size, bs, *seg_f32.shape[1:]).mean(dim=0)
            seg_u8 = seg_f32.clip(0, 255).to(torch.uint8)
        if match_input_res and tuple(seg_u8.shape[-2:]) != input_size:                   # :473-474 (nearest)
            seg_f32 = torch.nn.functional.interpolate(seg_f32, input_size, mode="nearest")
            seg_u8 = seg_f32.clip(0, 255).to(torch.uint8)
        if output_type == "pt":
            return MarigoldSegOutput(seg_colored=seg_u8, uncertainty=None, seg_u8=seg_u8)
        from PIL import Image                                                            # :534-545
        arr = seg_u8.cpu().numpy()
        imgs = [Image.fromarray(np.ascontiguousarray(a.transpose(1, 2, 0))) for a in arr]
        return MarigoldSegOutput(seg_colored=imgs[0] if len(imgs) == 1 else imgs, uncertainty=None, seg_u8=seg_u8)

    @staticmethod
    def _pil_to_tensor(img, processing_res):
        """pipeline:326-349 for PIL inputs (resize_max_res, RGB, /255*2-1)."""
        if processing_res > 0:
            w, h = img.size
            s = min(processing_res / w, processing_res / h)
            img = img.resize((int(w * s), int(h * s)))
        a = np.asarray(img.convert("RGB")).transpose(2, 0, 1)
        return torch.from_numpy(a / 255.0 * 2.0 - 1.0).float()[None]
