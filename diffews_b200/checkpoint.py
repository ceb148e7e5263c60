"""Real-checkpoint loading: the diffusers directory layout the reference loads in
evaluation_util/main_oss.py:338-369 —

    <ckpt>/unet/diffusion_pytorch_model.safetensors   (+ config.json)     UNet2DConditionModel keys + conv_in_ref.*
    <ckpt>/vae/diffusion_pytorch_model.safetensors    (+ config.json)     AutoencoderKL keys
    <scheduler_load_path>/scheduler_config.json                           DDIM config (scheduler_1.0_1.0)

The engines take diffusers-named state dicts verbatim, so loading is "read the tensors, read the few config fields the
engines need, hand them over".  `.bin` (torch.save) shards are accepted too.  Nothing here imports diffusers.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

_WEIGHT_NAMES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
                 "diffusion_pytorch_model.bin")


def load_state_dict(folder_or_file: str) -> Dict[str, torch.Tensor]:
    """State dict of one diffusers sub-model (a folder holding diffusion_pytorch_model.* or a weight file)."""
    path = folder_or_file
    if os.path.isdir(path):
        for name in _WEIGHT_NAMES:
            if os.path.exists(os.path.join(path, name)):
                path = os.path.join(path, name)
                break
        else:
            raise FileNotFoundError(f"no {_WEIGHT_NAMES} under {folder_or_file}")
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path, device="cpu")
    sd = torch.load(path, map_location="cpu", weights_only=True)
    return sd.get("state_dict", sd)


def load_config(folder: str) -> dict:
    p = os.path.join(folder, "config.json")
    if not os.path.exists(p):
        return {}
    with open(p) as f:
        return json.load(f)


def _heads_from_config(cfg: dict, block_out_channels):
    """diffusers stores the head COUNT of SD-2.x under the (mis-named) `attention_head_dim` key."""
    h = cfg.get("num_attention_heads") or cfg.get("attention_head_dim")
    if h is None:
        return tuple(c // 64 for c in block_out_channels)
    if isinstance(h, int):
        return (h,) * len(block_out_channels)
    return tuple(h)


def with_support_stem(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """A plain SD-2.1 UNet has no 8-channel support stem: initialise `conv_in_ref` from `conv_in` exactly as
    train_tools/load_ckpt_and_modify_ref8in_tag4in.py:6-28 does (weights repeated over the two inputs and halved).
    A DiffewS checkpoint already carries it and is returned unchanged."""
    if "conv_in_ref.weight" in sd:
        return sd
    sd = dict(sd)
    sd["conv_in_ref.weight"] = sd["conv_in.weight"].repeat(1, 2, 1, 1) / 2
    sd["conv_in_ref.bias"] = sd["conv_in.bias"].clone()
    return sd


def load_unet(ckpt_dir: str, device="cuda", precision=None, subfolder: Optional[str] = "unet"):
    """diffews/models/unet_2d_condition.py `MyUNet2DConditionModel.from_pretrained(ckpt, subfolder='unet')` equivalent."""
    from .unet import MyUNet2DConditionModel
    folder = os.path.join(ckpt_dir, subfolder) if subfolder else ckpt_dir
    sd = load_state_dict(folder)
    cfg = load_config(folder)
    boc = tuple(cfg.get("block_out_channels", (320, 640, 1280, 1280)))
    sd = with_support_stem(sd)
    return MyUNet2DConditionModel(sd, device=device, block_out_channels=boc, heads=_heads_from_config(cfg, boc),
                                  cross_attention_dim=cfg.get("cross_attention_dim", 1024), precision=precision)


def load_vae(ckpt_dir: str, device="cuda", precision=None, subfolder: Optional[str] = "vae"):
    from .vae import AutoencoderKL
    folder = os.path.join(ckpt_dir, subfolder) if subfolder else ckpt_dir
    sd = load_state_dict(folder)
    cfg = load_config(folder)
    return AutoencoderKL(sd, device=device, block_out_channels=tuple(cfg.get("block_out_channels", (128, 256, 512, 512))),
                         precision=precision)


def load_pipeline(ckpt_dir: str, scheduler_dir: str, text_embeds: torch.Tensor, device="cuda"):
    """The objects main_oss.py:339-373 builds, from a released DiffewS checkpoint directory.  `text_embeds` is the
    [1, Lctx, 1024] empty-prompt embedding (the CLIP text encoder itself is outside the path)."""
    from .pipeline import MarigoldPipelineRGBLatentNoise
    from .scheduler import DDIMSchedulerCustomized
    pipe = MarigoldPipelineRGBLatentNoise(
        unet=load_unet(ckpt_dir, device), vae=load_vae(ckpt_dir, device),
        scheduler=DDIMSchedulerCustomized.from_config_file(os.path.join(scheduler_dir, "scheduler_config.json")),
        text_embeds=text_embeds)
    pipe.test_timestep = 1
    return pipe
