"""Golden vectors for the pipeline wiring, produced by the reference's OWN
`MarigoldPipelineRGBLatentNoise.single_infer` / `encode_rgb` / `decode_seg`.

diffews/marigold_pipeline_rgb_latent_noise.py is executed UNMODIFIED.  What it imports but the container lacks is
replaced by stand-ins that carry no arithmetic of the path:
  * diffusers.DiffusionPipeline -> a base class whose register_modules() sets attributes; UNet2DConditionModel /
    AutoencoderKL / ControlNetModel / BaseOutput / image_processor names -> placeholders (type annotations, isinstance)
  * matplotlib, marigold.models, marigold.image_projector -> placeholders (unused on the seg path); the `marigold`
    package is registered by path so marigold/util/{image_util,batchsize,ensemble}.py are the reference's real files
    and marigold/__init__.py (which pulls the legacy pipelines) is not run
The sub-modules the pipeline DRIVES are plugged in from this repo: the oracle UNet / VAE (oracle/sd21.py, reduced
width, seeded) behind a thin adapter that adds `.sample` / `.device`, and the restated DDIM scheduler
(diffews_b200/scheduler.py, whose tables are pinned against the reference scheduler file by make_golden_attn.py).
So the golden pins exactly what the reference file itself contributes: order of the three encodes, the scale factors,
the support/query channel concat, the embed repeats, the clear / support pass / query pass / clear protocol with
`t * test_timestep`, `pred_original_sample`, decode, clip, * 0.5 + 0.5, * 255.

Output: tests/golden/pipeline_reference.json (sub-sampled seg values + statistics for 1-shot B=2 and 3-shot B=1).
    python scripts/make_golden_pipeline.py
"""
import base64
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class DiffusionPipeline:
        device = torch.device("cpu")                 # properties of the real base class that __call__ reads
        dtype = torch.float32

        def __init__(self):
            pass

        def register_modules(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

    class BaseOutput:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    ph = lambda n: type(n, (), {})                                                      # noqa: E731
    mod("diffusers", DiffusionPipeline=DiffusionPipeline, UNet2DConditionModel=ph("UNet2DConditionModel"),
        AutoencoderKL=ph("AutoencoderKL"), ControlNetModel=ph("ControlNetModel"), DDIMScheduler=ph("DDIMScheduler"),
        DDPMScheduler=ph("DDPMScheduler"))
    mod("diffusers.utils", BaseOutput=BaseOutput)
    mod("diffusers.utils.torch_utils", randn_tensor=torch.randn)
    mod("diffusers.image_processor", PipelineImageInput=object, VaeImageProcessor=ph("VaeImageProcessor"))
    mod("diffusers.configuration_utils", ConfigMixin=object, register_to_config=lambda f: f)
    mod("diffusers.schedulers")
    mod("diffusers.schedulers.scheduling_ddim", DDIMSchedulerOutput=dict)
    plt = mod("matplotlib.pyplot")
    mod("matplotlib", pyplot=plt, colormaps={})
    pkg = mod("marigold")
    pkg.__path__ = [os.path.join(REF, "marigold")]                  # real sub-modules, without running marigold/__init__.py
    util = mod("marigold.util")
    util.__path__ = [os.path.join(REF, "marigold", "util")]
    mod("marigold.image_projector", ImageProjModel=ph("ImageProjModel"))
    mod("marigold.models", DPTHead=ph("DPTHead"), CustomUNet2DConditionModel=ph("CustomUNet2DConditionModel"))


class UNetAdapter:
    """oracle UNet (returns a tensor) behind the diffusers call protocol the pipeline uses (`.sample`)."""

    def __init__(self, unet):
        self.unet = unet

    def clear_attn_bank(self):
        self.unet.clear_attn_bank()

    def __call__(self, sample, timestep, encoder_hidden_states=None, is_target=True):
        return types.SimpleNamespace(sample=self.unet(sample, timestep, encoder_hidden_states, is_target=is_target))


def subsample(seg: torch.Tensor) -> dict:
    sub = seg[:, :, ::4, ::4].contiguous().to(torch.float32)
    return {"shape": list(seg.shape), "mean": float(seg.double().mean()), "abs_mean": float(seg.double().abs().mean()),
            "min": float(seg.min()), "max": float(seg.max()), "sub_shape": list(sub.shape),
            "sub_f32_b64": base64.b64encode(sub.numpy().tobytes()).decode()}


def cases():
    """(B, k, image size): shared with tests/test_oracle.py."""
    return [(2, 1, 64), (1, 3, 64)]


def main():
    install_stubs()
    spec = importlib.util.spec_from_file_location("ref_pipeline", os.path.join(REF, "diffews", "marigold_pipeline_rgb_latent_noise.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from diffews_b200.scheduler import DDIMSchedulerCustomized
    from diffews_b200.synthetic import make_batch, pipeline_inputs, prompt_embedding
    from oracle.sd21 import build_models
    torch.set_num_threads(8)
    unet, vae = build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    vae.device = torch.device("cpu")                                 # pipeline:771 `depth_latent.to(self.vae.device)`
    pipe = ref.MarigoldPipelineRGBLatentNoise(unet=UNetAdapter(unet), vae=vae, scheduler=DDIMSchedulerCustomized(),
                                              tokenizer=None, text_embeds=prompt_embedding(), text_encoder=object(),
                                              image_encoder=None, image_projector=None, controlnet=None,
                                              customized_head=None)
    pipe.test_timestep = 1                                           # main_oss.py:373
    out = {"made_by": "scripts/make_golden_pipeline.py: unmodified MarigoldPipelineRGBLatentNoise.single_infer driving the "
                      "oracle UNet / VAE (reduced width, seed 0) and the restated scheduler", "cases": []}
    for B, k, size in cases():
        ref_imgs, tag, gt = pipeline_inputs(make_batch(0, B, size, k))
        with torch.no_grad():
            seg = pipe.single_infer(ref_imgs, tag, gt, None, 1, False, mode="seg", seed=0)
            lat = pipe.encode_rgb(tag)
        rec = {"B": B, "k": k, "size": size, "seg": subsample(seg), "encode_rgb_tag_mean": float(lat.double().mean()),
               "encode_rgb_tag_abs_mean": float(lat.double().abs().mean())}
        out["cases"].append(rec)
        print(B, k, rec["seg"]["mean"], rec["seg"]["min"], rec["seg"]["max"])
    # __call__ as test_diffusion makes it (main_oss.py:113-123; bsz = 1 is the only batch size the reference eval supports):
    # tensor inputs, a real file in rgb_paths (opened for CLIP image features that the text-embed mode never uses).
    import tempfile
    from PIL import Image
    B, k, size = 1, 1, 64
    ref_imgs, tag, gt = pipeline_inputs(make_batch(0, B, size, k))
    with tempfile.TemporaryDirectory() as d:
        fn = os.path.join(d, "q.png")
        Image.fromarray(np.zeros((size, size, 3), np.uint8)).save(fn)
        res = pipe([ref_imgs, tag, gt], denoising_steps=1, ensemble_size=1, processing_res=size, batch_size=1,
                   show_progress_bar=False, mode="seg", rgb_paths=[fn], seed=0)
    img = np.asarray(res.seg_colored)
    out["call"] = {"B": B, "k": k, "size": size, "type": type(res.seg_colored).__name__, "shape": list(img.shape),
                   "dtype": str(img.dtype), "uncertainty_is_none": res.uncertainty is None,
                   "u8_b64": base64.b64encode(np.ascontiguousarray(img[::2, ::2]).tobytes()).decode(),
                   "sum": int(img.astype(np.int64).sum())}
    print("__call__:", out["call"]["type"], out["call"]["shape"], out["call"]["sum"])
    path = os.path.join(ROOT, "tests", "golden", "pipeline_reference.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
