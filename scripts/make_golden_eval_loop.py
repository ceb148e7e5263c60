"""End-to-end golden: the reference's OWN evaluation loop, `test_diffusion(pipe, dataloader, args)`
(evaluation_util/main_oss.py:84-171), executed UNMODIFIED on the synthetic COCO-20i tree.

Everything on the path that is the reference's own code runs as written: evaluation_util/main_oss.py (input folding
:99-110, the pipeline call :113-123, the inline rthres :125-137), evaluation_util/data/{dataset,coco}.py with
torchvision's transform, diffews/marigold_pipeline_rgb_latent_noise.py (__call__, single_infer, encode / decode),
evaluation_util/common/{evaluation,logger,utils,vis}.py.  Stand-ins (no arithmetic of the path): the diffusers /
accelerate / detectron2 / pycocotools / cv2 / matplotlib names those files import, `torch.Tensor.cuda` -> identity (no
GPU here).  Plugged in from this repo, as in make_golden_pipeline.py: the oracle UNet / VAE (reduced width, seed 0)
and the restated DDIM scheduler.  The loader is capped to the first N episodes by a proxy object (the reference's
COCO dataset reports 1000 episodes).

Output: tests/golden/eval_loop_reference.json — per-episode area_inter / area_union as returned by the reference
Evaluator inside the loop (recorded by wrapping the classmethod from outside), mIoU and FB-IoU returned by
test_diffusion.      python scripts/make_golden_eval_loop.py
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import data_tree  # noqa: E402
import make_golden_pipeline as mgp  # noqa: E402

REF = "/root/reference"
N_EPISODES, IMG = 6, 64


def install_more_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    ph = lambda n: type(n, (), {})                                                      # noqa: E731
    sys.modules["diffusers.schedulers"].DDPMScheduler = ph("DDPMScheduler")
    mod("accelerate", Accelerator=ph("Accelerator"))
    mod("accelerate.utils", ProjectConfiguration=ph("ProjectConfiguration"), set_seed=lambda s: None)
    mod("detectron2")
    mod("detectron2.structures")
    mod("detectron2.structures.masks")
    mod("pycocotools")
    sys.modules["pycocotools"].mask = mod("pycocotools.mask")
    mod("cv2")
    pkg = mod("diffews")
    pkg.__path__ = [os.path.join(REF, "diffews")]
    mod("diffews.models")
    mod("diffews.models.unet_2d_condition", MyUNet2DConditionModel=ph("MyUNet2DConditionModel"))


class FirstN:
    """The caller's choice of loader: the first n batches of the reference DataLoader (same dataset object)."""

    def __init__(self, loader, n):
        self.loader, self.n, self.dataset = loader, n, loader.dataset

    def __len__(self):
        return self.n

    def __iter__(self):
        for i, b in enumerate(self.loader):
            if i >= self.n:
                return
            yield b


def main():
    # transformers is installed and real; import what the reference files take from it BEFORE the stand-ins exist, so its
    # own optional-dependency probes (accelerate, ...) see the true environment
    from transformers import CLIPImageProcessor, CLIPTextModel, CLIPTokenizer, CLIPVisionModelWithProjection  # noqa: F401
    mgp.install_stubs()
    install_more_stubs()
    sys.path.insert(0, REF)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        main_oss = importlib.import_module("evaluation_util.main_oss")      # runs random.seed(0), fix_randseed(0)
        from diffews_b200.scheduler import DDIMSchedulerCustomized
        from diffews_b200.synthetic import prompt_embedding
        from oracle.sd21 import build_models
        torch.set_num_threads(8)
        unet, vae = build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
        vae.device = torch.device("cpu")
        pipe = main_oss.MarigoldPipeline(unet=mgp.UNetAdapter(unet), vae=vae, scheduler=DDIMSchedulerCustomized(),
                                         tokenizer=None, text_embeds=prompt_embedding(), text_encoder=object(),
                                         image_encoder=None, image_projector=None, controlnet=None, customized_head=None)
        pipe.test_timestep = 1                                               # main_oss.py:373
        args = argparse.Namespace(denoise_steps=1, ensemble_size=1, img_size=IMG, bsz=1, r_threshold=0.25, threshold=0,
                                  benchmark="coco")
        recorded = []
        orig_cp = main_oss.Evaluator.classify_prediction.__func__

        def spy(cls, pred_mask, batch):
            inter, union = orig_cp(cls, pred_mask, batch)
            recorded.append({"class_id": int(batch["class_id"][0]), "query_name": batch["query_name"][0],
                             "support_names": [s[0] for s in batch["support_names"]],
                             "pred_fg": int(pred_mask.sum()), "area_inter": inter.tolist(), "area_union": union.tolist()})
            return inter, union
        main_oss.Evaluator.classify_prediction = classmethod(spy)
        with tempfile.TemporaryDirectory() as tree:
            data_tree.build_coco_tree(tree)
            main_oss.FSSDataset.initialize(img_size=IMG, datapath=tree, use_original_imgsize=False)   # main_oss.py:395
            main_oss.Evaluator.initialize()
            main_oss.Visualizer.initialize(False)
            np.random.seed(0)                                                # = utils.fix_randseed(0) at import, restated
            loader = main_oss.FSSDataset.build_dataloader("coco", 1, 0, 0, "test", 1)                 # main_oss.py:396-399
            with torch.no_grad():
                miou, fb_iou = main_oss.test_diffusion(pipe, FirstN(loader, N_EPISODES), args)
    finally:
        torch.Tensor.cuda = orig_cuda
    out = {"made_by": "scripts/make_golden_eval_loop.py: unmodified evaluation_util/main_oss.py test_diffusion on the synthetic "
                      "COCO-20i tree, oracle UNet / VAE + restated scheduler plugged in",
           "img_size": IMG, "episodes": recorded, "miou": float(miou), "fb_iou": float(fb_iou)}
    path = os.path.join(ROOT, "tests", "golden", "eval_loop_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("mIoU", float(miou), "FB-IoU", float(fb_iou), [e["area_inter"] for e in recorded][:3])
    print("wrote", path)


if __name__ == "__main__":
    main()
