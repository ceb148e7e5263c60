"""Batched, data-parallel evaluation loop — the B200 counterpart of `test_diffusion` (evaluation_util/main_oss.py:84-171).

One process per GPU; episodes are independent, so rank r takes episodes r, r+world, r+2*world, ... (no data-path
collective); the only exchange is one NCCL all-reduce of the int64 [2, nclass] intersection/union buffers at the end
(AverageMeter.all_reduce), after which every rank computes the same mIoU / FB-IoU.
"""
from __future__ import annotations

import torch

from .evaluation import NCLASS, AverageMeter, Evaluator


def shard_episodes(n_total: int, rank: int, world: int) -> range:
    """Episode indices owned by `rank` (round-robin, like a DistributedSampler without padding)."""
    return range(rank, n_total, world)


def build_engine_from_state_dicts(unet_sd, vae_sd, text_embeds, device="cuda", unet_precision=None, vae_precision=None,
                                  unet_channels=(320, 640, 1280, 1280), heads=(5, 10, 20, 20),
                                  vae_channels=(128, 256, 512, 512)):
    """Pipeline from diffusers-named state dicts (a checkpoint's tensors, or diffews_b200.synthetic.random_*_state_dict)."""
    from .pipeline import MarigoldPipelineRGBLatentNoise
    from .unet import MyUNet2DConditionModel
    from .vae import AutoencoderKL
    unet = MyUNet2DConditionModel(unet_sd, device=device, block_out_channels=unet_channels, heads=heads,
                                  precision=unet_precision)
    vae = AutoencoderKL(vae_sd, device=device, block_out_channels=vae_channels, precision=vae_precision)
    return MarigoldPipelineRGBLatentNoise(unet, vae, text_embeds=text_embeds)


def build_engine_from_modules(unet_module, vae_module, text_embeds, device="cuda", unet_precision=None,
                              vae_precision=None):
    from .pipeline import MarigoldPipelineRGBLatentNoise
    from .unet import MyUNet2DConditionModel
    from .vae import AutoencoderKL
    kw_u = {"precision": unet_precision} if unet_precision is not None else {}
    kw_v = {"precision": vae_precision} if vae_precision is not None else {}
    unet = MyUNet2DConditionModel.from_module(unet_module, device=device, **kw_u)
    vae = AutoencoderKL.from_module(vae_module, device=device, **kw_v)
    return MarigoldPipelineRGBLatentNoise(unet, vae, text_embeds=text_embeds)


class EpisodeRunner:
    def __init__(self, pipe, benchmark: str = "coco", class_ids=None, r_threshold: float = 0.25, img_size: int = 512):
        """`class_ids`: the classes mIoU is averaged over.  The reference builds `AverageMeter(dataloader.dataset)`
        (main_oss.py:87), i.e. the FOLD's classes (20 of 80 for COCO-20i): `run(dataloader)` does the same from
        `dataloader.dataset.class_ids`.  Leaving it None outside `run()` averages over every class of the benchmark,
        which is only right for synthetic episodes that draw from all of them (bench.py)."""
        self.pipe = pipe
        self.r_threshold = r_threshold
        self.img_size = img_size
        self.benchmark = benchmark
        self._class_ids_given = class_ids is not None
        nclass = NCLASS[benchmark]
        self.meter = AverageMeter(benchmark=benchmark, class_ids=class_ids if class_ids is not None else range(nclass),
                                  device=pipe.device)

    # ---- CUDA-graph mode: the ~1700 kernel launches of one batch become a single graph launch ------------------
    TENSOR_KEYS = ("query_img", "query_mask", "support_imgs", "support_masks", "class_id")
    OPTIONAL_KEYS = ("query_ignore_idx",)      # PASCAL boundary pixels (pascal.py:78-83 -> evaluation.py:16-21)

    @torch.no_grad()
    def enable_cuda_graph(self, example_batch: dict):
        """Capture one full step (VAE encodes, both UNet passes, decode, rthres + counts + accumulation) for the shapes
        of `example_batch`.  Later `step()` calls with the same shapes copy their inputs into the captured buffers and
        replay the graph; other shapes fall back to eager launches."""
        dev = self.pipe.device
        keys = self.TENSOR_KEYS + tuple(k for k in self.OPTIONAL_KEYS if example_batch.get(k) is not None)
        self._static = {k: example_batch[k].to(dev).clone() for k in keys}
        self.pipe.validate_inputs = False          # float(t.min()) would synchronise inside the capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = (self.meter.intersection_buf.clone(), self.meter.union_buf.clone())
        with torch.cuda.stream(side):
            for _ in range(2):                     # warm-up: lazy attribute sets, caches, allocator pools
                self._eager_step(self._static)
        torch.cuda.current_stream(dev).wait_stream(side)
        from . import ops
        self._graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self._graph):
            self._static_out = self._eager_step(self._static)
        self.launches_in_graph = ops.launch_count() - n0     # kernels of libdiffews_b200.so captured per replay
        self.meter.intersection_buf.copy_(saved[0])
        self.meter.union_buf.copy_(saved[1])
        self._graph_shapes = {k: tuple(v.shape) for k, v in self._static.items()}

    @torch.no_grad()
    def step(self, batch: dict):
        """`batch`: collated episode batch, on the device (main_oss.py:94 `utils.to_cuda(batch)`) or in pinned host
        memory.  Returns per-episode int64 (area_inter [B,2], area_union [B,2]) and updates the meter.
        In graph mode the returned tensors are the graph's static outputs: the next replay overwrites them, so clone
        them to keep per-batch results (eager mode returns fresh tensors).  The rthres threshold uses the PER-EPISODE
        maximum (the reference's `pred_mask.max()` is a whole-batch max, identical at its only supported bsz = 1)."""
        g = getattr(self, "_graph", None)
        if g is not None and all(k in batch and tuple(batch[k].shape) == s for k, s in self._graph_shapes.items()) \
                and not any(batch.get(k) is not None and k not in self._graph_shapes for k in self.OPTIONAL_KEYS):
            for k, dst in self._static.items():
                dst.copy_(batch[k], non_blocking=True)
            g.replay()
            return self._static_out
        dev = self.pipe.device
        return self._eager_step({k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v)
                                 for k, v in batch.items()})

    # ---- input prefetch: the H2D copy of batch i+1 overlaps the kernels of batch i ------------------------------------
    @torch.no_grad()
    def prefetch(self, batch: dict):
        """Start copying a (pinned host) batch of the captured shapes into a staging buffer on a copy stream.  The next
        `step_prefetched()` consumes it.  Call it right after launching a step: the copy then runs under that step."""
        assert getattr(self, "_graph", None) is not None, "enable_cuda_graph first"
        dev = self.pipe.device
        if not hasattr(self, "_stage"):
            self._stage = {k: torch.empty_like(v) for k, v in self._static.items()}
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_free = None
        cs = self._copy_stream
        if self._stage_free is not None:
            cs.wait_event(self._stage_free)          # the previous step has moved the staging buffer into the graph's inputs
        with torch.cuda.stream(cs):
            for k, dst in self._stage.items():
                dst.copy_(batch[k], non_blocking=True)
            self._stage_ready = torch.cuda.Event()
            self._stage_ready.record(cs)

    @torch.no_grad()
    def step_prefetched(self):
        """Run one step on the batch handed to the last `prefetch()` (device-to-device move into the captured input
        buffers, then the graph).  Same return value as `step()`."""
        cur = torch.cuda.current_stream(self.pipe.device)
        cur.wait_event(self._stage_ready)
        for k, dst in self._static.items():
            dst.copy_(self._stage[k], non_blocking=True)
        self._stage_free = torch.cuda.Event()
        self._stage_free.record(cur)
        self._graph.replay()
        return self._static_out

    @torch.no_grad()
    def _eager_step(self, batch: dict):
        query_img, query_mask = batch["query_img"], batch["query_mask"]
        support_imgs, support_masks = batch["support_imgs"], batch["support_masks"]
        # main_oss.py:99-104: masks [b,k,h,w] -> [b,k,3,h,w] in [-1,1]; shots folded into the batch dim
        support_masks = support_masks.unsqueeze(2).repeat(1, 1, 3, 1, 1) * 2 - 1
        support_imgs = support_imgs.reshape(-1, *support_imgs.shape[-3:])
        support_masks = support_masks.reshape(-1, *support_masks.shape[-3:])
        out = self.pipe([support_imgs, query_img, support_masks], denoising_steps=1, ensemble_size=1,
                        processing_res=self.img_size, batch_size=query_img.shape[0], show_progress_bar=False,
                        mode="seg", rgb_paths=batch.get("rgb_path", []), seed=0, output_type="pt")
        inter, union = Evaluator.rthres_classify(out.seg_u8, batch, self.r_threshold)
        self.meter.update_counts(inter, union, batch["class_id"])
        self.last_seg_u8 = out.seg_u8      # uint8 [B,3,H,W] of this step (a static buffer of the graph in graph mode)
        return inter, union

    def finish(self):
        self.meter.all_reduce()
        return self.meter.compute_iou()

    @torch.no_grad()
    def run(self, dataloader, max_batches=None, use_cuda_graph: bool = True, log_every: int = 0):
        """`test_diffusion(pipe, dataloader, args)` (evaluation_util/main_oss.py:84-171): every batch of the loader
        through the pipeline, rthres, intersection/union, class accumulation; returns (mIoU, FB-IoU) after the
        cross-rank all-reduce.  `dataloader` yields the reference's batch dicts (diffews_b200.data.EpisodeLoader, or the
        reference's own DataLoader).  Full batches replay one CUDA graph; the short last batch runs eagerly."""
        ds = getattr(dataloader, "dataset", None)
        if ds is not None and hasattr(ds, "class_ids"):
            # main_oss.py:87 `AverageMeter(dataloader.dataset)`: mIoU averages over the fold's classes only
            ids = torch.as_tensor(list(ds.class_ids), dtype=torch.long, device=self.meter.device)
            if self._class_ids_given and not torch.equal(ids, self.meter.class_ids_interest):
                raise ValueError("EpisodeRunner class_ids differ from dataloader.dataset.class_ids")
            if getattr(ds, "benchmark", self.benchmark) != self.benchmark:
                raise ValueError(f"dataset benchmark {ds.benchmark!r} != runner benchmark {self.benchmark!r}")
            self.meter.class_ids_interest = ids
        for i, batch in enumerate(dataloader):
            if max_batches is not None and i >= max_batches:
                break
            if use_cuda_graph and getattr(self, "_graph", None) is None and i == 0:
                self.enable_cuda_graph(batch)
            self.step(batch)
            if log_every and (i + 1) % log_every == 0:
                miou, fb, _ = self.meter.compute_iou()
                print(f"[batch {i + 1}] mIoU {float(miou):.2f}  FB-IoU {float(fb):.2f}", flush=True)
        miou, fb_iou, _ = self.finish()
        return float(miou), float(fb_iou)
