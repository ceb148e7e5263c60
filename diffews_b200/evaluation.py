"""Evaluator / AverageMeter — drop-ins for evaluation_util/common/evaluation.py and logger.py (reference), with the
counting done by the fused rthres + intersection/union kernel and kept in int64.

Reference semantics (evaluation.py:12-39): per episode, `inter = histc(pred[pred==gt], 2, 0, 1)`,
`union = histc(pred) + histc(gt) - inter`; PASCAL ignore pixels (value 255) fall outside the histogram range.
logger.py:30-51: class-indexed accumulation (`index_add_`) and mIoU / FB-IoU.  The reference accumulates pixel counts
in float32, which is inexact beyond 2^24 pixels per class; this path accumulates int64 and converts
to float only inside compute_iou().
"""
from __future__ import annotations

import torch

from . import ops

NCLASS = {"pascal": 20, "coco": 80, "fss": 1000, "paco_part": 448, "pascal_part": 100, "lvis": 1203}


class Evaluator:
    ignore_index = 255

    @classmethod
    def initialize(cls):
        cls.ignore_index = 255

    @staticmethod
    def _u8(t: torch.Tensor) -> torch.Tensor:
        return t if t.dtype == torch.uint8 else t.to(torch.uint8)

    @classmethod
    def classify_prediction(cls, pred_mask: torch.Tensor, batch: dict):
        """Reference signature: pred_mask [B,H,W] {0,1} (any dtype), batch['query_mask'] [B,H,W] {0,1},
        optional batch['query_ignore_idx'].  Returns (area_inter [2,B], area_union [2,B]) as float32 like the
        reference (values are exact integers)."""
        inter, union = cls.classify_prediction_counts(pred_mask, batch)
        return inter.t().float(), union.t().float()

    @classmethod
    def classify_prediction_counts(cls, pred_mask, batch):
        """Same, int64 [B,2] counts (bin 0 background, bin 1 foreground)."""
        gt = cls._u8(batch["query_mask"]).contiguous()
        ign = batch.get("query_ignore_idx")
        ign = cls._u8(ign).contiguous() if ign is not None else None
        inter, union, _ = ops.rthres_iou_hist(cls._u8(pred_mask).contiguous(), gt, ign, 0.0, want_mask=False)
        return inter, union

    @classmethod
    def rthres_classify(cls, pred_u8: torch.Tensor, batch: dict, r_threshold: float = 0.25, want_mask: bool = False):
        """Fused main_oss.py:128-134 + evaluation.py:12-39: pred_u8 [B,3,H,W] uint8 (the pipeline's seg output).
        The dynamic threshold uses the per-episode max.  Returns int64 [B,2] inter, [B,2] union (and the mask)."""
        gt = cls._u8(batch["query_mask"]).contiguous()
        ign = batch.get("query_ignore_idx")
        ign = cls._u8(ign).contiguous() if ign is not None else None
        inter, union, mask = ops.rthres_iou_hist(pred_u8.contiguous(), gt, ign, r_threshold, want_mask=want_mask)
        return (inter, union, mask) if want_mask else (inter, union)


class AverageMeter:
    """logger.py:10-51 with int64 device buffers.  `dataset` needs `.benchmark` and `.class_ids` like FSSDataset's
    datasets (evaluation_util/data/coco.py), or pass benchmark / class_ids directly."""

    def __init__(self, dataset=None, benchmark: str | None = None, class_ids=None, device="cuda"):
        if dataset is not None:
            benchmark, class_ids = dataset.benchmark, dataset.class_ids
        self.benchmark = benchmark
        self.nclass = NCLASS[benchmark]
        self.device = torch.device(device)
        self.class_ids_interest = torch.as_tensor(list(class_ids), dtype=torch.long, device=self.device)
        self.intersection_buf = torch.zeros([2, self.nclass], dtype=torch.int64, device=self.device)
        self.union_buf = torch.zeros([2, self.nclass], dtype=torch.int64, device=self.device)
        self.loss_buf = []

    def update_counts(self, inter_b2: torch.Tensor, union_b2: torch.Tensor, class_id: torch.Tensor):
        """int64 [B,2] counts straight from the kernel."""
        ops.iou_accumulate(inter_b2, union_b2, class_id.to(self.device, torch.int64).contiguous(),
                           self.intersection_buf, self.union_buf)

    def update(self, inter_b, union_b, class_id, loss=None):
        """Reference signature: inter_b / union_b are [2,B] (float) tensors."""
        self.update_counts(inter_b.t().round().to(torch.int64).contiguous(),
                           union_b.t().round().to(torch.int64).contiguous(), class_id)
        self.loss_buf.append(torch.tensor(0.0) if loss is None else loss)

    def reset(self):
        """Zero the counters IN PLACE (a captured CUDA graph holds the buffers' device pointers)."""
        self.intersection_buf.zero_()
        self.union_buf.zero_()
        self.loss_buf = []

    def all_reduce(self):
        """Data-parallel evaluation: sum the integer counts over ranks (NCCL, order-independent, bit-exact).
        The result is copied back INTO the existing buffers: EpisodeRunner's CUDA graph captured their device pointers
        (dfw_iou_accumulate), so the attributes must never be rebound to new tensors."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = torch.stack([self.intersection_buf, self.union_buf])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM)
            self.intersection_buf.copy_(packed[0])
            self.union_buf.copy_(packed[1])

    def compute_iou(self):
        inter = self.intersection_buf.float()
        union = self.union_buf.float()
        iou = inter / torch.max(torch.stack([union, torch.ones_like(union)]), dim=0)[0]
        iou = iou.index_select(1, self.class_ids_interest)
        miou = iou[1].mean() * 100
        fb_iou = (inter.index_select(1, self.class_ids_interest).sum(dim=1) /
                  union.index_select(1, self.class_ids_interest).sum(dim=1)).mean() * 100
        return miou, fb_iou, iou[1][: min(len(iou[1]), 20)]
