"""ORACLE (test infrastructure) — rthres binarisation, intersection/union, AverageMeter.  CPU, plain torch.

Follows:
  evaluation_util/main_oss.py:128-137      to_tensor(PIL) ; dynamic_thres = pred.max()*r ; pred.mean(dim=1) > thres
  evaluation_util/common/evaluation.py:12-39   Evaluator.classify_prediction
  evaluation_util/common/logger.py:10-51       AverageMeter.update / compute_iou
"""
from __future__ import annotations

import torch


def to_tensor_u8(pred_u8: torch.Tensor) -> torch.Tensor:
    """torchvision.transforms.functional.to_tensor on a uint8 image: fp32, value / 255 (true division).
    main_oss.py:128.  Input here is already [B,3,H,W] uint8 (the PIL HWC->CHW permute is layout only)."""
    return pred_u8.to(torch.float32).div(255)


def rthres_mask(pred_u8: torch.Tensor, r_threshold: float = 0.25) -> torch.Tensor:
    """main_oss.py:131-134 for ONE episode ([1,3,H,W] uint8) -> [1,H,W] float {0,1}.
    `pred.max()` is a whole-tensor max, which is per-episode at the only batch size the reference supports (1)."""
    assert pred_u8.shape[0] == 1
    pred = to_tensor_u8(pred_u8)
    dynamic_thres = pred.max() * r_threshold
    return (pred.mean(dim=1) > dynamic_thres).to(torch.float32)


def classify_prediction(pred_mask: torch.Tensor, batch: dict, ignore_index: int = 255):
    """evaluation.py:12-39, verbatim semantics (float masks, torch.histc with 2 bins over [0,1])."""
    gt_mask = batch.get("query_mask")
    query_ignore_idx = batch.get("query_ignore_idx")
    if query_ignore_idx is not None:
        assert torch.logical_and(query_ignore_idx, gt_mask).sum() == 0
        query_ignore_idx = query_ignore_idx * ignore_index
        gt_mask = gt_mask + query_ignore_idx
        pred_mask[gt_mask == ignore_index] = ignore_index
    area_inter, area_pred, area_gt = [], [], []
    for _pred_mask, _gt_mask in zip(pred_mask, gt_mask):
        _inter = _pred_mask[_pred_mask == _gt_mask]
        if _inter.size(0) == 0:
            _area_inter = torch.tensor([0, 0], device=_pred_mask.device, dtype=torch.float32)
        else:
            _area_inter = torch.histc(_inter, bins=2, min=0, max=1)
        area_inter.append(_area_inter)
        area_pred.append(torch.histc(_pred_mask, bins=2, min=0, max=1))
        area_gt.append(torch.histc(_gt_mask, bins=2, min=0, max=1))
    area_inter = torch.stack(area_inter).t()
    area_pred = torch.stack(area_pred).t()
    area_gt = torch.stack(area_gt).t()
    area_union = area_pred + area_gt - area_inter
    return area_inter, area_union


NCLASS = {"pascal": 20, "coco": 80, "fss": 1000, "paco_part": 448, "pascal_part": 100, "lvis": 1203}


class AverageMeter:
    """logger.py:10-51 on CPU.  `exact=True` accumulates in int64 (what the B200 path does); `exact=False`
    reproduces the reference's float32 buffers (inexact beyond 2^24 pixels per class — SURVEY Appendix A)."""

    def __init__(self, benchmark: str, class_ids, exact: bool = True):
        self.benchmark = benchmark
        self.class_ids_interest = torch.as_tensor(list(class_ids), dtype=torch.long)
        self.nclass = NCLASS[benchmark]
        dt = torch.int64 if exact else torch.float32
        self.intersection_buf = torch.zeros([2, self.nclass], dtype=dt)
        self.union_buf = torch.zeros([2, self.nclass], dtype=dt)
        self.loss_buf = []

    def update(self, inter_b, union_b, class_id, loss=None):
        self.intersection_buf.index_add_(1, class_id, inter_b.to(self.intersection_buf.dtype))
        self.union_buf.index_add_(1, class_id, union_b.to(self.union_buf.dtype))
        self.loss_buf.append(torch.tensor(0.0) if loss is None else loss)

    def compute_iou(self):
        inter = self.intersection_buf.float()
        union = self.union_buf.float()
        iou = inter / torch.max(torch.stack([union, torch.ones_like(union)]), dim=0)[0]
        iou = iou.index_select(1, self.class_ids_interest)
        miou = iou[1].mean() * 100
        fb_iou = (inter.index_select(1, self.class_ids_interest).sum(dim=1) /
                  union.index_select(1, self.class_ids_interest).sum(dim=1)).mean() * 100
        return miou, fb_iou, iou[1][: min(len(iou[1]), 20)]
