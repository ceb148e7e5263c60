"""Deterministic synthetic episodes with the reference's batch layout (evaluation_util/data/coco.py:49-60):
query_img [3,S,S], query_mask [S,S], support_imgs [k,3,S,S], support_masks [k,S,S], class_id.  SURVEY §8(d)."""
from __future__ import annotations

import torch


def _ellipse_mask(g: torch.Generator, size: int) -> torch.Tensor:
    yy, xx = torch.meshgrid(torch.arange(size, dtype=torch.float32), torch.arange(size, dtype=torch.float32),
                            indexing="ij")
    mask = torch.zeros(size, size, dtype=torch.bool)
    n = int(torch.randint(1, 4, (1,), generator=g))
    for _ in range(n):
        cy, cx = (torch.rand(2, generator=g) * 0.6 + 0.2) * size
        ry, rx = (torch.rand(2, generator=g) * 0.25 + 0.08) * size
        mask |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
    return mask.to(torch.float32)


def make_episode(idx: int, size: int = 512, nshot: int = 1, nclass: int = 80) -> dict:
    g = torch.Generator().manual_seed(1234 + idx)
    ep = {
        "query_img": torch.rand(3, size, size, generator=g) * 2 - 1,
        "support_imgs": torch.rand(nshot, 3, size, size, generator=g) * 2 - 1,
        "query_mask": _ellipse_mask(g, size),
        "support_masks": torch.stack([_ellipse_mask(g, size) for _ in range(nshot)]),
        "class_id": torch.randint(0, nclass, (1,), generator=g)[0],
    }
    return ep


def make_batch(start: int, B: int, size: int = 512, nshot: int = 1, nclass: int = 80) -> dict:
    """Collated like a DataLoader batch: query_img [B,3,S,S], query_mask [B,S,S], support_imgs [B,k,3,S,S],
    support_masks [B,k,S,S], class_id [B]."""
    eps = [make_episode(start + i, size, nshot, nclass) for i in range(B)]
    return {k: torch.stack([e[k] for e in eps]) for k in eps[0]}


def pipeline_inputs(batch: dict):
    """evaluation_util/main_oss.py:99-110: masks -> 3 channels in [-1,1]; shots folded into the batch dim."""
    sm = batch["support_masks"].unsqueeze(2).repeat(1, 1, 3, 1, 1) * 2 - 1
    si = batch["support_imgs"]
    si = si.reshape(-1, *si.shape[-3:])
    sm = sm.reshape(-1, *sm.shape[-3:])
    return [si, batch["query_img"], sm]


def prompt_embedding(lctx: int = 2, dim: int = 1024) -> torch.Tensor:
    return torch.randn(1, lctx, dim, generator=torch.Generator().manual_seed(7))
