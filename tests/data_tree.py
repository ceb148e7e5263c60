"""Deterministic synthetic dataset trees in the on-disk layouts the reference's datasets read (COCO-20i, PASCAL-5i,
FSS-1000).  Shared by scripts/make_golden_data.py (which runs the UNMODIFIED reference datasets on them) and
tests/test_data_layer.py (which rebuilds the same trees).  Images are PNG-encoded (lossless, so the decoded pixels do
not depend on a JPEG library) under the `.jpg` names the reference expects — PIL sniffs the format from the content."""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
from PIL import Image


def _rand_image(rs: np.random.RandomState, h: int, w: int) -> np.ndarray:
    """Smooth gradients + noise + hard edges, so both the interpolation and the rounding paths are exercised."""
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 3), np.float64)
    for c in range(3):
        img[..., c] = 127 + 100 * np.sin(xx / rs.uniform(3, 20) + rs.uniform(0, 6)) * np.cos(yy / rs.uniform(3, 20))
    img += rs.uniform(-30, 30, img.shape)
    x0, y0 = rs.randint(0, w // 2), rs.randint(0, h // 2)
    img[y0:y0 + h // 3, x0:x0 + w // 3] = rs.randint(0, 2, 3) * 255
    return np.clip(img, 0, 255).astype(np.uint8)


def _rand_label(rs: np.random.RandomState, h: int, w: int, classes, boundary: bool) -> np.ndarray:
    lab = np.zeros((h, w), np.uint8)
    for c in classes:
        y0, x0 = rs.randint(0, h - 4), rs.randint(0, w - 4)
        y1, x1 = rs.randint(y0 + 3, h + 1), rs.randint(x0 + 3, w + 1)
        if boundary:
            lab[max(y0 - 1, 0):y1 + 1, max(x0 - 1, 0):x1 + 1] = 255
        lab[y0:y1, x0:x1] = c + 1
    return lab


def _save(arr: np.ndarray, path: str):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        Image.fromarray(arr).save(f, format="PNG")


def build_coco_tree(root: str, seed: int = 11, n_images: int = 30, fold: int = 0) -> str:
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "COCO2014")
    val_classes = [fold + 4 * v for v in range(20)]
    per_image = [[] for _ in range(n_images)]
    for i, c in enumerate(val_classes):                      # every class in >= 3 images
        for j in (i, (i * 7 + 3), (i * 11 + 5)):
            if c not in per_image[j % n_images]:
                per_image[j % n_images].append(c)
    classwise = {c: [] for c in val_classes}
    for i in range(n_images):
        h, w = rs.randint(40, 131), rs.randint(40, 131)
        if i % 6 == 0:
            h = w = 48                                        # same size as the target: the passes are skipped in PIL
        name = f"val2014/COCO_val2014_{i:012d}.jpg"
        _save(_rand_image(rs, h, w), os.path.join(base, name))
        _save(_rand_label(rs, h, w, per_image[i], False), os.path.join(base, "annotations", name[:-4] + ".png"))
        for c in per_image[i]:
            classwise[c].append(name)
    os.makedirs(os.path.join(base, "splits", "val"), exist_ok=True)
    with open(os.path.join(base, "splits", "val", f"fold{fold}.pkl"), "wb") as f:
        pickle.dump(classwise, f)
    return root


def build_pascal_tree(root: str, seed: int = 12, n_images: int = 16, fold: int = 0) -> str:
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "VOC2012")
    lines = []
    for i in range(n_images):
        h, w = rs.randint(36, 120), rs.randint(36, 120)
        name = f"2008_{i:06d}"
        classes = sorted({fold * 5 + i % 5, fold * 5 + (i // 2) % 5})
        _save(_rand_image(rs, h, w), os.path.join(base, "JPEGImages", name + ".jpg"))
        _save(_rand_label(rs, h, w, classes, True), os.path.join(base, "SegmentationClassAug", name + ".png"))
        lines += [f"{name}__{c + 1}" for c in classes]
    os.makedirs(os.path.join(base, "splits", "val"), exist_ok=True)
    with open(os.path.join(base, "splits", "val", f"fold{fold}.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return root


def build_fss_tree(root: str, seed: int = 13, categories=("abacus", "bat", "crt_screen")) -> str:
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "FSS-1000")
    for cat in categories:
        for i in range(1, 11):
            h, w = rs.randint(30, 100), rs.randint(30, 100)
            _save(_rand_image(rs, h, w), os.path.join(base, "data", cat, f"{i}.jpg"))
            m = rs.randint(0, 256, (h, w)).astype(np.uint8)                     # grey values around the 128 threshold
            m[h // 4:h // 2, w // 4:w // 2] = 255
            _save(np.stack([m, m, m], axis=-1), os.path.join(base, "data", cat, f"{i}.png"))   # RGB png -> convert('L')
    os.makedirs(os.path.join(base, "splits"), exist_ok=True)
    with open(os.path.join(base, "splits", "test.txt"), "w") as f:
        f.write("\n".join(categories) + "\n")
    return root


def metric_cases():
    """Seeded masks for Evaluator.classify_prediction (evaluation_util/common/evaluation.py:12-39): [B,H,W] float {0,1}
    prediction / ground truth, optional PASCAL ignore boundary.  Shared with tests/test_oracle.py (same generator)."""
    cases = []
    for seed, (B, H, W, ignore, kind) in enumerate([(1, 64, 64, False, "rand"), (3, 48, 80, False, "rand"),
                                                    (2, 64, 64, True, "rand"), (1, 32, 32, False, "all_bg"),
                                                    (1, 32, 32, False, "disjoint"), (2, 40, 40, True, "all_fg")]):
        g = torch.Generator().manual_seed(100 + seed)
        pred = (torch.rand(B, H, W, generator=g) > 0.6).float()
        gt = (torch.rand(B, H, W, generator=g) > 0.5).float()
        if kind == "all_bg":
            pred.zero_(); gt.zero_()
        elif kind == "disjoint":
            pred[:] = 1; gt.zero_()
        elif kind == "all_fg":
            pred[:] = 1; gt[:] = 1
        ign = None
        if ignore:
            ign = (torch.rand(B, H, W, generator=g) > 0.9).float()
            gt = gt * (1 - ign)                     # evaluation.py:17 asserts ignore and gt are disjoint
        cases.append({"seed": 100 + seed, "B": B, "H": H, "W": W, "ignore": ignore, "kind": kind, "pred": pred, "gt": gt,
                      "ign": ign})
    return cases


def attn_cases():
    """Seeded weights / inputs for one KV-bank self-attention layer (C = 128 = 2 heads x 64, 12 tokens): the reference
    protocol is clear bank -> support call on [B*k, S, C] (stores K, V) -> query call on [B, S, C] (attends to
    [self ; folded bank]).  Shared by scripts/make_golden_attn.py and tests/test_oracle.py."""
    cases = []
    for B, k in ((2, 1), (2, 3), (1, 5)):
        g = torch.Generator().manual_seed(1000 + 10 * B + k)
        C, S = 128, 12
        w = {n: torch.randn(C, C, generator=g) * C ** -0.5 for n in ("to_q", "to_k", "to_v", "to_out")}
        w["to_out_bias"] = torch.randn(C, generator=g) * 0.1
        cases.append({"B": B, "k": k, "C": C, "S": S, "heads": 2, "w": w,
                      "x_support": torch.randn(B * k, S, C, generator=g), "x_query": torch.randn(B, S, C, generator=g)})
    return cases
