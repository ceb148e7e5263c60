"""ORACLE (test infrastructure) — the episode data layer either side of the hot path.  CPU, numpy.

Follows:
  evaluation_util/data/dataset.py:36-40   transform = Resize((S,S)) -> ToTensor() -> Normalize([0.5],[0.5])
  evaluation_util/data/coco.py:32-58      __getitem__: transform(query), nearest mask resize, stacked supports
  evaluation_util/data/coco.py:84-115     load_frame: class-uniform episode sampling with np.random.choice
  evaluation_util/data/pascal.py:37-83    __getitem__ / extract_ignore_idx (boundary = floor(mask / 255))
  evaluation_util/data/pascal.py:101-110  sample_episode

PINNED — unlike the UNet/VAE oracle this part is checked against the real thing: the reference's own coco.py / pascal.py
import cleanly in the build container (torch, PIL, numpy, torchvision only), so scripts/make_golden_data.py runs the
UNMODIFIED reference datasets on a synthetic dataset tree and commits the outputs (tests/golden/data_layer.json).
The arithmetic of `Resize` lives in Pillow (ImagingResample, src/libImaging/Resample.c: two-pass separable convolution
with 22-bit fixed-point coefficients and a uint8 intermediate image), restated here from its published algorithm and
compared bit for bit with the installed Pillow in tests/test_data_layer.py; `F.interpolate(mode="nearest")` is torch's
floor(dst * float32(in / out)) rule (ATen UpSample.h nearest_neighbor_compute_source_index), compared with torch there.
"""
from __future__ import annotations

import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2          # Resample.c: coefficients are scaled by 2^22 for 8-bit channels


def pil_bilinear_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear ("triangle", support 1) filter over the
    full axis (box = 0 .. in_size).  Returns (xmin [out], xcount [out], kk [out, ksize] int32).  All arithmetic is IEEE
    double in the reference's order (no FMA contraction)."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    xmin = np.zeros(out_size, np.int32)
    xcnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx], xcnt[xx] = lo, n
    return xmin, xcnt, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass of ImagingResample along axis 0 of a uint8 array [in, ...] -> [out, ...] (uint8, clip8 of the rounded
    fixed-point sum: (sum + 2^21) >> 22 clamped to 0..255)."""
    xmin, xcnt, kk = pil_bilinear_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = xcnt[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[xmin[xx] + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear_u8(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`PIL.Image.resize((out_w, out_h), BILINEAR)` on an 8-bit image (torchvision Resize on a PIL input,
    dataset.py:37): horizontal pass first, then vertical, each only if that axis changes size (Resample.c
    ImagingResample need_horizontal / need_vertical)."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    h, w, _ = img_hwc.shape
    x = img_hwc
    if w != out_w:
        x = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(x.transpose(1, 0, 2)), out_w).transpose(1, 0, 2))
    if h != out_h:
        x = _resample_axis0(x, out_h)
    return np.ascontiguousarray(x)


def to_tensor_normalize(u8_hwc: np.ndarray, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """ToTensor() (uint8 HWC -> fp32 CHW, true division by 255) then Normalize([mean],[std]) (sub, then div), fp32."""
    t = torch.from_numpy(np.array(u8_hwc, copy=True)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    return t.sub(mean).div(std)


def transform_image(img_hwc: np.ndarray, size: int) -> torch.Tensor:
    """dataset.py:36-40 on one decoded RGB image."""
    return to_tensor_normalize(pil_resize_bilinear_u8(img_hwc, size, size))


def nearest_src_index(out_size: int, in_size: int) -> np.ndarray:
    """Source index of every destination index for F.interpolate(mode='nearest') on CPU (ATen UpSampleKernel.cpp
    nearest_idx): identity when sizes match, >> 1 for an exact 2x upscale, else min(floor(dst * float32(in/out)), in-1)
    with the product taken in fp32."""
    dst = np.arange(out_size, dtype=np.int64)
    if out_size == in_size:
        return dst
    if out_size == 2 * in_size:
        return dst >> 1
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.floor(dst.astype(np.float32) * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def nearest_resize(mask_hw: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    iy = nearest_src_index(out_h, mask_hw.shape[0])
    ix = nearest_src_index(out_w, mask_hw.shape[1])
    return mask_hw[iy][:, ix]


def coco_mask(label_hw: np.ndarray, class_sample: int, size: int) -> torch.Tensor:
    """coco.py:92-93 (label == class+1 -> 1 else 0) then :41-43 nearest resize -> float [S,S]."""
    m = (label_hw == class_sample + 1).astype(np.float32)
    return torch.from_numpy(nearest_resize(m, size, size).copy())


def pascal_mask(label_hw: np.ndarray, class_sample: int, size: int):
    """pascal.py:42-43 nearest resize of the raw class mask, then :78-83 extract_ignore_idx -> (mask, boundary)."""
    c = nearest_resize(label_hw.astype(np.float32), size, size)
    boundary = np.floor(c / 255)
    m = (c == class_sample + 1).astype(np.float32)
    return torch.from_numpy(m.copy()), torch.from_numpy(boundary.copy())


# ---- episode sampling ---------------------------------------------------------------------------------------------

def coco_class_ids(fold: int, split: str, nclass: int = 80, nfolds: int = 4):
    """coco.py:60-66."""
    val = [fold + nfolds * v for v in range(nclass // nfolds)]
    return val if split != "trn" else [x for x in range(nclass) if x not in val]


def pascal_class_ids(fold: int, split: str, nclass: int = 20, nfolds: int = 4):
    """pascal.py:112-120."""
    n = nclass // nfolds
    val = [fold * n + i for i in range(n)]
    return val if split != "trn" else [x for x in range(nclass) if x not in val]


def coco_sample_episode(class_ids, classwise: dict, shot: int):
    """coco.py:84-101: the exact np.random call sequence (global numpy RNG, as the reference uses it)."""
    class_sample = np.random.choice(class_ids, 1, replace=False)[0]
    query_name = np.random.choice(classwise[class_sample], 1, replace=False)[0]
    support_names = []
    while True:
        support_name = np.random.choice(classwise[class_sample], 1, replace=False)[0]
        if query_name != support_name:
            support_names.append(support_name)
        if len(support_names) == shot:
            break
    return query_name, support_names, class_sample


def pascal_sample_episode(img_metadata, classwise: dict, idx: int, shot: int):
    """pascal.py:38, :101-110."""
    idx %= len(img_metadata)
    query_name, class_sample = img_metadata[idx]
    support_names = []
    while True:
        support_name = np.random.choice(classwise[class_sample], 1, replace=False)[0]
        if query_name != support_name:
            support_names.append(support_name)
        if len(support_names) == shot:
            break
    return query_name, support_names, class_sample
