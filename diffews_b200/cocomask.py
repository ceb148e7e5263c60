"""Host-side decoding of COCO-style instance segmentations (LVIS-92i / PACO-Part annotations) to bitmasks.

The reference's LVIS dataset turns every annotation into a mask with `detectron2.structures.masks.polygons_to_bitmask`
(= pycocotools `frPyObjects` -> `merge` -> `decode`) for polygon lists, `pycocotools.mask.decode` for RLE dicts and
takes 2-D arrays as they are (evaluation_util/data/lvis.py:99-121).  Neither detectron2 nor pycocotools is installed
here and both are third-party (the reference pins neither), so the two algorithms are RESTATED from the published
pycocotools sources (common/maskApi.c: rleFrString, rleFrPoly, rleDecode, rleMerge) -- parity for polygon / RLE input is
therefore unpinned against the real library (tests check them against hand-verifiable shapes); the array path and
everything around it (sampling, union of instances, resizing) is pinned by the reference's own lvis.py
(tests/golden/data_layer.json, scripts/make_golden_data.py).  Pure numpy; runs in the decode thread pool.
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def rle_counts_from_string(s) -> List[int]:
    """maskApi.c rleFrString: COCO's LEB128-like compressed run lengths (6 bits per char, offset 48, delta-coded)."""
    if isinstance(s, bytes):
        s = s.decode("ascii")
    cnts, p, m = [], 0, 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if m > 2:
            x += cnts[m - 2]
        cnts.append(x)
        m += 1
    return cnts


def rle_decode(counts: Sequence[int], h: int, w: int) -> np.ndarray:
    """maskApi.c rleDecode: runs alternate 0 / 1 starting with 0, in COLUMN-major order -> uint8 [h, w]."""
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, v = 0, 0
    for c in counts:
        if v:
            flat[pos:pos + c] = 1
        pos += c
        v ^= 1
    return flat.reshape(w, h).T.copy()


def decode_rle_dict(segm: dict) -> np.ndarray:
    h, w = segm["size"]
    counts = segm["counts"]
    if not isinstance(counts, (list, tuple)):
        counts = rle_counts_from_string(counts)
    return rle_decode(counts, h, w)


def _c_round(v: float) -> int:
    return int(v + 0.5)            # (int)(x + .5) of maskApi.c: truncation toward zero after the shift


def polygon_to_rle_counts(xy: Sequence[float], h: int, w: int) -> List[int]:
    """maskApi.c rleFrPoly: the boundary is traced on a 5x up-sampled grid, the crossings of the pixel-column boundaries
    are collected and sorted into run lengths (column-major)."""
    k = len(xy) // 2
    scale = 5.0
    x = [_c_round(scale * xy[2 * j]) for j in range(k)]
    y = [_c_round(scale * xy[2 * j + 1]) for j in range(k)]
    x.append(x[0]); y.append(y[0])
    u, v = [], []
    for j in range(k):
        xs, xe, ys, ye = x[j], x[j + 1], y[j], y[j + 1]
        dx, dy = abs(xe - xs), abs(ys - ye)
        flip = (dx >= dy and xs > xe) or (dx < dy and ys > ye)
        if flip:
            xs, xe, ys, ye = xe, xs, ye, ys
        if dx >= dy:
            s = (ye - ys) / dx if dx else 0.0
            for d in range(dx + 1):
                t = dx - d if flip else d
                u.append(t + xs); v.append(_c_round(ys + s * t))
        else:
            s = (xe - xs) / dy
            for d in range(dy + 1):
                t = dy - d if flip else d
                v.append(t + ys); u.append(_c_round(xs + s * t))
    a = []
    for j in range(1, len(u)):
        if u[j] != u[j - 1]:
            xd = float(u[j] if u[j] < u[j - 1] else u[j] - 1)
            xd = (xd + 0.5) / scale - 0.5
            if math.floor(xd) != xd or xd < 0 or xd > w - 1:
                continue
            yd = float(v[j] if v[j] < v[j - 1] else v[j - 1])
            yd = (yd + 0.5) / scale - 0.5
            yd = 0.0 if yd < 0 else (float(h) if yd > h else yd)
            a.append(int(xd) * h + int(math.ceil(yd)))
    a.append(h * w)
    a.sort()
    prev, diffs = 0, []
    for t in a:
        diffs.append(t - prev)
        prev = t
    b, j = [diffs[0]], 1
    while j < len(diffs):
        if diffs[j] > 0:
            b.append(diffs[j]); j += 1
        else:
            j += 1
            if j < len(diffs):
                b[-1] += diffs[j]; j += 1
    return b


def polygons_to_bitmask(polygons: Sequence[Sequence[float]], h: int, w: int) -> np.ndarray:
    """detectron2 polygons_to_bitmask: union (rleMerge, intersect = 0) of the polygons' masks -> bool [h, w]."""
    if len(polygons) == 0:
        return np.zeros((h, w), dtype=bool)
    m = np.zeros((h, w), dtype=np.uint8)
    for p in polygons:
        m |= rle_decode(polygon_to_rle_counts(np.asarray(p, dtype=np.float64).reshape(-1).tolist(), h, w), h, w)
    return m.astype(bool)


def segmentation_to_mask(segm, h: int, w: int) -> np.ndarray:
    """lvis.py:99-121 `get_mask`: polygon list | RLE dict | 2-D array -> uint8 {0,1} [h, w]."""
    if isinstance(segm, list):
        return polygons_to_bitmask([np.asarray(p) for p in segm], h, w).astype(np.uint8)
    if isinstance(segm, dict):
        return decode_rle_dict(segm).astype(np.uint8)
    if isinstance(segm, np.ndarray):
        assert segm.ndim == 2, "Expect segmentation of 2 dimensions, got {}.".format(segm.ndim)
        return (segm != 0).astype(np.uint8) if segm.dtype == bool else segm.astype(np.uint8)
    raise NotImplementedError(type(segm))
