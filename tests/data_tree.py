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


# ---- instance-segmentation benchmarks (LVIS-92i, PACO-Part, PASCAL-Part) --------------------------------------------
def rle_encode(mask_hw: np.ndarray):
    """Uncompressed COCO RLE of a binary mask: column-major run lengths starting with a run of zeros."""
    flat = np.asarray(mask_hw, dtype=np.uint8).T.reshape(-1)
    counts, prev, run = [], 0, 0
    for v in flat:
        if v != prev:
            counts.append(run)
            run, prev = 0, v
        run += 1
    counts.append(run)
    return counts


def rle_to_string(counts) -> str:
    """pycocotools maskApi.c rleToString (the inverse of what cocomask.rle_counts_from_string reads)."""
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = not (x == -1 if (ch & 0x10) else x == 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    return "".join(out)


def _blob(rs, h, w):
    m = np.zeros((h, w), np.uint8)
    y0, x0 = rs.randint(0, h - 6), rs.randint(0, w - 6)
    y1, x1 = rs.randint(y0 + 4, h + 1), rs.randint(x0 + 4, w + 1)
    yy, xx = np.mgrid[0:h, 0:w]
    cy, cx, ry, rx = (y0 + y1) / 2, (x0 + x1) / 2, (y1 - y0) / 2, (x1 - x0) / 2
    m[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = 1
    return m, (x0, y0, x1, y1)


def _segm(rs, h, w, kind):
    """One segmentation in one of the four encodings lvis.py:99-121 accepts."""
    m, (x0, y0, x1, y1) = _blob(rs, h, w)
    if kind == "array":
        return m
    if kind == "rle_list":
        return {"size": [h, w], "counts": rle_encode(m)}
    if kind == "rle_str":
        return {"size": [h, w], "counts": rle_to_string(rle_encode(m))}
    # polygons: a triangle and a quadrilateral with fractional vertices inside the blob's box
    t = [x0 + .3, y0 + .2, x1 - .4, y0 + 1.7, (x0 + x1) / 2 + .25, y1 - .6]
    q = [x0 + .5, (y0 + y1) / 2, (x0 + x1) / 2, y0 + .5, x1 - .5, (y0 + y1) / 2 + .3, (x0 + x1) / 2 - .2, y1 - .5]
    return [t, q] if rs.randint(0, 2) else [q]


_KINDS = ("array", "rle_list", "rle_str", "poly")


def build_lvis_tree(root: str, seed: int = 14, n_images: int = 24, n_classes: int = 20) -> str:
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "LVIS")
    sizes = {}
    for i in range(n_images):
        h, w = rs.randint(40, 110), rs.randint(40, 110)
        name = f"val2017/{i:012d}.jpg"
        sizes[name] = (h, w)
        _save(_rand_image(rs, h, w), os.path.join(base, "coco", name))
    names = sorted(sizes)
    val = {}
    for c in range(n_classes):
        cid = 3 * c + 7                                         # sparse category ids, like LVIS
        per = {}
        for j in range(1 + (c % 4)):                            # classes with a single image are dropped at shot >= 1
            name = names[(c * 5 + j * 7) % n_images]
            h, w = sizes[name]
            per[name] = {"annotations": [{"segmentation": _segm(rs, h, w, _KINDS[(c + j + a) % 4])}
                                         for a in range(1 + (c + j) % 3)]}
        val[cid] = per
    os.makedirs(base, exist_ok=True)
    with open(os.path.join(base, "lvis_val.pkl"), "wb") as f:
        pickle.dump(val, f)
    with open(os.path.join(base, "lvis_train.pkl"), "wb") as f:
        pickle.dump({1000 + c: {} for c in range(5)}, f)
    return root


def build_paco_tree(root: str, seed: int = 15, n_images: int = 20) -> str:
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "PACO-Part")
    sizes, img2anno = {}, {}
    for i in range(n_images):
        h, w = rs.randint(48, 120), rs.randint(48, 120)
        sizes[i] = (h, w, f"/data/x/coco/val2017/{i:012d}.jpg")
        _save(_rand_image(rs, h, w), os.path.join(base, "coco", f"val2017/{i:012d}.jpg"))
        img2anno[i] = []
    train_cids = list(range(100, 100 + 448))                    # paco_part.py:87 asserts 448 training categories
    cid2img = {}
    ann = 0
    for v in range(8):                                          # fold-0 validation categories: train_cids[0 + 4 v]
        cid = train_cids[4 * v]
        lst = []
        for j in range(2 + v % 3):
            i = (v * 3 + j * 5) % n_images
            h, w, path = sizes[i]
            lst.append({i: path})
            if j == 0:
                lst.append({i: path})                           # a duplicate, removed by paco_part.py:70-81
            for o in range(1 + (v + j) % 2):                    # objects of this category in the image
                bw, bh = rs.randint(16, w - 4), rs.randint(16, h - 4)
                bx, by = rs.randint(0, w - bw), rs.randint(0, h - bh)
                box = [bx + .4, by + .7, float(bw), float(bh)]
                for p in range(1 + (o + j) % 2):                # parts of the object
                    img2anno[i].append({"category_id": cid, "obj_ann_id": ann, "obj_bbox": box,
                                        "segmentation": _segm(rs, h, w, _KINDS[(v + j + o + p) % 4])})
                ann += 1
        cid2img[cid] = lst
    os.makedirs(os.path.join(base, "paco"), exist_ok=True)
    with open(os.path.join(base, "paco", "paco_part_val.pkl"), "wb") as f:
        pickle.dump({"cid2img": cid2img, "img2anno": img2anno}, f)
    with open(os.path.join(base, "paco", "paco_part_train.pkl"), "wb") as f:
        pickle.dump({"cid2img": {c: [] for c in train_cids}, "img2anno": {}}, f)
    return root


def build_pascal_part_tree(root: str, seed: int = 16, n_images: int = 14) -> str:
    import json
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "Pascal-Part/VOCdevkit/VOC2010")
    objs = {"cat": ["HEAD", "TAIL"], "dog": ["HEAD", "LEG"]}
    index = {"animals": {"object": {o: {"part": {p: {"train": ["t0"], "val": []} for p in ps}} for o, ps in objs.items()}},
             "indoor": {"object": {}}, "person": {"object": {}}, "vehicles": {"object": {}}}
    os.makedirs(os.path.join(base, "Annotations_Part_json_merged_part_classes"), exist_ok=True)
    for i in range(n_images):
        h, w = rs.randint(48, 120), rs.randint(48, 120)
        iid = f"2008_{i:06d}"
        _save(_rand_image(rs, h, w), os.path.join(base, "JPEGImages", iid + ".jpg"))
        anno = {"object": []}
        for o_i, (o, ps) in enumerate(objs.items()):
            if (i + o_i) % 3 == 2:
                continue
            for inst in range(1 + (i + o_i) % 2):
                bw, bh = rs.randint(20, w - 2), rs.randint(20, h - 2)
                bx, by = rs.randint(0, w - bw), rs.randint(0, h - bh)
                parts = []
                for p_i, p in enumerate(ps):
                    if (i + inst + p_i) % 4 == 3:
                        continue                                # this instance lacks the part: resampled (:118-119)
                    masks = []
                    for _ in range(1 + (i + p_i) % 2):
                        m, _b = _blob(rs, h, w)
                        masks.append({"size": [h, w], "counts": rle_to_string(rle_encode(m))})
                    parts.append({"name": p, "mask": masks})
                    lst = index["animals"]["object"][o]["part"][p]["val"]
                    if iid not in lst:
                        lst.append(iid)
                anno["object"].append({"name": o, "bndbox": {"xmin": bx, "ymin": by, "xmax": bx + bw, "ymax": by + bh},
                                       "parts": parts})
        with open(os.path.join(base, "Annotations_Part_json_merged_part_classes", iid + ".json"), "w") as f:
            json.dump(anno, f)
    with open(os.path.join(base, "all_obj_part_to_image.json"), "w") as f:
        json.dump(index, f)
    return root


def build_pascal_cd_tree(root: str, seed: int = 17, n_images: int = 12) -> str:
    """pascal_voc_cd.py: VOC2012 with four validation split files and cd_folds.pth / class_names.pth.  Built in its own
    root (it would overwrite build_pascal_tree's fold0.txt)."""
    rs = np.random.RandomState(seed)
    base = os.path.join(root, "VOC2012")
    folds = {0: [1, 4, 9, 11, 12], 1: [2, 6, 13, 18, 5], 2: [3, 7, 16, 17, 20], 3: [8, 10, 14, 15, 19]}
    lines = {f: [] for f in range(4)}
    for i in range(n_images):
        h, w = rs.randint(36, 120), rs.randint(36, 120)
        name = f"2009_{i:06d}"
        classes = sorted({folds[0][i % 5] - 1, folds[1][i % 5] - 1, folds[0][(i // 2) % 5] - 1})
        _save(_rand_image(rs, h, w), os.path.join(base, "JPEGImages", name + ".jpg"))
        _save(_rand_label(rs, h, w, classes, True), os.path.join(base, "SegmentationClassAug", name + ".png"))
        for c in classes:
            lines[i % 4].append(f"{name}__{c + 1}")
    os.makedirs(os.path.join(base, "splits", "val"), exist_ok=True)
    for f_id in range(4):
        with open(os.path.join(base, "splits", "val", f"fold{f_id}.txt"), "w") as f:
            f.write("\n".join(lines[f_id]) + "\n")
    torch.save(folds, os.path.join(base, "cd_folds.pth"))
    torch.save([f"class{c}" for c in range(20)], os.path.join(base, "class_names.pth"))
    return root
