"""Golden vectors produced by the UNMODIFIED reference: the episode data layer (datasets) and the intersection / union
metric (Evaluator.classify_prediction) — the two parts of the path whose reference modules import without diffusers.

Runs in the build container only (it imports /root/reference/evaluation_util/data/{coco,pascal,fss}.py by file path —
they need torch, PIL, numpy only — and torchvision's Resize/ToTensor/Normalize exactly as dataset.py:36-40 builds them)
on the synthetic trees of tests/data_tree.py, and writes tests/golden/data_layer.json: per episode the sampled names,
class id, and SHA-256 of every tensor's bytes (fp32, C order), plus a few raw values.  The tests rebuild the same trees
and compare the oracle (CPU) and the CUDA data layer (GPU) against these.

    python scripts/make_golden_data.py
"""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

import numpy as np
import torch
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import data_tree  # noqa: E402

REF = "/root/reference/evaluation_util/data"
IMG_SIZE = 48


def _install_segm_standins():
    """lvis.py / paco_part.py / pascal_part.py import cv2 (unused), detectron2.structures.masks (polygons_to_bitmask,
    and through its star-import `mask_util`) and pycocotools.mask (decode).  None is installed.  The stand-ins forward
    to diffews_b200.cocomask, the restatement of pycocotools' maskApi.c: for ARRAY segmentations (half of the synthetic
    annotations) no stand-in arithmetic is involved at all; for polygon / RLE segmentations the rasteriser under the
    reference's code is the restatement itself, so those fixtures pin everything AROUND the rasteriser (sampling, union,
    box crop, resize), not the rasteriser."""
    import types
    from diffews_b200 import cocomask

    def decode(rle):
        if isinstance(rle, (list, tuple)):
            if len(rle) == 0:
                return np.zeros((0, 0, 0), np.uint8)
            return np.stack([cocomask.decode_rle_dict(r) for r in rle], axis=-1)
        return cocomask.decode_rle_dict(rle)

    mask_util = types.ModuleType("pycocotools.mask")
    mask_util.decode = decode
    pyco = types.ModuleType("pycocotools")
    pyco.mask = mask_util
    d2m = types.ModuleType("detectron2.structures.masks")
    d2m.polygons_to_bitmask = lambda polygons, height, width: cocomask.polygons_to_bitmask(polygons, height, width)
    d2m.mask_util = mask_util
    d2s = types.ModuleType("detectron2.structures")
    d2s.masks = d2m
    d2 = types.ModuleType("detectron2")
    d2.structures = d2s
    sys.modules.update({"cv2": types.ModuleType("cv2"), "pycocotools": pyco, "pycocotools.mask": mask_util,
                        "detectron2": d2, "detectron2.structures": d2s, "detectron2.structures.masks": d2m})


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().to(torch.float32).numpy().tobytes()).hexdigest()


def episodes(ds, n, root):
    out = []
    for idx in range(n):
        b = ds[idx]
        e = {"query_name": os.path.relpath(b["query_name"], root) if os.path.isabs(str(b["query_name"])) else b["query_name"],
             "support_names": [os.path.relpath(s, root) if os.path.isabs(str(s)) else s for s in b["support_names"]],
             "class_id": int(b["class_id"]),
             "query_img": sha(b["query_img"]), "query_mask": sha(b["query_mask"]),
             "support_imgs": sha(b["support_imgs"]), "support_masks": sha(b["support_masks"]),
             "query_img_shape": list(b["query_img"].shape), "support_imgs_shape": list(b["support_imgs"].shape),
             "query_mask_sum": float(b["query_mask"].sum()), "query_img_first": b["query_img"].flatten()[:4].tolist()}
        if "query_ignore_idx" in b:
            e["query_ignore_idx"] = sha(b["query_ignore_idx"])
        if "support_ignore_idxs" in b:
            e["support_ignore_idxs"] = sha(b["support_ignore_idxs"])
        if "org_query_imsize" in b:
            e["org_query_imsize"] = list(b["org_query_imsize"])
        out.append(e)
    return out


def metric_golden():
    spec = importlib.util.spec_from_file_location("ref_evaluation", "/root/reference/evaluation_util/common/evaluation.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.Evaluator.initialize()
    out = []
    for c in data_tree.metric_cases():
        batch = {"query_mask": c["gt"].clone()}
        if c["ign"] is not None:
            batch["query_ignore_idx"] = c["ign"].clone()
        inter, union = m.Evaluator.classify_prediction(c["pred"].clone(), batch)      # float32 [2,B]
        out.append({k: c[k] for k in ("seed", "B", "H", "W", "ignore", "kind")} |
                   {"area_inter": inter.tolist(), "area_union": union.tolist(), "dtype": str(inter.dtype)})
    return out


def meter_golden(metric):
    """The UNMODIFIED evaluation_util/common/logger.py AverageMeter (update / compute_iou, logger.py:10-51) fed with the
    reference Evaluator's own outputs.  The class hard-codes `.cuda()` (logger.py:15, :30-31); there is no GPU in the
    build container, so `torch.Tensor.cuda` is patched to the identity for the duration of the run — the module's
    source is not touched."""
    spec = importlib.util.spec_from_file_location("ref_logger", "/root/reference/evaluation_util/common/logger.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)

    class DS:
        benchmark = "coco"
        class_ids = [4 * v for v in range(20)]                # COCO-20i fold 0 (coco.py:60-66)

    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        meter = m.AverageMeter(DS())
        g = torch.Generator().manual_seed(5)
        class_ids = []
        for rep in range(3):                                   # the same episodes land in different classes
            for c in metric:
                cid = torch.tensor(DS.class_ids)[torch.randint(0, 20, (c["B"],), generator=g)]
                class_ids.append(cid.tolist())
                meter.update(torch.tensor(c["area_inter"]), torch.tensor(c["area_union"]), cid, None)
        miou, fb_iou, head = meter.compute_iou()
    finally:
        torch.Tensor.cuda = orig
    return {"benchmark": "coco", "class_ids_interest": DS.class_ids, "update_class_ids": class_ids,
            "miou": float(miou), "fb_iou": float(fb_iou), "iou_head": head.tolist(),
            "intersection_buf": meter.intersection_buf.tolist(), "union_buf": meter.union_buf.tolist(),
            "buf_dtype": str(meter.intersection_buf.dtype)}


def main():
    tf = transforms.Compose([transforms.Resize(size=(IMG_SIZE, IMG_SIZE)), transforms.ToTensor(),
                             transforms.Normalize([0.5], [0.5])])                      # dataset.py:36-40
    gold = {"img_size": IMG_SIZE, "made_by": "scripts/make_golden_data.py (reference datasets, unmodified)",
            "versions": {"torch": torch.__version__, "numpy": np.__version__}}
    with tempfile.TemporaryDirectory() as root:
        data_tree.build_coco_tree(root)
        data_tree.build_pascal_tree(root)
        data_tree.build_fss_tree(root)
        coco, pascal, fss = _load("coco"), _load("pascal"), _load("fss")
        for shot in (1, 2):
            np.random.seed(0)
            ds = coco.DatasetCOCO(root, fold=0, transform=tf, split="val", shot=shot, use_original_imgsize=False)
            gold[f"coco_shot{shot}"] = episodes(ds, 6, root)
        np.random.seed(0)
        ds = pascal.DatasetPASCAL(root, fold=0, transform=tf, split="val", shot=1, use_original_imgsize=False)
        gold["pascal_shot1"] = episodes(ds, 6, root)
        np.random.seed(0)
        ds = fss.DatasetFSS(root, fold=0, transform=tf, split="test", shot=2, use_original_imgsize=False)
        gold["fss_shot2"] = episodes(ds, 5, root)
        # instance-segmentation benchmarks (import stand-ins: see _install_segm_standins)
        sys.path.insert(0, ROOT)
        _install_segm_standins()
        data_tree.build_lvis_tree(root)
        data_tree.build_paco_tree(root)
        data_tree.build_pascal_part_tree(root)
        lvis, paco, ppart = _load("lvis"), _load("paco_part"), _load("pascal_part")
        for shot in (1, 2):
            np.random.seed(0)
            ds = lvis.DatasetLVIS(root, fold=0, transform=tf, split="val", shot=shot, use_original_imgsize=False)
            gold[f"lvis_shot{shot}"] = episodes(ds, 8, root)
            gold[f"lvis_shot{shot}_meta"] = {"nclass": ds.nclass, "class_ids": ds.class_ids, "class_ids_ori": ds.class_ids_ori}
        np.random.seed(0)
        ds = paco.DatasetPACOPart(root, fold=0, transform=tf, split="val", shot=1, use_original_imgsize=False)
        gold["paco_part_shot1"] = episodes(ds, 8, root)
        gold["paco_part_shot1_meta"] = {"nclass": ds.nclass, "class_ids": ds.class_ids, "class_ids_ori": ds.class_ids_ori}
        for shot in (1, 2):
            np.random.seed(0)
            ds = ppart.DatasetPASCALPart(root, fold=0, transform=tf, split="val", shot=shot, use_original_imgsize=False)
            gold[f"pascal_part_shot{shot}"] = episodes(ds, 8, root)
            gold[f"pascal_part_shot{shot}_meta"] = {"nclass": ds.nclass, "class_ids": ds.class_ids, "len": len(ds),
                                                     "cat_part_name": ds.cat_part_name}
    with tempfile.TemporaryDirectory() as root:
        data_tree.build_pascal_cd_tree(root)
        np.random.seed(0)
        ds = _load("pascal_voc_cd").DatasetPASCALCD(root, fold=0, transform=tf, split="val", shot=1)
        gold["pascal_cd_shot1"] = episodes(ds, 6, root)
        gold["pascal_cd_shot1_meta"] = {"nclass": ds.nclass, "class_ids": ds.class_ids, "len": len(ds),
                                        "n_metadata": len(ds.img_metadata)}
    mpath = os.path.join(ROOT, "tests", "golden", "metric_reference.json")
    with open(mpath, "w") as f:
        cases = metric_golden()
        json.dump({"made_by": "scripts/make_golden_data.py: unmodified evaluation_util/common/evaluation.py "
                              "Evaluator.classify_prediction and evaluation_util/common/logger.py AverageMeter on CPU "
                              "(Tensor.cuda patched to the identity: no GPU in the build container)",
                   "cases": cases, "meter": meter_golden(cases)}, f, indent=1)
    print("wrote", mpath)
    path = os.path.join(ROOT, "tests", "golden", "data_layer.json")
    with open(path, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
