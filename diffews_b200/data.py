"""Episode data layer — the B200 counterpart of evaluation_util/data/ (the caller-side data format).

Reference: `FSSDataset.initialize / build_dataloader` (evaluation_util/data/dataset.py:14-52) hands a torchvision
`Resize((S,S)) -> ToTensor -> Normalize([0.5],[0.5])` transform to a per-benchmark `Dataset` whose `__getitem__`
samples an episode with `np.random.choice`, decodes it with PIL, resizes every image on the CPU and nearest-resizes the
class mask (coco.py:32-58, pascal.py:37-76, fss.py:35-62); a `DataLoader(bsz, shuffle=split=='trn')` collates.

Here the split is: episode SAMPLING and file DECODING stay on the host (same numpy call sequence as the reference, so
a seeded run visits the same episodes; decoding runs on a thread pool), while every per-pixel operation — the Pillow
bilinear resample, ToTensor, Normalize, the class-mask selection and its nearest resize — runs on the GPU in four
launches per batch (`dfw_resize_normalize_u8`, `dfw_mask_nearest`; csrc/preproc.cu) on the raw decoded bytes, which
are shipped in ONE pinned-host -> device copy (a 640x480 JPEG is 0.9 MB decoded against 3.1 MB as the fp32 512x512
tensor the reference moves).  The resulting batch dict has the reference's keys, shapes and dtypes and is bit-identical
to the reference's CPU tensors (tests/test_data_layer.py, golden vectors made by the unmodified reference datasets).

No CPU fallback: collation needs the CUDA library; `raw_episode()` (host only) works anywhere.
"""
from __future__ import annotations

import glob
import os
import pickle
from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional

import numpy as np
import torch

from . import ops

MASK_EQ, MASK_GE128 = 0, 1      # dfw_mask_nearest modes


class EpisodeTransform:
    """Stands in for the torchvision Compose of dataset.py:36-40: it only carries the parameters; the arithmetic runs
    in `EpisodeCollator` on the GPU."""

    def __init__(self, img_size: int, mean: float = 0.5, std: float = 0.5):
        self.size = (img_size, img_size)
        self.mean, self.std = mean, std


def _read_rgb(path: str) -> np.ndarray:
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def _read_label(path: str, to_l: bool = False) -> np.ndarray:
    from PIL import Image
    im = Image.open(path)
    if to_l:
        im = im.convert("L")
    a = np.asarray(im)
    if a.ndim != 2 or a.dtype != np.uint8:
        raise ValueError(f"{path}: label masks must be single-channel 8-bit images, got {a.shape} {a.dtype}")
    return a


class _EpisodeDataset:
    """Common part: `__getitem__` returns a RAW episode (decoded uint8 arrays, no resizing)."""
    benchmark = ""
    mask_mode = MASK_EQ
    has_ignore = False

    def __len__(self):
        raise NotImplementedError

    def sample_names(self, idx: int):
        """-> (query_name, support_names, class_sample): consumes the global numpy RNG exactly like the reference."""
        raise NotImplementedError

    def paths(self, name: str):
        """-> (image path, label path)."""
        raise NotImplementedError

    def rgb_path(self, query_name: str) -> str:
        return self.paths(query_name)[0]

    def decode(self, sampled, pool: Optional[ThreadPoolExecutor] = None) -> dict:
        query_name, support_names, class_sample = sampled[:3]
        names = [query_name] + list(support_names)
        jobs = []
        for n in names:
            ip, lp = self.paths(n)
            jobs.append((_read_rgb, ip))
            jobs.append((self._read_label, lp))
        if pool is None:
            res = [f(p) for f, p in jobs]
        else:
            res = list(pool.map(lambda fp: fp[0](fp[1]), jobs))
        imgs, labels = res[0::2], res[1::2]
        h, w = imgs[0].shape[:2]
        return {"query_img": imgs[0], "query_label": labels[0], "support_imgs": imgs[1:], "support_labels": labels[1:],
                "query_name": query_name, "support_names": list(support_names), "class_sample": int(class_sample),
                "mask_param": int(class_sample) + 1, "org_query_imsize": (w, h), "rgb_path": self.rgb_path(query_name)}

    def _read_label(self, path):
        return _read_label(path)

    def raw_episode(self, idx: int) -> dict:
        return self.decode(self.sample_names(idx))

    __getitem__ = raw_episode


class DatasetCOCO(_EpisodeDataset):
    """COCO-20i (evaluation_util/data/coco.py).  Files: COCO2014/splits/{trn,val}/fold{f}.pkl (class -> image names),
    COCO2014/<name> images, COCO2014/annotations/<name with .png> class-index masks."""
    benchmark = "coco"

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize):
        self.split = "val" if split in ["val", "test"] else "trn"              # coco.py:14
        self.fold, self.nfolds, self.nclass = fold, 4, 80
        self.shot = shot
        self.base_path = os.path.join(datapath, "COCO2014")
        self.transform = transform
        self.use_original_imgsize = use_original_imgsize
        nclass_trn = self.nclass // self.nfolds                                 # coco.py:60-66
        val = [self.fold + self.nfolds * v for v in range(nclass_trn)]
        self.class_ids = val if self.split != "trn" else [x for x in range(self.nclass) if x not in val]
        with open(f"{self.base_path}/splits/{self.split}/fold{self.fold}.pkl", "rb") as f:   # coco.py:68-71
            self.img_metadata_classwise = pickle.load(f)
        meta = []
        for k in self.img_metadata_classwise.keys():                            # coco.py:73-77
            meta += self.img_metadata_classwise[k]
        self.img_metadata = sorted(list(set(meta)))

    def __len__(self):
        return len(self.img_metadata) if self.split == "trn" else 1000          # coco.py:28-29

    def sample_names(self, idx):
        # coco.py:84-101 — idx is ignored: class-uniform sampling from the global numpy RNG
        class_sample = np.random.choice(self.class_ids, 1, replace=False)[0]
        query_name = np.random.choice(self.img_metadata_classwise[class_sample], 1, replace=False)[0]
        support_names = []
        while True:
            support_name = np.random.choice(self.img_metadata_classwise[class_sample], 1, replace=False)[0]
            if query_name != support_name:
                support_names.append(support_name)
            if len(support_names) == self.shot:
                break
        return query_name, support_names, class_sample

    def paths(self, name):
        mask_path = os.path.join(self.base_path, "annotations", name)
        return os.path.join(self.base_path, name), mask_path[:mask_path.index(".jpg")] + ".png"     # coco.py:79-82


class DatasetPASCAL(_EpisodeDataset):
    """PASCAL-5i (evaluation_util/data/pascal.py): VOC2012/JPEGImages, SegmentationClassAug, splits/{split}/fold{f}.txt
    lines `<name>__<class 1..20>`; label 255 = boundary -> query_ignore_idx."""
    benchmark = "pascal"
    has_ignore = True

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize):
        self.split = "val" if split in ["val", "test"] else "trn"
        self.fold, self.nfolds, self.nclass = fold, 4, 20
        self.base_path = os.path.join(datapath, "VOC2012")
        self.shot = shot
        self.use_original_imgsize = use_original_imgsize
        self.img_path = os.path.join(datapath, "VOC2012/JPEGImages/")
        self.ann_path = os.path.join(datapath, "VOC2012/SegmentationClassAug/")
        self.transform = transform
        n = self.nclass // self.nfolds                                          # pascal.py:112-120
        val = [self.fold * n + i for i in range(n)]
        self.class_ids = val if self.split != "trn" else [x for x in range(self.nclass) if x not in val]
        self.img_metadata = self._build_img_metadata()
        self.img_metadata_classwise = {c: [] for c in range(self.nclass)}       # pascal.py:148-155
        for img_name, img_class in self.img_metadata:
            self.img_metadata_classwise[img_class] += [img_name]

    def _build_img_metadata(self):                                              # pascal.py:122-146
        def read(split, fold_id):
            with open(os.path.join(self.base_path, "splits/%s/fold%d.txt" % (split, fold_id)), "r") as f:
                rows = f.read().split("\n")[:-1]
            return [[r.split("__")[0], int(r.split("__")[1]) - 1] for r in rows]
        if self.split == "trn":
            out = []
            for fold_id in range(self.nfolds):
                if fold_id != self.fold:
                    out += read(self.split, fold_id)
            return out
        return read(self.split, self.fold)

    def __len__(self):
        return len(self.img_metadata) if self.split == "trn" else 1000          # pascal.py:34-35

    def sample_names(self, idx):
        idx %= len(self.img_metadata)                                           # pascal.py:38, :101-110
        query_name, class_sample = self.img_metadata[idx]
        support_names = []
        while True:
            support_name = np.random.choice(self.img_metadata_classwise[class_sample], 1, replace=False)[0]
            if query_name != support_name:
                support_names.append(support_name)
            if len(support_names) == self.shot:
                break
        return query_name, support_names, class_sample

    def paths(self, name):
        return os.path.join(self.img_path, name) + ".jpg", os.path.join(self.ann_path, name) + ".png"


class DatasetPASCALCD(DatasetPASCAL):
    """PASCAL cross-domain folds (evaluation_util/data/pascal_voc_cd.py): the PASCAL-5i files, with the fold's classes
    read from VOC2012/cd_folds.pth (1-based ids), validation episodes drawn from ALL four split files filtered to those
    classes, and no `query_ignore_idx` in the batch (pascal_voc_cd.py:61) — the support ignore masks stay."""
    emit_query_ignore = False

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize=False):
        self._fold_classes = torch.load(os.path.join(datapath, "VOC2012", "cd_folds.pth"))
        self.class_names = torch.load(os.path.join(datapath, "VOC2012", "class_names.pth"))
        super().__init__(datapath, fold, transform, split, shot, False)          # :43 always resizes the query mask
        val = [x - 1 for x in self._fold_classes[self.fold]]                       # pascal_voc_cd.py:111-120
        self.class_ids = val if self.split != "trn" else [x for x in range(self.nclass) if x not in val]

    def _build_img_metadata(self):                                              # pascal_voc_cd.py:122-147
        keep = self._fold_classes[self.fold]

        def read(split, fold_id):
            with open(os.path.join(self.base_path, f"splits/{split}/fold{fold_id}.txt"), "r") as f:
                rows = f.read().split("\n")[:-1]
            return [[r.split("__")[0], int(r.split("__")[1]) - 1] for r in rows if int(r.split("__")[1]) in keep]
        out = []
        for fold_id in range(self.nfolds):
            if self.split == "trn" and fold_id == self.fold:
                continue
            out += read(self.split, fold_id)
        return out


class DatasetFSS(_EpisodeDataset):
    """FSS-1000 (evaluation_util/data/fss.py): FSS-1000/data/<category>/{1..10}.jpg + .png, splits/{split}.txt."""
    benchmark = "fss"
    mask_mode = MASK_GE128

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize):
        self.split = split
        self.shot = shot
        self.nclass = 1000
        self.use_original_imgsize = False                                        # fss.py:38 always resizes
        self.base_path = os.path.join(datapath, "FSS-1000/data")
        with open(os.path.join(datapath, "FSS-1000/splits/%s.txt" % split), "r") as f:
            self.categories = sorted(f.read().split("\n")[:-1])                  # fss.py:22-24
        self.class_ids = {"trn": range(0, 520), "val": range(520, 760), "test": range(760, 1000)}[split]
        self.img_metadata = []
        for cat in self.categories:                                              # fss.py:108-115
            for p in sorted(glob.glob("%s/*" % os.path.join(self.base_path, cat))):
                if os.path.basename(p).split(".")[1] == "jpg":
                    self.img_metadata.append(p)
        self.transform = transform

    def __len__(self):
        return len(self.img_metadata)

    def sample_names(self, idx):
        query_name = self.img_metadata[idx]                                      # fss.py:86-104
        class_sample = self.categories.index(query_name.split("/")[-2])
        class_sample += {"trn": 0, "val": 520, "test": 760}[self.split]
        support_names = []
        while True:
            support_name = np.random.choice(range(1, 11), 1, replace=False)[0]
            support_name = os.path.join(os.path.dirname(query_name), str(support_name)) + ".jpg"
            if query_name != support_name:
                support_names.append(support_name)
            if len(support_names) == self.shot:
                break
        return query_name, support_names, class_sample

    def paths(self, name):
        return name, os.path.join(os.path.dirname(name), name.split("/")[-1].split(".")[0]) + ".png"

    def rgb_path(self, query_name):
        return query_name

    def _read_label(self, path):
        return _read_label(path, to_l=True)                                      # fss.py:80-84 (.convert('L'))


# --------------------------------------------------------------------------------------------------------------------
class _SegmDataset(_EpisodeDataset):
    """Benchmarks whose masks come from instance segmentations (polygons / RLE / arrays) instead of a class-index PNG:
    the host rasterises and unions them (`cocomask`, in the decode pool) and crops to the object box where the
    reference does; the device then sees a {0,1} label with `mask_param` = 1, i.e. the same `dfw_mask_nearest` launch.
    `sample_names` returns (query_name, support_names, class_id, extra) — `extra` carries what the sampling loop of the
    reference already looked up (segmentations, boxes), so that decoding needs no RNG."""

    def image_path(self, name: str) -> str:
        raise NotImplementedError

    def paths(self, name):
        return self.image_path(name), None

    @staticmethod
    def _union(segms, h, w) -> np.ndarray:
        # lvis.py:131-135: cat of float masks, .sum(0) > 0
        from . import cocomask
        acc = np.zeros((h, w), np.float32)
        for s in segms:
            acc += cocomask.segmentation_to_mask(s, h, w).astype(np.float32)
        return (acc > 0).astype(np.uint8)

    @staticmethod
    def _crop(img, mask, box_xyxy):
        x0, y0, x1, y1 = box_xyxy
        return np.ascontiguousarray(img[y0:y1, x0:x1]), np.ascontiguousarray(mask[y0:y1, x0:x1])

    def _one(self, job):
        name, segms, box = job
        img = _read_rgb(self.image_path(name))
        h, w = img.shape[:2]
        mask = segms if isinstance(segms, np.ndarray) else self._union(segms, h, w)
        if box is not None:
            img, mask = self._crop(img, mask, box)
        return img, mask

    def decode(self, sampled, pool: Optional[ThreadPoolExecutor] = None) -> dict:
        query_name, support_names, class_id, extra = sampled
        jobs = [(query_name, extra["query_segms"], extra.get("query_box"))]
        jobs += [(n, s, b) for n, s, b in zip(support_names, extra["support_segms"],
                                              extra.get("support_boxes") or [None] * len(support_names))]
        res = [self._one(j) for j in jobs] if pool is None else list(pool.map(self._one, jobs))
        h, w = res[0][0].shape[:2]
        raw = {"query_img": res[0][0], "query_label": res[0][1], "support_imgs": [r[0] for r in res[1:]],
               "support_labels": [r[1] for r in res[1:]], "query_name": query_name, "support_names": list(support_names),
               "class_sample": int(class_id), "mask_param": 1, "org_query_imsize": (w, h),
               "rgb_path": self.image_path(query_name)}
        if "category" in extra:
            raw["category"] = extra["category"]
        return raw


class DatasetLVIS(_SegmDataset):
    """LVIS-92i (evaluation_util/data/lvis.py).  Files: LVIS/lvis_{train,val}.pkl = {category id: {image name:
    {'annotations': [{'segmentation': polygons | RLE | array}, ...]}}}, images under LVIS/coco/<name>."""
    benchmark = "lvis"

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize):
        self.split = "val" if split in ["val", "test"] else "trn"              # lvis.py:17
        self.fold, self.nfolds = fold, 10
        self.shot = shot
        self.anno_path = os.path.join(datapath, "LVIS")
        self.base_path = os.path.join(datapath, "LVIS", "coco")
        self.transform = transform
        self.use_original_imgsize = use_original_imgsize
        with open(os.path.join(self.anno_path, "lvis_train.pkl"), "rb") as f:   # lvis.py:66-90
            train_anno = pickle.load(f)
        with open(os.path.join(self.anno_path, "lvis_val.pkl"), "rb") as f:
            val_anno = pickle.load(f)
        train_cat_ids = [i for i in list(train_anno.keys()) if len(train_anno[i]) > self.shot]
        val_cat_ids = [i for i in list(val_anno.keys()) if len(val_anno[i]) > self.shot]
        class_ids_val = [val_cat_ids[self.fold + self.nfolds * v] for v in range(len(val_cat_ids) // self.nfolds)]
        if self.split == "trn":
            self.class_ids_ori = [x for x in train_cat_ids if x not in class_ids_val]
            self.nclass, self.img_metadata_classwise = len(train_cat_ids), train_anno
        else:
            self.class_ids_ori = class_ids_val
            self.nclass, self.img_metadata_classwise = len(val_cat_ids), val_anno
        self.class_ids_c = {cid: i for i, cid in enumerate(self.class_ids_ori)}
        self.class_ids = sorted(list(self.class_ids_c.values()))
        meta = []
        for k in self.img_metadata_classwise.keys():                            # lvis.py:92-96
            meta.extend(list(self.img_metadata_classwise[k].keys()))
        self.img_metadata = sorted(list(set(meta)))

    def __len__(self):
        return len(self.img_metadata) if self.split == "trn" else 2300          # lvis.py:33-34

    def image_path(self, name):
        return os.path.join(self.base_path, name)

    def sample_names(self, idx):
        idx %= len(self.class_ids)                                              # lvis.py:37, :123-157
        class_sample = self.class_ids_ori[idx]
        per_class = self.img_metadata_classwise[class_sample]
        query_name = np.random.choice(list(per_class.keys()), 1, replace=False)[0]
        query_segms = [a["segmentation"] for a in per_class[query_name]["annotations"]]
        support_names, support_segms = [], []
        while True:
            support_name = np.random.choice(list(per_class.keys()), 1, replace=False)[0]
            if query_name != support_name:
                support_names.append(support_name)
                support_segms.append([a["segmentation"] for a in per_class[support_name]["annotations"]])
            if len(support_names) == self.shot:
                break
        return query_name, support_names, self.class_ids_c[class_sample], {"query_segms": query_segms,
                                                                            "support_segms": support_segms}


class DatasetPACOPart(_SegmDataset):
    """PACO-Part (evaluation_util/data/paco_part.py).  Files: PACO-Part/paco/paco_part_{train,val}.pkl =
    {'cid2img': {category: [{image id: path}, ...]}, 'img2anno': {image id: [{'category_id', 'obj_ann_id', 'obj_bbox'
    (xywh), 'segmentation'}, ...]}}, images under PACO-Part/coco/<last two path components>."""
    benchmark = "paco_part"

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize, box_crop=True):
        self.split = "val" if split in ["val", "test"] else "trn"
        self.fold, self.nfolds, self.nclass = fold, 4, 448
        self.shot = shot
        self.img_path = os.path.join(datapath, "PACO-Part", "coco")
        self.anno_path = os.path.join(datapath, "PACO-Part", "paco")
        self.transform = transform
        self.use_original_imgsize = use_original_imgsize
        self.box_crop = box_crop
        with open(os.path.join(self.anno_path, "paco_part_train.pkl"), "rb") as f:    # paco_part.py:62-100
            train_anno = pickle.load(f)
        with open(os.path.join(self.anno_path, "paco_part_val.pkl"), "rb") as f:
            test_anno = pickle.load(f)
        dedup = {}
        for cid in test_anno["cid2img"]:                                        # first occurrence of every image id
            seen = []
            dedup.setdefault(cid, [])
            for img in test_anno["cid2img"][cid]:
                img_id = list(img.keys())[0]
                if img_id not in seen:
                    seen.append(img_id)
                    dedup[cid].append(img)
        test_anno["cid2img"] = dedup
        train_cat_ids = list(train_anno["cid2img"].keys())
        test_cat_ids = [i for i in list(test_anno["cid2img"].keys()) if len(test_anno["cid2img"][i]) > self.shot]
        assert len(train_cat_ids) == self.nclass
        class_ids_val = [train_cat_ids[self.fold + self.nfolds * v] for v in range(self.nclass // self.nfolds)]
        class_ids_val = [x for x in class_ids_val if x in test_cat_ids]
        anno = train_anno if self.split == "trn" else test_anno
        self.class_ids_ori = [x for x in train_cat_ids if x not in class_ids_val] if self.split == "trn" else class_ids_val
        self.cid2img, self.img2anno = anno["cid2img"], anno["img2anno"]
        self.class_ids_c = {cid: i for i, cid in enumerate(self.class_ids_ori)}
        self.class_ids = sorted(list(self.class_ids_c.values()))
        self.img_metadata = []
        for k in self.cid2img.keys():                                           # paco_part.py:102-106
            self.img_metadata += self.cid2img[k]

    def __len__(self):
        return len(self.img_metadata) if self.split == "trn" else 2500          # paco_part.py:32-33

    def image_path(self, name):
        return os.path.join(self.img_path, name)

    def _objects(self, img_id, class_sample):
        objs = {}
        for anno in self.img2anno[img_id]:                                      # paco_part.py:136-147
            if anno["category_id"] == class_sample:
                o = objs.setdefault(anno["obj_ann_id"], {"obj_bbox": [], "segms": []})
                o["obj_bbox"].append(anno["obj_bbox"])
                o["segms"].append(anno["segmentation"])
        return objs

    @staticmethod
    def _xyxy(b):
        return [int(b[0]), int(b[1]), int(b[0] + b[2]), int(b[1] + b[3])]       # paco_part.py:199 slicing bounds

    def sample_names(self, idx):
        # paco_part.py:126-185 — idx ignored; class, query image, query object, then per shot (image, object)
        class_sample = np.random.choice(self.class_ids_ori, 1, replace=False)[0]
        query = np.random.choice(self.cid2img[class_sample], 1, replace=False)[0]
        query_id, query_name = list(query.keys())[0], list(query.values())[0]
        query_name = "/".join(query_name.split("/")[-2:])
        qobjs = self._objects(query_id, class_sample)
        sel = np.random.choice(list(qobjs.keys()), 1, replace=False)[0]
        query_box, query_segms = qobjs[sel]["obj_bbox"][0], qobjs[sel]["segms"]
        support_names, support_segms, support_boxes = [], [], []
        while True:
            support = np.random.choice(self.cid2img[class_sample], 1, replace=False)[0]
            support_id, support_name = list(support.keys())[0], list(support.values())[0]
            support_name = "/".join(support_name.split("/")[-2:])
            if query_name != support_name:
                support_names.append(support_name)
                sobjs = self._objects(support_id, class_sample)
                ssel = np.random.choice(list(sobjs.keys()), 1, replace=False)[0]
                support_boxes.append(sobjs[ssel]["obj_bbox"][0])
                support_segms.append(sobjs[ssel]["segms"])
            if len(support_names) == self.shot:
                break
        extra = {"query_segms": query_segms, "support_segms": support_segms}
        if self.box_crop:
            extra["query_box"] = self._xyxy(query_box)
            extra["support_boxes"] = [self._xyxy(b) for b in support_boxes]
        return query_name, support_names, self.class_ids_c[class_sample], extra


class DatasetPASCALPart(_SegmDataset):
    """PASCAL-Part (evaluation_util/data/pascal_part.py).  Files under Pascal-Part/VOCdevkit/VOC2010/:
    all_obj_part_to_image.json ({super-category: {'object': {obj: {'part': {part: {'train': [...], 'val': [...]}}}}}}),
    JPEGImages/<id>.jpg, Annotations_Part_json_merged_part_classes/<id>.json ({'object': [{'name', 'bndbox',
    'parts': [{'name', 'mask': [RLE, ...]}]}]}).  fold selects the super-category."""
    benchmark = "pascal_part"

    def __init__(self, datapath, fold, transform, split, shot, use_original_imgsize, box_crop=True):
        import json
        self.split = "val" if split in ["val", "test"] else "train"
        self.cat = ["animals", "indoor", "person", "vehicles"][fold]
        self.shot = shot
        self.transform = transform
        self.use_original_imgsize = use_original_imgsize
        self.box_crop = box_crop
        base = os.path.join(datapath, "Pascal-Part/VOCdevkit/VOC2010")
        self.img_file = os.path.join(base, "JPEGImages/{}.jpg")
        self.anno_file = os.path.join(base, "Annotations_Part_json_merged_part_classes/{}.json")
        with open(os.path.join(base, "all_obj_part_to_image.json"), "r") as f:
            self.cat_annos = json.load(f)[self.cat]
        self.cat_part_name = []
        for obj in self.cat_annos["object"]:                                    # pascal_part.py:34-46
            for part in self.cat_annos["object"][obj]["part"]:
                p = self.cat_annos["object"][obj]["part"][part]
                if len(p["train"]) > 0 and len(p["val"]) > 0:
                    if obj + "+" + part == "aeroplane+TAIL":
                        continue
                    self.cat_part_name.append(obj + "+" + part)
        self.class_ids = self.cat_part_id = list(range(len(self.cat_part_name)))
        self.nclass = len(self.cat_part_id)
        self.img_metadata = []
        for obj in self.cat_annos["object"]:                                    # pascal_part.py:56-63
            for part in self.cat_annos["object"][obj]["part"]:
                self.img_metadata.extend(self.cat_annos["object"][obj]["part"][part][self.split])

    def __len__(self):
        # pascal_part.py:50-54: the 'trn' branch is unreachable there too (split is 'train' or 'val')
        return min(len(self.img_metadata), 2500)

    def image_path(self, name):
        return self.img_file.format(name)

    def _sample_part(self, img_id, obj_n, part_n):
        """pascal_part.py:103-131 for one candidate image: (object, union of the part's masks) or None."""
        import json
        from . import cocomask
        with open(self.anno_file.format(img_id), "r") as f:
            anno = json.load(f)
        objs = [o for o in anno["object"] if o["name"] == obj_n]
        assert len(objs) > 0
        sel_obj = np.random.choice(objs, 1, replace=False)[0]
        rles = []
        for p in sel_obj["parts"]:
            if p["name"] == part_n:
                rles.extend(p["mask"])
        if not rles:
            return None
        part_mask = sum(cocomask.decode_rle_dict(m).astype(np.int64) for m in rles) > 0
        if part_mask.size == 0:
            return None
        return sel_obj, part_mask.astype(np.uint8)

    def sample_names(self, idx):
        idx %= len(self.class_ids)                                              # pascal_part.py:67, :98-178
        class_sample, class_id = self.cat_part_name[idx], self.class_ids[idx]
        obj_n, part_n = class_sample.split("+")
        pool = self.cat_annos["object"][obj_n]["part"][part_n][self.split]
        while True:
            query_id = np.random.choice(pool, 1, replace=False)[0]
            got = self._sample_part(query_id, obj_n, part_n)
            if got is not None:
                break
        q_obj, q_mask = got
        box = lambda o: [int(o["bndbox"][b]) for b in o["bndbox"]]              # xyxy, dict order (pascal_part.py:136)
        support_ids, support_masks, support_boxes = [], [], []
        while True:
            while True:
                sid = np.random.choice(pool, 1, replace=False)[0]
                if sid == query_id or sid in support_ids:
                    continue
                got = self._sample_part(sid, obj_n, part_n)
                if got is not None:
                    break
            support_ids.append(sid)
            support_masks.append(got[1])
            support_boxes.append(box(got[0]))
            if len(support_ids) == self.shot:
                break
        extra = {"query_segms": q_mask, "support_segms": support_masks, "category": class_sample}
        if self.box_crop:
            extra["query_box"], extra["support_boxes"] = box(q_obj), support_boxes
        return str(query_id), [str(s) for s in support_ids], class_id, extra


# --------------------------------------------------------------------------------------------------------------------
class EpisodeCollator:
    """Raw episodes -> the reference's batch dict on the device (keys / shapes / dtypes of coco.py:45-57, plus the
    PASCAL ignore tensors), all pixel work on the GPU."""

    def __init__(self, dataset: _EpisodeDataset, device="cuda"):
        self.ds = dataset
        self.device = torch.device(device)
        t = dataset.transform
        if isinstance(t, int):
            t = EpisodeTransform(t)
        self.size = tuple(t.size)
        self.mean, self.std = float(getattr(t, "mean", 0.5)), float(getattr(t, "std", 0.5))
        self._pinned = None

    @staticmethod
    def pack(arrays: List[np.ndarray], params: List[int]):
        """Lay the descriptor table and the pixel arrays out in one byte buffer: [n x DfwImageDesc | images ...], every
        image 16-byte aligned.  Returns (total bytes, offsets list, descriptor bytes)."""
        n = len(arrays)
        desc = np.zeros(n, dtype=np.dtype([("offset", "<i8"), ("h", "<i4"), ("w", "<i4"), ("row_stride", "<i4"),
                                           ("param", "<i4")]))
        assert desc.dtype.itemsize == ops.IMAGE_DESC_BYTES
        off = (n * ops.IMAGE_DESC_BYTES + 15) & ~15
        offsets = []
        for i, a in enumerate(arrays):
            assert a.dtype == np.uint8 and a.ndim in (2, 3)
            h, w = a.shape[:2]
            row = w * (3 if a.ndim == 3 else 1)
            desc[i] = (off, h, w, row, params[i])
            offsets.append(off)
            off = (off + h * row + 15) & ~15
        return off, offsets, desc

    def _upload(self, arrays, params):
        total, offsets, desc = self.pack(arrays, params)
        if self._pinned is None or self._pinned.numel() < total:
            self._pinned = torch.empty(max(total, 1 << 20), dtype=torch.uint8).pin_memory()
        host = self._pinned.numpy()
        host[:desc.nbytes] = desc.view(np.uint8)
        for a, o in zip(arrays, offsets):
            host[o:o + a.size] = np.ascontiguousarray(a).reshape(-1)
        dev = torch.empty(total, dtype=torch.uint8, device=self.device)
        dev.copy_(self._pinned[:total], non_blocking=True)
        # the pinned staging buffer is reused by the next batch: make the host wait for this copy (not for the kernels)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._copy_done = ev
        return dev

    def __call__(self, raws: List[dict]) -> dict:
        if self.device.type != "cuda":
            raise RuntimeError("EpisodeCollator runs on the CUDA library only (no CPU fallback); use raw_batches() "
                               "for host-side inspection")
        if getattr(self, "_copy_done", None) is not None:
            self._copy_done.synchronize()
        B, k = len(raws), len(raws[0]["support_imgs"])
        S_h, S_w = self.size
        imgs = [r["query_img"] for r in raws] + [s for r in raws for s in r["support_imgs"]]
        labels = [r["query_label"] for r in raws] + [s for r in raws for s in r["support_labels"]]
        cls1 = [r["mask_param"] for r in raws] + [r["mask_param"] for r in raws for _ in range(k)]
        n = len(imgs)
        for im, lb in zip(imgs, labels):
            if im.shape[:2] != lb.shape:
                raise ValueError(f"image {im.shape[:2]} and label mask {lb.shape} sizes differ")
        # one buffer, one copy: [n image descriptors | n label descriptors | pixels ...]
        buf = self._upload(imgs + labels, [0] * n + cls1)
        label_desc = n * ops.IMAGE_DESC_BYTES
        out, _ = ops.resize_normalize_u8(buf, 0, n, max(a.shape[0] for a in imgs), max(a.shape[1] for a in imgs), S_h, S_w,
                                         self.mean, self.std)
        ign = self.ds.has_ignore
        masks, bnd = ops.mask_nearest(buf, label_desc, n, S_h, S_w, self.ds.mask_mode, want_boundary=ign)
        batch = {
            "rgb_path": [r["rgb_path"] for r in raws],
            "query_img": out[:B],
            "query_mask": masks[:B],
            "query_name": [r["query_name"] for r in raws],
            "org_query_imsize": [torch.tensor([r["org_query_imsize"][0] for r in raws]),
                                 torch.tensor([r["org_query_imsize"][1] for r in raws])],
            "support_imgs": out[B:].view(B, k, 3, S_h, S_w),
            "support_masks": masks[B:].view(B, k, S_h, S_w),
            "support_names": [[r["support_names"][j] for r in raws] for j in range(k)],   # default_collate transposes
            "class_id": torch.tensor([r["class_sample"] for r in raws], dtype=torch.int64).to(self.device),
        }
        if ign:
            if getattr(self.ds, "emit_query_ignore", True):
                batch["query_ignore_idx"] = bnd[:B]
            batch["support_ignore_idxs"] = bnd[B:].view(B, k, S_h, S_w)
        if "category" in raws[0]:
            batch["category"] = [r["category"] for r in raws]                           # pascal_part.py:91
        if getattr(self.ds, "use_original_imgsize", False):
            # coco.py:41 / pascal.py:42: the query mask keeps its own size (the reference can then only run bsz = 1)
            if B != 1:
                raise ValueError("use_original_imgsize needs bsz = 1 (query masks of different sizes cannot be stacked)")
            h, w = labels[0].shape
            qm, qb = ops.mask_nearest(buf, label_desc, 1, h, w, self.ds.mask_mode, want_boundary=ign)
            batch["query_mask"] = qm
            if ign:
                batch["query_ignore_idx"] = qb
        return batch


class EpisodeLoader:
    """`DataLoader(dataset, batch_size=bsz, shuffle=split=='trn', num_workers=0)` (dataset.py:44-50) with the collation
    on the GPU.  Sampling order == the reference's single-process loader: indices 0..len-1 in order at test time, a
    torch.randperm drawn like RandomSampler at training time; the last batch may be short.

    Data-parallel evaluation: with `world` > 1 every rank draws EVERY episode from the numpy RNG (names
    only — cheap), so all ranks walk the reference's single-process episode sequence, and rank r decodes / collates
    batches r, r + world, ...  The all-reduced int64 counts are then bit-identical to a one-GPU run of the same seed."""

    def __init__(self, dataset: _EpisodeDataset, bsz: int, shuffle: bool = False, device="cuda", decode_threads: int = 8,
                 rank: Optional[int] = None, world: Optional[int] = None):
        self.dataset = dataset
        self.batch_size = bsz
        self.shuffle = shuffle
        self.collate = EpisodeCollator(dataset, device)
        self.pool = ThreadPoolExecutor(decode_threads) if decode_threads > 1 else None
        if rank is None or world is None:
            import torch.distributed as dist
            on = dist.is_available() and dist.is_initialized()
            rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
        if not 0 <= rank < world:
            raise ValueError(f"rank {rank} outside world {world}")
        self.rank, self.world = rank, world

    def _num_batches(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __len__(self):
        n = self._num_batches()
        return (n - self.rank + self.world - 1) // self.world

    def _indices(self):
        n = len(self.dataset)
        # torch's DataLoader iterator first draws its `_base_seed` from the global torch RNG (both modes), then — on the
        # first batch — RandomSampler draws the permutation seed: same consumption here, so a `torch.manual_seed(s)` run
        # visits the reference's indices
        torch.empty((), dtype=torch.int64).random_()
        if not self.shuffle:
            return list(range(n))
        seed = int(torch.empty((), dtype=torch.int64).random_().item())         # torch RandomSampler.__iter__
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g).tolist()

    def raw_batches(self):
        idx = self._indices()
        for bi, s in enumerate(range(0, len(idx), self.batch_size)):
            sampled = [self.dataset.sample_names(i) for i in idx[s:s + self.batch_size]]     # RNG order = reference
            if bi % self.world != self.rank:
                continue
            yield [self.dataset.decode(sm, self.pool) for sm in sampled]

    def __iter__(self):
        for raws in self.raw_batches():
            yield self.collate(raws)


class FSSDataset:
    """evaluation_util/data/dataset.py:14-52, same classmethod protocol."""
    datasets = {"coco": DatasetCOCO, "pascal": DatasetPASCAL, "fss": DatasetFSS, "paco_part": DatasetPACOPart,
                "pascal_part": DatasetPASCALPart, "lvis": DatasetLVIS, "pascal_cd": DatasetPASCALCD}

    @classmethod
    def initialize(cls, img_size, datapath, use_original_imgsize):
        cls.datapath = datapath
        cls.use_original_imgsize = use_original_imgsize
        cls.transform = EpisodeTransform(img_size)

    @classmethod
    def build_dataloader(cls, benchmark, bsz, nworker, fold, split, shot=1, device="cuda", rank=None, world=None):
        if benchmark not in cls.datasets:
            raise NotImplementedError(f"benchmark {benchmark!r}: only {sorted(cls.datasets)} exist (dataset.py:18-26)")
        shuffle = split == "trn"
        dataset = cls.datasets[benchmark](cls.datapath, fold=fold, transform=cls.transform, split=split, shot=shot,
                                          use_original_imgsize=cls.use_original_imgsize)
        return EpisodeLoader(dataset, bsz, shuffle=shuffle, device=device, decode_threads=max(1, nworker), rank=rank,
                             world=world)
