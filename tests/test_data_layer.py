"""Episode data layer: oracle vs the reference's own outputs, then the CUDA path vs the oracle.

Pinning chain: tests/golden/data_layer.json holds what the UNMODIFIED reference datasets (evaluation_util/data/
{coco,pascal,fss}.py + torchvision transform, dataset.py:36-40) produced on the synthetic trees of tests/data_tree.py
(scripts/make_golden_data.py).  CPU tests: the oracle restatement (oracle/data.py) reproduces those tensors bit for
bit, and agrees with the installed Pillow / torch on random sizes.  GPU tests: the CUDA data layer reproduces the
oracle (and the golden hashes) bit for bit — this is byte / index work, so the bar is exact equality.
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import data_tree  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "data_layer.json")))
S = GOLD["img_size"]


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().to(torch.float32).numpy().tobytes()).hexdigest()


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("fss_data"))
    data_tree.build_coco_tree(root)
    data_tree.build_pascal_tree(root)
    data_tree.build_fss_tree(root)
    data_tree.build_lvis_tree(root)
    data_tree.build_paco_tree(root)
    data_tree.build_pascal_part_tree(root)
    data_tree.build_pascal_cd_tree(os.path.join(root, "cd"))          # its own root: it rewrites VOC2012/splits
    return root


CASES = {"coco_shot1": ("coco", "val", 1), "coco_shot2": ("coco", "val", 2), "pascal_shot1": ("pascal", "val", 1),
         "fss_shot2": ("fss", "test", 2), "lvis_shot1": ("lvis", "val", 1), "lvis_shot2": ("lvis", "val", 2),
         "paco_part_shot1": ("paco_part", "val", 1), "pascal_part_shot1": ("pascal_part", "val", 1),
         "pascal_part_shot2": ("pascal_part", "val", 2), "pascal_cd_shot1": ("pascal_cd", "val", 1)}


def _dataset(tree, key):
    from diffews_b200 import data
    bench, split, shot = CASES[key]
    if bench == "pascal_cd":
        tree = os.path.join(tree, "cd")
    data.FSSDataset.initialize(S, tree, False)
    cls = data.FSSDataset.datasets[bench]
    return cls(tree, fold=0, transform=data.FSSDataset.transform, split=split, shot=shot, use_original_imgsize=False)


def _rel(name, root):
    return os.path.relpath(name, root) if os.path.isabs(str(name)) else name


def _oracle_episode(ds, raw):
    """The reference's per-episode tensors from a raw (decoded) episode, through the CPU oracle."""
    from oracle import data as O
    out = {"query_img": O.transform_image(raw["query_img"], S),
           "support_imgs": torch.stack([O.transform_image(a, S) for a in raw["support_imgs"]])}
    c = raw["class_sample"]
    labs = [raw["query_label"]] + raw["support_labels"]
    if ds.benchmark == "pascal":
        mb = [O.pascal_mask(l, c, S) for l in labs]
        masks, bnd = [m for m, _ in mb], [b for _, b in mb]
        out["query_ignore_idx"], out["support_ignore_idxs"] = bnd[0], torch.stack(bnd[1:])
        if not getattr(ds, "emit_query_ignore", True):
            del out["query_ignore_idx"]                      # pascal_voc_cd.py:61
    elif ds.benchmark == "fss":
        masks = [torch.from_numpy(O.nearest_resize((l >= 128).astype(np.float32), S, S).copy()) for l in labs]
    elif ds.benchmark in ("lvis", "paco_part", "pascal_part"):
        # lvis.py:41-48: the union bitmask .float(), nearest-resized
        masks = [torch.from_numpy(O.nearest_resize((l > 0).astype(np.float32), S, S).copy()) for l in labs]
    else:
        masks = [O.coco_mask(l, c, S) for l in labs]
    out["query_mask"], out["support_masks"] = masks[0], torch.stack(masks[1:])
    return out


# ---- CPU: oracle pinned against Pillow / torch / the reference's golden outputs --------------------------------------
def test_oracle_resize_matches_pillow():
    from PIL import Image
    from oracle import data as O
    rng = np.random.default_rng(0)
    for it in range(40):
        h, w = int(rng.integers(5, 160)), int(rng.integers(5, 160))
        oh, ow = int(rng.integers(4, 120)), int(rng.integers(4, 120))
        if it % 7 == 0:
            ow = w
        if it % 11 == 0:
            oh = h
        a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        if it % 3 == 0:
            a = (a // 128 * 255).astype(np.uint8)
        ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(ref, O.pil_resize_bilinear_u8(a, oh, ow)), (h, w, oh, ow)


def test_oracle_nearest_matches_torch():
    from oracle import data as O
    for i in list(range(1, 200)) + [333, 480, 500, 640, 767, 1333]:
        for o in (1, 3, 48, 64, 96, 384, 417, 512, 768, 2 * i, i):
            m = torch.arange(i, dtype=torch.float32)[None, None, :, None]
            ref = torch.nn.functional.interpolate(m, (o, 1), mode="nearest").flatten().long().numpy()
            assert np.array_equal(ref, O.nearest_src_index(o, i)), (i, o)


@pytest.mark.parametrize("key", list(CASES))
def test_oracle_and_host_sampling_match_reference_golden(tree, key):
    """Host half of the product (file layout parsing, np.random call order, decoding) + the oracle's pixel arithmetic
    == the unmodified reference datasets, tensor for tensor."""
    ds = _dataset(tree, key)
    np.random.seed(0)
    for idx, g in enumerate(GOLD[key]):
        raw = ds.raw_episode(idx)
        assert _rel(raw["query_name"], tree) == g["query_name"]
        assert [_rel(s, tree) for s in raw["support_names"]] == g["support_names"]
        assert raw["class_sample"] == g["class_id"]
        if "org_query_imsize" in g:
            assert list(raw["org_query_imsize"]) == g["org_query_imsize"]
        o = _oracle_episode(ds, raw)
        assert list(o["query_img"].shape) == g["query_img_shape"]
        assert list(o["support_imgs"].shape) == g["support_imgs_shape"]
        assert o["query_img"].flatten()[:4].tolist() == g["query_img_first"]
        for k in ("query_img", "query_mask", "support_imgs", "support_masks", "query_ignore_idx", "support_ignore_idxs"):
            if k in g:
                assert sha(o[k]) == g[k], (key, idx, k)


def test_class_ids_and_lengths(tree):
    from oracle import data as O
    ds = _dataset(tree, "coco_shot1")
    assert ds.class_ids == O.coco_class_ids(0, "val") and len(ds) == 1000 and ds.benchmark == "coco"
    ds = _dataset(tree, "pascal_shot1")
    assert ds.class_ids == O.pascal_class_ids(0, "val") == [0, 1, 2, 3, 4] and len(ds) == 1000
    ds = _dataset(tree, "fss_shot2")
    assert list(ds.class_ids) == list(range(760, 1000)) and len(ds) == 30


def test_segmentation_benchmark_metadata_matches_reference(tree):
    """Class splits of the instance-segmentation benchmarks (lvis.py:66-90, paco_part.py:62-100, pascal_part.py:31-46)
    against what the unmodified reference built on the same trees."""
    for key in ("lvis_shot1", "lvis_shot2", "paco_part_shot1", "pascal_part_shot1"):
        ds, meta = _dataset(tree, key), GOLD[key + "_meta"]
        assert ds.nclass == meta["nclass"] and list(ds.class_ids) == meta["class_ids"]
        if "class_ids_ori" in meta:
            assert list(ds.class_ids_ori) == meta["class_ids_ori"]
        if "cat_part_name" in meta:
            assert ds.cat_part_name == meta["cat_part_name"] and len(ds) == meta["len"]
    assert len(_dataset(tree, "lvis_shot1")) == 2300 and len(_dataset(tree, "paco_part_shot1")) == 2500
    ds, meta = _dataset(tree, "pascal_cd_shot1"), GOLD["pascal_cd_shot1_meta"]
    assert (ds.nclass, list(ds.class_ids), len(ds), len(ds.img_metadata)) == (meta["nclass"], meta["class_ids"],
                                                                              meta["len"], meta["n_metadata"])


def test_cocomask_rle_and_polygons():
    """diffews_b200/cocomask.py (restated pycocotools maskApi.c; the library is not installed): RLE string <-> counts <->
    mask round trips, and polygons with hand-checkable answers."""
    from diffews_b200 import cocomask
    rs = np.random.RandomState(0)
    for _ in range(40):
        h, w = rs.randint(1, 40), rs.randint(1, 40)
        m = (rs.rand(h, w) > rs.rand()).astype(np.uint8)
        counts = data_tree.rle_encode(m)
        assert sum(counts) == h * w
        assert cocomask.rle_counts_from_string(data_tree.rle_to_string(counts)) == counts
        assert np.array_equal(cocomask.rle_decode(counts, h, w), m)
        assert np.array_equal(cocomask.decode_rle_dict({"size": [h, w], "counts": data_tree.rle_to_string(counts)}), m)
    # an integer-cornered rectangle [x0,x1) x [y0,y1) covers exactly (x1-x0)(y1-y0) pixels (COCO's bbox-polygon area)
    for x0, y0, x1, y1 in ((10, 10, 20, 20), (0, 0, 30, 25), (3, 7, 4, 8), (5, 0, 29, 3)):
        m = cocomask.polygons_to_bitmask([[x0, y0, x1, y0, x1, y1, x0, y1]], 25, 30)
        ref = np.zeros((25, 30), bool)
        ref[y0:y1, x0:x1] = True
        assert np.array_equal(m, ref), (x0, y0, x1, y1)
    # vertex order / starting vertex do not matter; the union of two polygons is the OR of their masks
    a = [4.3, 3.2, 25.6, 5.7, 14.25, 20.4]
    b = [2.5, 12.0, 15.0, 2.5, 27.5, 12.3, 14.8, 22.5]
    ma, mb = cocomask.polygons_to_bitmask([a], 25, 30), cocomask.polygons_to_bitmask([b], 25, 30)
    assert np.array_equal(cocomask.polygons_to_bitmask([a[2:] + a[:2]], 25, 30), ma)
    assert np.array_equal(cocomask.polygons_to_bitmask([a, b], 25, 30), ma | mb)
    # a convex polygon's mask lies within one pixel of the exact point-in-polygon set of pixel centres
    yy, xx = np.mgrid[0:25, 0:30]
    inside = np.ones((25, 30), bool)
    pts = np.asarray(b).reshape(-1, 2)
    for i in range(len(pts)):
        (xa, ya), (xb, yb) = pts[i], pts[(i + 1) % len(pts)]
        inside &= ((xb - xa) * (yy + .5 - ya) - (yb - ya) * (xx + .5 - xa)) >= 0
    assert abs(int(mb.sum()) - int(inside.sum())) <= 0.15 * inside.sum()
    grown = np.zeros_like(inside)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            grown |= np.roll(np.roll(inside, dy, 0), dx, 1)
    assert not (mb & ~grown).any()


def test_pack_layout():
    from diffews_b200 import ops
    from diffews_b200.data import EpisodeCollator
    a = np.zeros((5, 7, 3), np.uint8)
    b = np.zeros((4, 9), np.uint8)
    total, offs, desc = EpisodeCollator.pack([a, b], [0, 6])
    assert desc.dtype.itemsize == ops.IMAGE_DESC_BYTES == 24
    assert offs[0] == 48 and offs[0] % 16 == 0 and offs[1] == ((48 + 105 + 15) & ~15)
    assert tuple(desc[0]) == (48, 5, 7, 21, 0) and tuple(desc[1]) == (offs[1], 4, 9, 9, 6)
    assert total >= offs[1] + 36


def test_loader_visits_reference_order_and_keeps_short_batch(tree):
    """DataLoader(shuffle=False, num_workers=0) semantics: indices in order, last batch short, RNG consumed per item."""
    from diffews_b200.data import EpisodeLoader
    ds = _dataset(tree, "fss_shot2")
    np.random.seed(0)
    loader = EpisodeLoader(ds, bsz=4, device="cpu", decode_threads=2)
    assert len(loader) == 8
    names = []
    sizes = []
    for raws in loader.raw_batches():
        sizes.append(len(raws))
        names += [_rel(r["query_name"], tree) for r in raws]
    assert sizes == [4] * 7 + [2]
    assert names[:5] == [g["query_name"] for g in GOLD["fss_shot2"]]


def test_loader_shards_batches_over_ranks_without_changing_the_episode_sequence(tree):
    """world = 2: every rank consumes the numpy RNG for every episode, rank r keeps batches r, r+2, ...; the union is
    the single-process sequence (so all-reduced counts equal a one-GPU run of the same seed)."""
    from diffews_b200.data import EpisodeLoader
    ds = _dataset(tree, "coco_shot2")
    def names(rank, world):
        np.random.seed(0)
        loader = EpisodeLoader(ds, bsz=2, device="cpu", decode_threads=1, rank=rank, world=world)
        out = []
        for i, raws in enumerate(loader.raw_batches()):
            out.append([(r["query_name"], tuple(r["support_names"]), r["class_sample"]) for r in raws])
            if i == 2:
                break
        return out
    single = names(0, 1)
    r0, r1 = names(0, 2), names(1, 2)
    assert r0[0] == single[0] and r1[0] == single[1] and r0[1] == single[2]
    assert [e[0] for b in single[:1] for e in b] == [g["query_name"] for g in GOLD["coco_shot2"][:2]]
    lens = [len(EpisodeLoader(ds, bsz=16, device="cpu", decode_threads=1, rank=r, world=8)) for r in range(8)]
    assert lens == [8, 8, 8, 8, 8, 8, 8, 7] and sum(lens) == 63                 # 1000 episodes / 16 -> 63 batches


def test_collate_without_gpu_fails_loudly(tree):
    from diffews_b200.data import EpisodeLoader
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ds = _dataset(tree, "coco_shot1")
    np.random.seed(0)
    with pytest.raises(Exception):
        next(iter(EpisodeLoader(ds, bsz=1, device="cpu", decode_threads=1)))


# ---- GPU: the CUDA data layer == oracle == reference golden, bit for bit ---------------------------------------------
def _upload(arrays, params):
    from diffews_b200.data import EpisodeCollator
    total, offs, desc = EpisodeCollator.pack(arrays, params)
    host = np.zeros(total, np.uint8)
    host[:desc.nbytes] = desc.view(np.uint8)
    for a, o in zip(arrays, offs):
        host[o:o + a.size] = a.reshape(-1)
    return torch.from_numpy(host).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("key", list(CASES))
@pytest.mark.parametrize("bsz", [1, 3])
def test_gpu_loader_matches_reference_golden(lib_built, tree, key, bsz):
    from diffews_b200.data import EpisodeLoader
    ds = _dataset(tree, key)
    np.random.seed(0)
    loader = EpisodeLoader(ds, bsz=bsz, device="cuda", decode_threads=4)
    gold = GOLD[key]
    seen = 0
    for batch in loader:
        B = batch["query_img"].shape[0]
        assert batch["query_img"].dtype == torch.float32 and batch["query_img"].is_cuda
        assert batch["class_id"].dtype == torch.int64 and batch["class_id"].is_cuda
        for b in range(B):
            if seen >= len(gold):
                break
            g = gold[seen]
            assert _rel(batch["query_name"][b], tree) == g["query_name"]
            assert [_rel(batch["support_names"][j][b], tree) for j in range(len(g["support_names"]))] == g["support_names"]
            assert int(batch["class_id"][b]) == g["class_id"]
            for k in ("query_img", "query_mask", "support_imgs", "support_masks", "query_ignore_idx",
                      "support_ignore_idxs"):
                if k in g:
                    assert sha(batch[k][b]) == g[k], (key, seen, k)
            seen += 1
        if seen >= len(gold):
            break
    assert seen == len(gold)


@pytest.mark.gpu
def test_gpu_resize_matches_pillow_bytes(lib_built):
    """Mixed sizes in one batch: down / up-scaling, identity axes, exact 2x, a 13x reduction (wide coefficient
    windows), and the BASELINE sizes (480x640 -> 512^2, 800x1333 -> 768^2)."""
    from PIL import Image
    from diffews_b200 import ops
    from oracle import data as O
    rng = np.random.default_rng(5)
    for out_h, out_w, shapes in [(64, 64, [(64, 64), (32, 32), (30, 200), (200, 30), (64, 100), (100, 64), (850, 70), (5, 5)]),
                                 (512, 512, [(480, 640), (640, 427), (512, 512), (333, 500)]),
                                 (768, 768, [(800, 1333), (768, 1024)]),
                                 (40, 56, [(123, 77), (40, 56), (20, 28)])]:
        imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
        imgs[0][: imgs[0].shape[0] // 2] = 255          # saturated rows: exercises the clip
        buf = _upload(imgs, [0] * len(imgs))
        f32, u8 = ops.resize_normalize_u8(buf, 0, len(imgs), max(h for h, _ in shapes), max(w for _, w in shapes), out_h,
                                          out_w, want_u8=True)
        torch.cuda.synchronize()
        for i, a in enumerate(imgs):
            ref = np.asarray(Image.fromarray(a).resize((out_w, out_h), Image.BILINEAR))
            got = u8[i].cpu().numpy()
            assert np.array_equal(ref, got), (shapes[i], out_h, out_w, int(np.abs(ref.astype(int) - got).max()))
            assert torch.equal(f32[i].cpu(), O.to_tensor_normalize(ref)), shapes[i]


@pytest.mark.gpu
def test_gpu_mask_nearest_matches_torch(lib_built):
    from diffews_b200 import ops
    rng = np.random.default_rng(6)
    shapes = [(480, 640), (100, 37), (64, 64), (32, 32), (500, 333), (7, 3)]
    labs = [rng.integers(0, 6, (h, w), dtype=np.uint8) for h, w in shapes]
    for l in labs:
        l[rng.integers(0, 2, l.shape).astype(bool) & (rng.integers(0, 8, l.shape) == 0)] = 255
    params = [1, 2, 3, 4, 5, 1]
    buf = _upload(labs, params)
    for out_h, out_w in [(64, 64), (512, 512), (48, 80)]:
        m, bnd = ops.mask_nearest(buf, 0, len(labs), out_h, out_w, 0, want_boundary=True)
        m2, _ = ops.mask_nearest(buf, 0, len(labs), out_h, out_w, 1)
        for i, l in enumerate(labs):
            t = torch.from_numpy(l.astype(np.float32))[None, None]
            r = torch.nn.functional.interpolate(t, (out_h, out_w), mode="nearest")[0, 0]
            assert torch.equal(m[i].cpu(), (r == params[i]).float()), (shapes[i], out_h, out_w)
            assert torch.equal(bnd[i].cpu(), (r / 255).floor()), (shapes[i], out_h, out_w)
            assert torch.equal(m2[i].cpu(), (r >= 128).float()), (shapes[i], out_h, out_w)


@pytest.mark.gpu
def test_gpu_original_imgsize_query_mask(lib_built, tree):
    """use_original_imgsize=True keeps the query mask at the source size (coco.py:41), bsz = 1 only."""
    from diffews_b200 import data
    data.FSSDataset.initialize(S, tree, True)
    np.random.seed(0)
    loader = data.FSSDataset.build_dataloader("coco", 1, 0, 0, "val", 1)
    batch = next(iter(loader))
    w, h = int(batch["org_query_imsize"][0][0]), int(batch["org_query_imsize"][1][0])
    assert tuple(batch["query_mask"].shape) == (1, h, w) and tuple(batch["query_img"].shape) == (1, 3, S, S)
    assert float(batch["query_mask"].sum()) > 0
    data.FSSDataset.initialize(S, tree, False)


@pytest.fixture(scope="module")
def small_pipe():
    from diffews_b200.runner import build_engine_from_modules
    from diffews_b200.synthetic import prompt_embedding
    from oracle.sd21 import build_models
    unet_o, vae_o = build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    return build_engine_from_modules(unet_o, vae_o, prompt_embedding())


@pytest.mark.gpu
def test_gpu_loader_feeds_the_runner(lib_built, tree, small_pipe):
    """The loader's batch dict drives EpisodeRunner.step unchanged (keys / layouts of main_oss.py:94-110)."""
    from diffews_b200 import data
    from diffews_b200.runner import EpisodeRunner
    data.FSSDataset.initialize(64, tree, False)
    np.random.seed(0)
    loader = data.FSSDataset.build_dataloader("pascal", 2, 2, 0, "val", 1)
    runner = EpisodeRunner(small_pipe, benchmark="pascal", class_ids=loader.dataset.class_ids, img_size=64)
    it = iter(loader)
    for _ in range(2):
        batch = next(it)
        inter, union = runner.step(batch)
        assert inter.shape == (2, 2) and (union >= inter).all()
        # boundary (ignore) pixels are dropped from every count (evaluation.py:16-21 + histc range)
        valid = (batch["query_ignore_idx"] == 0).sum(dim=(1, 2))
        assert (inter.sum(dim=1) <= valid).all() and (union <= valid[:, None]).all()
    miou, fb, _ = runner.finish()
    assert 0.0 <= float(miou) <= 100.0


@pytest.mark.gpu
def test_gpu_run_loop_graph_equals_eager_and_honours_ignore(lib_built, tree, small_pipe):
    """EpisodeRunner.run (= test_diffusion, main_oss.py:84-171) over the PASCAL loader: the CUDA-graph replay and the
    eager launches accumulate identical int64 buffers, and the boundary pixels reach the metric kernel in both."""
    from diffews_b200 import data
    from diffews_b200.runner import EpisodeRunner
    bufs = []
    for graph in (False, True):
        data.FSSDataset.initialize(64, tree, False)
        np.random.seed(0)
        loader = data.FSSDataset.build_dataloader("pascal", 2, 2, 0, "val", 1)
        runner = EpisodeRunner(small_pipe, benchmark="pascal", class_ids=loader.dataset.class_ids, img_size=64)
        miou, fb = runner.run(loader, max_batches=3, use_cuda_graph=graph)
        assert 0.0 <= miou <= 100.0 and 0.0 <= fb <= 100.0
        if graph:
            assert "query_ignore_idx" in runner._graph_shapes
        bufs.append((runner.meter.intersection_buf.cpu().clone(), runner.meter.union_buf.cpu().clone()))
    assert torch.equal(bufs[0][0], bufs[1][0]) and torch.equal(bufs[0][1], bufs[1][1])
    # 3 batches x 2 episodes x 64^2 pixels minus the ignored boundary: strictly fewer pixels than the full images
    assert int(bufs[0][1].sum()) > 0


@pytest.mark.gpu
def test_gpu_evaluator_matches_unmodified_reference_evaluator(lib_built):
    """a12 on the GPU against the reference's own numbers: diffews_b200.evaluation.Evaluator.classify_prediction (one
    launch of dfw_rthres_iou_hist, int64 counts) == what the UNMODIFIED evaluation_util/common/evaluation.py returned
    for the same seeded masks (tests/golden/metric_reference.json), including the PASCAL ignore boundary, an all-
    background pair (empty foreground histogram) and a fully disjoint pair."""
    from diffews_b200.evaluation import Evaluator
    gold = json.load(open(os.path.join(HERE, "golden", "metric_reference.json")))["cases"]
    for c, g in zip(data_tree.metric_cases(), gold):
        batch = {"query_mask": c["gt"].cuda()}
        if c["ign"] is not None:
            batch["query_ignore_idx"] = c["ign"].cuda()
        inter, union = Evaluator.classify_prediction(c["pred"].cuda(), batch)
        assert inter.dtype == torch.float32 and tuple(inter.shape) == (2, c["B"])
        assert inter.cpu().tolist() == g["area_inter"], g["kind"]
        assert union.cpu().tolist() == g["area_union"], g["kind"]


def test_training_shuffle_draws_like_torch_random_sampler(tree):
    """split == 'trn' -> DataLoader(shuffle=True) (dataset.py:46-48).  torch's RandomSampler draws a seed from the global
    torch RNG and permutes with a private generator; EpisodeLoader._indices must consume the global RNG identically so a
    seeded training run visits the same indices as the reference's loader."""
    from torch.utils.data import DataLoader, Dataset
    from diffews_b200.data import EpisodeLoader

    class Idx(Dataset):
        def __len__(self):
            return 30

        def __getitem__(self, i):
            return i

    ds = _dataset(tree, "fss_shot2")                      # 30 items
    for seed in (0, 7):
        torch.manual_seed(seed)
        ref = [int(b) for b in DataLoader(Idx(), batch_size=1, shuffle=True, num_workers=0)]
        torch.manual_seed(seed)
        got = EpisodeLoader(ds, bsz=4, shuffle=True, device="cpu", decode_threads=1)._indices()
        assert got == ref and sorted(got) == list(range(30))
