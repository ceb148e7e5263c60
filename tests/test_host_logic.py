"""CPU tests of the host-side logic: weight re-layouts, synthetic episodes, scheduler, sharding, gloo all-reduce."""
import json
import os

import pytest
import torch

from diffews_b200.weights import conv_weight_to_gemm, geglu_permute


def test_conv_weight_layout():
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    g = conv_weight_to_gemm(w)
    assert g.shape == (2, 27)
    # K index = (kh*3 + kw)*Cin + c
    for co in range(2):
        for kh in range(3):
            for kw in range(3):
                for c in range(3):
                    assert g[co, (kh * 3 + kw) * 3 + c] == w[co, c, kh, kw]


def test_geglu_permutation_roundtrip():
    C = 64
    w = torch.randn(8 * C, C)
    b = torch.randn(8 * C)
    wp, bp = geglu_permute(w, b)
    x = torch.randn(5, C)
    h = x @ w.t() + b
    ref = h[:, :4 * C] * torch.nn.functional.gelu(h[:, 4 * C:])
    hp = x @ wp.t() + bp
    out = torch.empty(5, 4 * C)
    for t in range(8 * C // 256):           # what the DFW_EPI_GEGLU epilogue does per 256-column tile
        v, gte = hp[:, t * 256:t * 256 + 128], hp[:, t * 256 + 128:(t + 1) * 256]
        out[:, t * 128:(t + 1) * 128] = v * torch.nn.functional.gelu(gte)
    assert torch.allclose(out, ref, atol=1e-5)


def test_synthetic_episode_layout_and_determinism():
    from diffews_b200.synthetic import make_batch, pipeline_inputs
    b1, b2 = make_batch(3, 2, 64, 2), make_batch(3, 2, 64, 2)
    for k in b1:
        assert torch.equal(b1[k], b2[k])
    assert b1["query_img"].shape == (2, 3, 64, 64) and b1["support_imgs"].shape == (2, 2, 3, 64, 64)
    assert b1["support_masks"].shape == (2, 2, 64, 64) and b1["query_mask"].shape == (2, 64, 64)
    assert set(b1["query_mask"].unique().tolist()) <= {0.0, 1.0}
    frac = b1["query_mask"].mean().item()
    assert 0.01 < frac < 0.9
    si, q, sm = pipeline_inputs(b1)               # main_oss.py:99-110
    assert si.shape == (4, 3, 64, 64) and sm.shape == (4, 3, 64, 64) and q.shape == (2, 3, 64, 64)
    assert set(sm.unique().tolist()) <= {-1.0, 1.0}
    assert -1 <= si.min() and si.max() <= 1


def test_shard_episodes_partition():
    from diffews_b200.runner import shard_episodes
    for n, world in [(1000, 8), (17, 4), (5, 8)]:
        seen = []
        for r in range(world):
            seen += list(shard_episodes(n, r, world))
        assert sorted(seen) == list(range(n))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffews_b200.evaluation import AverageMeter
    from diffews_b200.runner import shard_episodes
    m = AverageMeter(benchmark="coco", class_ids=range(80), device="cpu")
    g = torch.Generator().manual_seed(0)
    inter = torch.randint(0, 1000, (40, 2), generator=g)
    union = inter + torch.randint(1, 1000, (40, 2), generator=g)
    cls = torch.randint(0, 80, (40,), generator=g)
    for i in shard_episodes(40, rank, world):      # host stand-in for the accumulation kernel (no GPU here)
        m.intersection_buf[:, cls[i]] += inter[i]
        m.union_buf[:, cls[i]] += union[i]
    ptrs = (m.intersection_buf.data_ptr(), m.union_buf.data_ptr())
    m.all_reduce()
    # a captured CUDA graph holds these pointers (EpisodeRunner.enable_cuda_graph): the reduction must land IN PLACE
    assert ptrs == (m.intersection_buf.data_ptr(), m.union_buf.data_ptr())
    first = (m.intersection_buf.clone(), m.union_buf.clone())
    m.reset()                                       # second fold on the same meter: zero in place, accumulate, reduce again
    assert ptrs == (m.intersection_buf.data_ptr(), m.union_buf.data_ptr()) and int(m.union_buf.sum()) == 0
    for i in shard_episodes(40, rank, world):
        m.intersection_buf[:, cls[i]] += inter[i]
        m.union_buf[:, cls[i]] += union[i]
    m.all_reduce()
    assert torch.equal(first[0], m.intersection_buf) and torch.equal(first[1], m.union_buf)
    miou, fb, _ = m.compute_iou()
    # plain lists, not tensors: a tensor in a Queue travels as a shared-memory handle that dies with this process
    q.put((rank, m.intersection_buf.tolist(), m.union_buf.tolist(), float(miou), float(fb)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_data_parallel_counts_equal_single_process():
    """world_size-2 gloo: all-reduced int64 [2,nclass] buffers == single-process accumulation over the same episodes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:                     # a port that is free right now
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    g = torch.Generator().manual_seed(0)
    inter = torch.randint(0, 1000, (40, 2), generator=g)
    union = inter + torch.randint(1, 1000, (40, 2), generator=g)
    cls = torch.randint(0, 80, (40,), generator=g)
    ib = torch.zeros(2, 80, dtype=torch.int64); ub = torch.zeros(2, 80, dtype=torch.int64)
    for i in range(40):
        ib[:, cls[i]] += inter[i]; ub[:, cls[i]] += union[i]
    for rank, i_buf, u_buf, miou, fb in res:
        assert i_buf == ib.tolist() and u_buf == ub.tolist()
    assert res[0][3] == res[1][3] and res[0][4] == res[1][4]


def test_checkpoint_layout_roundtrip(tmp_path):
    """diffusers directory layout (main_oss.py:338-369): safetensors + config.json per sub-model are read back verbatim;
    a plain SD-2.1 UNet gets the 8-channel support stem the reference's conversion script builds."""
    from safetensors.torch import save_file
    from diffews_b200 import checkpoint as ck
    g = torch.Generator().manual_seed(0)
    sd = {"conv_in.weight": torch.randn(8, 4, 3, 3, generator=g), "conv_in.bias": torch.randn(8, generator=g),
          "mid_block.resnets.0.norm1.weight": torch.randn(8, generator=g)}
    (tmp_path / "unet").mkdir()
    save_file(sd, str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    (tmp_path / "unet" / "config.json").write_text(json.dumps({"block_out_channels": [8, 16], "attention_head_dim": [5, 10]}))
    back = ck.load_state_dict(str(tmp_path / "unet"))
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    cfg = ck.load_config(str(tmp_path / "unet"))
    assert ck._heads_from_config(cfg, cfg["block_out_channels"]) == (5, 10)
    assert ck._heads_from_config({}, (320, 640)) == (5, 10) and ck._heads_from_config({"attention_head_dim": 8}, (64, 64)) == (8, 8)
    full = ck.with_support_stem(back)
    assert full["conv_in_ref.weight"].shape == (8, 8, 3, 3)
    assert torch.equal(full["conv_in_ref.weight"][:, :4], sd["conv_in.weight"] / 2)
    assert torch.equal(full["conv_in_ref.weight"][:, 4:], sd["conv_in.weight"] / 2)
    assert ck.with_support_stem(full) is full
    torch.save(sd, str(tmp_path / "unet" / "w.bin"))
    assert torch.equal(ck.load_state_dict(str(tmp_path / "unet" / "w.bin"))["conv_in.bias"], sd["conv_in.bias"])


def _gloo_loader_worker(rank, world, port, tree, q):
    import sys
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from diffews_b200 import data
    data.FSSDataset.initialize(48, tree, False)
    np.random.seed(0)
    loader = data.FSSDataset.build_dataloader("coco", 2, 1, 0, "val", 1, device="cpu")   # rank / world from the group
    names = []
    for i, raws in enumerate(loader.raw_batches()):
        names.append([(r["query_name"], r["support_names"][0], r["class_sample"]) for r in raws])
        if i == 1:
            break
    gathered = [None] * world
    dist.all_gather_object(gathered, names)
    q.put((rank, loader.rank, loader.world, len(loader), gathered))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_loader_shards_by_process_group_rank(tmp_path):
    """world_size-2 gloo: EpisodeLoader picks rank / world up from torch.distributed, every rank walks the reference's
    numpy episode sequence and keeps batches rank, rank + 2, ...; interleaving the ranks' batches gives exactly the
    single-process sequence (so the all-reduced counts equal a one-GPU run of the same seed)."""
    import socket
    import sys
    import numpy as np
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import data_tree
    from diffews_b200 import data
    tree = str(tmp_path)
    data_tree.build_coco_tree(tree)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_gloo_loader_worker, args=(r, 2, port, tree, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
    data.FSSDataset.initialize(48, tree, False)
    np.random.seed(0)
    single = data.FSSDataset.build_dataloader("coco", 2, 1, 0, "val", 1, device="cpu", rank=0, world=1)
    want = []
    for i, raws in enumerate(single.raw_batches()):
        want.append([(r["query_name"], r["support_names"][0], r["class_sample"]) for r in raws])
        if i == 3:
            break
    for rank, lr, lw, n, gathered in res:
        assert (lr, lw) == (rank, 2) and n == 250
        assert gathered[0] == [want[0], want[2]] and gathered[1] == [want[1], want[3]]


def test_random_init_state_dicts_match_oracle_module_shapes():
    """diffews_b200.synthetic.random_{unet,vae}_state_dict (bench.py's weight source, nothing from oracle/) carry exactly the
    keys and shapes of the oracle's SD-2.1 modules = the diffusers state-dict layout (865 910 724 + 23 360 / 83 653 863
    parameters)."""
    from diffews_b200.synthetic import random_unet_state_dict, random_vae_state_dict
    from oracle.sd21 import build_models
    for widths in (((320, 640, 1280, 1280), (5, 10, 20, 20), (128, 256, 512, 512)), ((64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))):
        unet_o, vae_o = build_models(0, *widths)
        for mine, ref in ((random_unet_state_dict(0, widths[0]), unet_o.state_dict()), (random_vae_state_dict(1, widths[2]), vae_o.state_dict())):
            assert set(mine) == set(ref), (sorted(set(mine) ^ set(ref))[:8])
            for k, v in ref.items():
                assert tuple(mine[k].shape) == tuple(v.shape), k
    n_unet = sum(v.numel() for v in random_unet_state_dict(0).values())
    n_vae = sum(v.numel() for v in random_vae_state_dict(1).values())
    assert n_unet == 865_910_724 + 23_360 and n_vae == 83_653_863


def _gloo_reducer_worker(rank, world, port, q):
    """The DDP bucket reducer of the training step (diffews_b200/train.py) on CPU tensors: the first step learns how many
    gradient contributions each parameter receives and reduces at the end; from the second step on a bucket is reduced
    the moment its last parameter is complete, i.e. while later (here: simulated) backward work is still outstanding."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from diffews_b200.train import GradReducer, ParamStore
    st = ParamStore("cpu", torch.float16)
    shapes = [(40, 24), (24,), (64, 40), (64,), (16, 64), (16,)]           # registration = execution order
    ps = [st.add(f"p{i}", torch.zeros(s), "linear" if len(s) == 2 else "vec") for i, s in enumerate(shapes)]
    st.finalize()
    st.reducer = GradReducer(st, None, bucket_bytes=4096)
    st.grad_scale = 1.0 / world
    assert len(st.reducer.buckets) >= 2
    uses = [2, 2, 2, 2, 1, 1]                                               # the last layer is only on the query pass
    out = []
    for step in range(3):
        st.begin_step()
        launched_mid = 0
        for use in range(2):                                                # query-pass backward, then support-pass backward
            for p, u in reversed(list(zip(ps, uses))):
                if use >= u:
                    continue
                g = torch.full(p.shape, float(rank + 1) * (step + 1))
                p.add_grad(g, st.grad_scale, st)
            launched_mid = max(launched_mid, sum(st.reducer.launched))
        overlapped = st.reducer.overlapped
        st.end_step()
        out.append((overlapped, [float(p.grad.flatten()[0]) for p in ps], [p.expected for p in ps]))
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_bucket_reducer_world2():
    import socket
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_gloo_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    uses = [2, 2, 2, 2, 1, 1]
    for rank, out in res:
        for step, (overlapped, vals, expected) in enumerate(out):
            assert expected == uses
            # mean over ranks of (rank + 1) * (step + 1) * uses
            want = [1.5 * (step + 1) * u for u in uses]
            assert vals == pytest.approx(want), (rank, step, vals, want)
            if step == 0:
                assert overlapped == 0            # counts not known yet: everything reduced at the end
            else:
                assert overlapped >= 1            # at least one bucket went out before the backward had finished
