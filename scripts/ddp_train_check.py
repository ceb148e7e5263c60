"""torchrun --nproc-per-node 2 scripts/ddp_train_check.py : data-parallel training step on 2 GPUs (NCCL).
Each rank runs forward + backward on its OWN synthetic episode; the bucket reducer all-reduces the gradients during the
backward.  Rank 0 then recomputes both ranks' gradients locally (same kernels, deterministic) and checks that the reduced
flat gradient buffer equals their sum bit for bit, that at least one bucket was launched before the backward had finished,
and that the parameters of both ranks are identical after the optimizer step.
ref: accelerator.prepare(unet, ...) -> DistributedDataParallel, train_icl_multitask_nocrop_nearest_nshot_v3.py:1226, :1386."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200.synthetic import random_unet_state_dict  # noqa: E402
from diffews_b200.train import Trainer, TrainableUNet  # noqa: E402

CH, HEADS = (64, 128, 256, 256), (1, 2, 4, 4)


def inputs(rank, k=2, hw=16):
    g = torch.Generator().manual_seed(100 + rank)
    return (torch.randn(k, 8, hw, hw, generator=g).cuda(), torch.randn(1, 4, hw, hw, generator=g).cuda(),
            torch.randn(1, 4, hw, hw, generator=g).cuda(), torch.randn(1, 5, 1024, generator=g).cuda(), 1.0)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    full = os.environ.get("DDP_FULL_WIDTH") == "1"
    sd = random_unet_state_dict(0) if full else random_unet_state_dict(0, CH)
    kw = {} if full else dict(block_out_channels=CH, heads=HEADS)
    unet = TrainableUNet(sd, device="cuda", **kw)
    tr = Trainer(unet, lr=1e-4, loss_scale=256.0, process_group=dist.group.WORLD, bucket_bytes=(64 << 20) if full else (1 << 20))
    assert tr.world == world and unet.store.reducer is not None
    hw = 32 if full else 16
    mine = inputs(rank, hw=hw)
    report = {"world": world, "buckets": len(unet.store.reducer.buckets)}
    for step in range(3):
        tr.forward_backward(*mine)
        torch.cuda.synchronize()
        if step == 0:
            assert unet.store.reducer.overlapped == 0
        else:
            report["overlapped_buckets"] = unet.store.reducer.overlapped
            assert unet.store.reducer.overlapped >= 1
    reduced = unet.store.flat_g.clone()
    # local recomputation of every rank's contribution, no reducer
    red, unet.store.reducer = unet.store.reducer, None
    acc = torch.zeros_like(reduced)
    for r in range(world):
        tr.forward_backward(*inputs(r, hw=hw))
        acc += unet.store.flat_g
    unet.store.reducer = red
    torch.cuda.synchronize()
    report["bit_exact"] = bool(torch.equal(acc, reduced))
    report["max_abs_diff"] = float((acc - reduced).abs().max())
    assert report["max_abs_diff"] <= 1e-6 * float(reduced.abs().max()), report
    # timing of the backward with and without the overlapped reduction
    def timed(n=5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            tr.forward_backward(*mine)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    report["fwd_bwd_ms_with_allreduce"] = round(timed(), 3)
    red, unet.store.reducer = unet.store.reducer, None
    report["fwd_bwd_ms_no_allreduce"] = round(timed(), 3)
    unet.store.reducer = red
    report["grad_mbytes"] = round(unet.store.total * 4 / 2 ** 20, 1)
    # one optimizer step: identical parameters on every rank
    tr.step(*mine)
    chk = unet.store.flat_w.double().sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    report["params_identical"] = all(float(b) == float(both[0]) for b in both)
    assert report["params_identical"]
    if rank == 0:
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
