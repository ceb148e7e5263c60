"""One optimisation step of the DiffewS UNet on the B200 kernels at the reference's training shape (BASELINE config 4:
7-shot, 512^2 -> 64x64 latents, batch 1: train_icl_multitask_nocrop_nearest_nshot_v3.py:1337-1396), full SD-2.1 width,
random-init weights, synthetic latents.  Timed with CUDA events over K steps after W warm-up steps; one more step runs
under the per-launch timer for the kernel-family breakdown.
  python scripts/bench_train.py [--nshot 7] [--latent 64] [--steps 5] [--warmup 3] [--out profiles/r02_train_step.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402
from diffews_b200.synthetic import random_unet_state_dict  # noqa: E402
from diffews_b200.train import Trainer, TrainableUNet  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nshot", type=int, default=7)
    ap.add_argument("--latent", type=int, default=64)
    ap.add_argument("--lctx", type=int, default=77)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default=None)
    ap.add_argument("--graph", action="store_true", help="capture forward + backward into a CUDA graph")
    ap.add_argument("--layer-table", default=None, help="write the per-shape table of the instrumented step here")
    args = ap.parse_args()
    dev = "cuda"
    sd = random_unet_state_dict(0)
    unet = TrainableUNet(sd, device=dev)
    tr = Trainer(unet, lr=1e-5, max_grad_norm=1.0, loss_scale=1024.0)
    g = torch.Generator(device=dev).manual_seed(0)
    k, hw = args.nshot, args.latent
    lat_ref = torch.randn(k, 8, hw, hw, device=dev, generator=g)
    lat_tag = torch.randn(1, 4, hw, hw, device=dev, generator=g)
    target = torch.randn(1, 4, hw, hw, device=dev, generator=g)
    ehs = torch.randn(1, args.lctx, 1024, device=dev, generator=g)
    step_args = (lat_ref, lat_tag, target, ehs, 1.0)
    losses = []
    if args.graph:
        tr.enable_cuda_graph(*step_args)
    for _ in range(args.warmup):
        losses.append(float(tr.step(*step_args)))
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.step(*step_args)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (ops.launch_count() - n0) // args.steps
    losses.append(float(loss))
    # phase split (events around forward+backward / clip+AdamW / operand refresh of one more step)
    tr.forward_backward(*step_args)          # (first eager pass after a capture re-creates the eager workspaces)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    tr.forward_backward(*step_args)
    ev[1].record()
    tr.opt.clip_grad_norm_(1.0); tr.opt.step()
    ev[2].record()
    unet.refresh_operands()
    ev[3].record()
    torch.cuda.synchronize()
    phases = {"forward_backward_ms": ev[0].elapsed_time(ev[1]), "clip_adamw_ms": ev[1].elapsed_time(ev[2]),
              "operand_refresh_ms": ev[2].elapsed_time(ev[3])}
    timer = ops.KernelTimer()
    ops.set_timer(timer)
    tr.forward_backward(*step_args)          # eager launches (a graph replay has no per-launch events)
    tr.opt.clip_grad_norm_(1.0); tr.opt.step()
    unet.refresh_operands()
    torch.cuda.synchronize()
    ops.set_timer(None)
    fam = {}
    for kind, d in timer.summary().items():
        e = {"launches": d["launches"], "ms": round(d["ms"], 3)}
        if kind in ("igemm", "attn"):
            e["tflops"] = round(d["flops"] / d["ms"] / 1e9, 1) if d["ms"] > 0 else None
        else:
            e["gb_per_s"] = round(d["flops"] / d["ms"] / 1e6, 1) if d["ms"] > 0 else None
        fam[kind] = e
    wg = {"launches": 0, "ms": 0.0, "flops": 0.0}
    for (kind, shape), d in timer.by_shape().items():
        if shape and shape.startswith("wgrad"):
            wg["launches"] += d["launches"]; wg["ms"] += d["ms"]; wg["flops"] += d["flops"]
    if args.layer_table:
        rows = sorted(timer.by_shape().items(), key=lambda kv: -kv[1]["ms"])
        tot = sum(v["ms"] for _, v in rows)
        with open(args.layer_table, "w") as f:
            f.write(f"# per-shape CUDA-event times of one instrumented training step ({k}-shot, latent {hw}); summed kernel time {tot:.1f} ms\n")
            f.write("kind\tshape\tlaunches\tms_total\tshare\tTFLOP/s (igemm, attn) | TB/s of algorithmic bytes (others)\n")
            for (kind, shape), v in rows:
                r = v["flops"] / (v["ms"] * 1e9) if v["ms"] > 0 else 0.0
                f.write(f"{kind}\t{shape}\t{v['launches']}\t{v['ms']:.3f}\t{v['ms'] / tot:.4f}\t{r:.2f}\n")
    tensor_flops = sum(d["flops"] for kind, d in timer.summary().items() if kind in ("igemm", "attn"))
    line = {
        "what": "one training step of the DiffewS UNet (support pass + query pass + backward through both + clip + AdamW)",
        "config": {"workload": f"{k}-shot, {hw}x{hw} latents ({hw * 8}^2 images), batch 1, SD-2.1 width, Lctx {args.lctx}, fp16 "
                               "operands / activations, fp32 master weights (BASELINE config 4 shape)", "data": "synthetic"},
        "launch": "CUDA graph (forward + backward), eager optimizer" if args.graph else "eager",
        "ms_per_step": round(ms, 3), "steps": args.steps, "warmup": args.warmup, "launches_per_step": int(launches),
        "phases": {k_: round(v, 3) for k_, v in phases.items()},
        "tensor_tflop_per_step": round(tensor_flops / 1e12, 3),
        "tensor_tflops_whole_step": round(tensor_flops / ms / 1e9, 1),
        "kernels": fam,
        "wgrad": {"launches": wg["launches"], "ms": round(wg["ms"], 3),
                  "tflops": round(wg["flops"] / wg["ms"] / 1e9, 1) if wg["ms"] > 0 else None},
        "losses": [round(l, 5) for l in losses],
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
    }
    s = json.dumps(line)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
