#!/usr/bin/env python
"""bench.py — episodes/sec of the DiffewS hot path (1-shot 512^2, single-step UNet) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                     (the reference's CPU path = the oracle port, host cores)

A "step" is one pass of the full hot path over one batch of B synthetic episodes per GPU (BASELINE config 2: B = 16,
1-shot, 512x512): 3 VAE encodes, support + query UNet pass with the KV bank, z0 = -v, VAE decode, uint8, rthres,
intersection/union, class-indexed accumulation.  Prints ONE JSON line (rank 0).

  value      episodes/s, whole job, inputs resident in HBM (device leg)
  e2e        episodes/s through the public API with inputs in pinned HOST memory (H2D of the batch and D2H of the
             per-episode counts inside the timed region)
  roofline   the dominant kernel family (tcgen05 implicit-GEMM conv/linear): algorithmic FLOPs of the launches in the
             timed region / their CUDA-event time, against MEASURED_PEAKS.json's sustained bf16 TFLOP/s
  cpu_baseline  the oracle (kind "port": the reference cannot be imported) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "episodes/sec (1-shot, 512^2, 1-step UNet)"
UNIT = "episodes/s"
FLOPS_PER_EPISODE_1SHOT_512 = 7.58e12
# BASELINE.json configs -> (episodes per GPU per step, shots, image size, algorithmic TFLOP per episode: SURVEY §8d)
CONFIGS = {2: (16, 1, 512, 7.58e12), 3: (8, 5, 512, 20.19e12), 5: (8, 1, 768, 18.48e12)}
TENSOR_FLOP_PER_CLK_PER_SM = 8192.0       # dense 16-bit tcgen05: 4096 MAC / clk / SM


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]            # samples under load: the upper half
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_oracle_episode_time(size: int, nshot: int, threads: int, unet_o=None, vae_o=None, episodes: int = 3):
    """`episodes` full episodes through the CPU oracle (bsz=1, like the reference's eval loop); mean seconds per episode
    (about 15 s of CPU work at 512^2: the bounded sample of the contract)."""
    import torch
    from diffews_b200.synthetic import make_batch, prompt_embedding
    from oracle.pipeline import evaluate_episode
    from oracle.sd21 import build_models
    torch.set_num_threads(threads)
    if unet_o is None:
        unet_o, vae_o = build_models(0)
    emb = prompt_embedding()
    warm = make_batch(10_000, 1, 64, nshot)
    evaluate_episode(unet_o, vae_o, emb, warm)          # warm-up at 64^2 (oneDNN primitive caches, thread pool)
    batches = [make_batch(i, 1, size, nshot) for i in range(episodes)]
    t0 = time.perf_counter()
    for batch in batches:
        evaluate_episode(unet_o, vae_o, emb, batch)
    return (time.perf_counter() - t0) / episodes


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port (the reference needs
    diffusers/xformers/accelerate, none installable here), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from diffews_b200.synthetic import make_batch, prompt_embedding
    from oracle.pipeline import evaluate_episode
    from oracle.sd21 import build_models
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet_o, vae_o = build_models(0)
    emb = prompt_embedding()
    budget_s = 240.0
    t_start = time.perf_counter()
    evaluate_episode(unet_o, vae_o, emb, make_batch(10_000, 1, 64, args.nshot))
    times = []
    done_warm = 0
    for i in range(args.warmup + args.steps):
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
        batch = make_batch(i, 1, args.size, args.nshot)
        t0 = time.perf_counter()
        evaluate_episode(unet_o, vae_o, emb, batch)
        dt = time.perf_counter() - t0
        if done_warm < min(args.warmup, 1):      # CPU steps are ~tens of seconds each: at most one untimed warm-up
            done_warm += 1
            continue
        times.append(dt)
    ms = 1000.0 * sum(times) / len(times)
    value = 1000.0 / ms
    sample = (f"{len(times)} timed episode(s) of the same workload at bsz=1 (the only batch size the reference's eval "
              f"loop supports), fp32 oracle port, {cores} threads; capped at {budget_s:.0f}s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": done_warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.nshot}-shot {args.size}x{args.size} episodes, single-step SD-2.1 UNet + VAE "
                               "+ rthres/IoU, random-init weights", "episodes_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


def _ncu_attn_pct():
    """sm__pipe_tensor_subpipe_hmma_cycles_active of the KV-fused attention kernel from the committed ncu summary (a
    stand-alone capture; never a number measured under the profiler in this run)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_ncu_attn.tsv")
    try:
        rows = [l.rstrip("\n").split("\t") for l in open(path)]
        col = next(i for i, h in enumerate(rows[0]) if h.startswith("sm__pipe_tensor_subpipe_hmma_cycles_active"))
        vals = [float(r[col]) for r in rows[1:] if "attn_kvfused" in r[0]]
        return round(sum(vals) / len(vals), 1)
    except (OSError, StopIteration, ValueError, ZeroDivisionError):
        return None


def _loader_leg(args, runner, dev, rank, world, B, barrier, dist):
    """episodes/s with every step's inputs coming from encoded files through diffews_b200.data.EpisodeLoader."""
    import shutil
    import tempfile
    import numpy as np
    import torch
    from PIL import Image
    from diffews_b200.data import DatasetCOCO, EpisodeLoader, EpisodeTransform
    threads = args.loader_threads or max(2, min(16, (os.cpu_count() or 8) // world))
    root = tempfile.mkdtemp(prefix=f"dfw_bench_tree_r{rank}_")
    try:
        # a COCO-20i-layout tree (fold 0 validation classes), 640x480 JPEG photos (quality 90) + PNG class masks
        rs = np.random.RandomState(7)
        base = os.path.join(root, "COCO2014")
        os.makedirs(os.path.join(base, "val2014")); os.makedirs(os.path.join(base, "annotations", "val2014"))
        os.makedirs(os.path.join(base, "splits", "val"))
        classes = [4 * v for v in range(20)]
        classwise = {c: [] for c in classes}
        n_img = 96
        yy, xx = np.mgrid[0:480, 0:640]
        for i in range(n_img):
            img = np.stack([127 + 100 * np.sin(xx / rs.uniform(5, 60) + rs.uniform(0, 6)) * np.cos(yy / rs.uniform(5, 60))
                            for _ in range(3)], -1) + rs.uniform(-20, 20, (480, 640, 3))
            name = f"val2014/COCO_val2014_{i:012d}.jpg"
            Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(base, name), format="JPEG", quality=90)
            lab = np.zeros((480, 640), np.uint8)
            for c in (classes[i % 20], classes[(i * 7 + 3) % 20]):
                y0, x0 = rs.randint(0, 300), rs.randint(0, 400)
                lab[y0:y0 + rs.randint(60, 180), x0:x0 + rs.randint(60, 240)] = c + 1
                classwise[c].append(name)
            Image.fromarray(lab).save(os.path.join(base, "annotations", name[:-4] + ".png"), format="PNG")
        import pickle
        with open(os.path.join(base, "splits", "val", "fold0.pkl"), "wb") as f:
            pickle.dump(classwise, f)
        np.random.seed(1234)
        ds = DatasetCOCO(root, fold=0, transform=EpisodeTransform(args.size), split="val", shot=args.nshot,
                         use_original_imgsize=False)
        loader = EpisodeLoader(ds, bsz=B, device=dev, decode_threads=threads, rank=rank, world=world)

        def batches():
            while True:
                for b in loader:
                    if b["query_img"].shape[0] == B:
                        yield b
        it = batches()
        host_out = torch.empty((2, B, 2), dtype=torch.int64).pin_memory()

        def run(nsteps):
            pending = None
            for _ in range(nsteps):
                batch = next(it)                                    # decode (thread pool) + one H2D copy + 3 launches
                inter, union = runner.step(batch)
                if pending is not None:                             # read the PREVIOUS step's counts: one step in flight
                    pending.synchronize()
                    _ = int(host_out[0, 0, 1])
                host_out[0].copy_(inter, non_blocking=True)
                host_out[1].copy_(union, non_blocking=True)
                pending = torch.cuda.Event()
                pending.record()
            pending.synchronize()
            _ = int(host_out[0, 0, 1])
        run(2)
        barrier()
        t0 = time.perf_counter()
        run(args.steps)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
        t = torch.tensor([wall_ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_ep = (1 + args.nshot) * (480 * 640 * 3 + 480 * 640)
        return {"value": world * B * args.steps / (float(t.item()) / 1000.0), "unit": UNIT, "decode_threads_per_rank": threads,
                "host_cores": os.cpu_count(), "h2d_bytes_per_step": int(B * per_ep),
                "input": f"{n_img} JPEG 640x480 (quality 90) + PNG class masks per rank, COCO-20i layout, PIL decode on a "
                         "thread pool, resize / normalise / mask on the GPU (dfw_resize_normalize_u8, dfw_mask_nearest)"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def _igemm_traffic():
    """DRAM bytes of one launch of the dominant igemm shape, from the committed ncu --set full capture (GB), else None."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_igemm_traffic.json")) as f:
            t = json.load(f)
        return {"unit": "GB", "dram_per_launch": round(t["dram_bytes_per_launch"] / 1e9, 3),
                "algorithmic_per_launch": round(t["algorithmic_bytes_per_launch"] / 1e9, 3), "layer": t["layer"],
                "source": "profiles/r02_ncu_t128.tsv"}
    except Exception:
        return None


_REAL_STDOUT_FD = None


def _quiet_stdout():
    """stdout carries exactly one JSON line: everything libraries print there meanwhile (NCCL's version banner, ...) goes
    to stderr; _emit() restores the real stdout for the line itself."""
    global _REAL_STDOUT_FD
    if _REAL_STDOUT_FD is None:
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT_FD is not None:
        os.dup2(_REAL_STDOUT_FD, 1)
    print(json.dumps(line), flush=True)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="episodes per GPU per step (BASELINE config 2: 16)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--nshot", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernel-timer", action="store_true")
    ap.add_argument("--no-loader", action="store_true", help="skip the loader-fed e2e leg (JPEG / PNG files -> EpisodeLoader)")
    ap.add_argument("--loader-threads", type=int, default=0, help="decode threads per rank (0: host cores / ranks, at most 16)")
    ap.add_argument("--vae-stream", default="half", choices=["half", "f32"], help="VAE residual-stream storage")
    ap.add_argument("--unet-stream", default="half", choices=["half", "f32"], help="UNet residual-stream storage")
    ap.add_argument("--gn-fusion", default="off", choices=["auto", "off", "all"],
                    help="GroupNorm + SiLU folded into the consuming VAE convolution (layers.FUSE_GN_INTO_CONV)")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch on the hot kernels (DFW_OPT_PDL)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="set a library option before the engines are built, e.g. --opt B_RESIDENT=0 (diffews_b200._lib.OPT_*)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of one CUDA graph per step")
    ap.add_argument("--layer-table", default=None, help="write a per-shape table of the timed tensor-core launches here")
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 = 1-shot 512^2 B16 (default workload), 3 = 5-shot 512^2 B8, 5 = 1-shot 768^2 B8")
    ap.add_argument("--operands", default="f16", choices=["f16", "bf16", "f32"],
                    help="tensor-core operand format; f32 = the fp32 evaluation mode (split-f16 operands, fp32 everywhere else)")
    args = ap.parse_args()
    if args.config is not None:
        args.batch, args.nshot, args.size, _ = CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the single JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    if args.warmup < 3:
        args.warmup = 3

    from diffews_b200 import ops
    from diffews_b200 import layers as _layers
    _layers.FUSE_GN_INTO_CONV = {"auto": "auto", "off": False, "all": True}[args.gn_fusion]
    from diffews_b200 import _lib as _dlib
    if args.pdl:
        ops.set_option(_dlib.OPT_PDL, 1)
    for kv in args.opt:
        name, val = kv.split("=")
        ops.set_option(getattr(_dlib, "OPT_" + name.upper()), int(val))
    from diffews_b200.runner import EpisodeRunner, build_engine_from_state_dicts
    from diffews_b200.synthetic import make_batch, prompt_embedding, random_unet_state_dict, random_vae_state_dict

    # weight source: deterministic random init with the diffusers key names (no checkpoints offline); nothing from oracle/
    from diffews_b200.layers import Precision
    half = torch.bfloat16 if args.operands == "bf16" else torch.float16
    if args.operands == "f32":
        vae_prec = unet_prec = Precision(half=half, f32=True)
        args.vae_stream = args.unet_stream = "f32"
    else:
        vae_prec = Precision(half=half, stream_f32=(args.vae_stream == "f32"), mid_f32=False)
        unet_prec = Precision(half=half, stream_f32=(args.unet_stream == "f32"), mid_f32=(args.unet_stream == "f32"))
    pipe = build_engine_from_state_dicts(random_unet_state_dict(0), random_vae_state_dict(1), prompt_embedding(), device=dev,
                                         vae_precision=vae_prec, unet_precision=unet_prec)
    runner = EpisodeRunner(pipe, "coco", img_size=args.size)
    B = args.batch

    # a small pool of distinct episode batches (rank-disjoint), pinned on the host
    pool = 2
    host_batches = []
    for i in range(pool):
        hb = make_batch((rank * pool + i) * B, B, args.size, args.nshot)
        host_batches.append({k: v.pin_memory() for k, v in hb.items()})
    dev_batches = [{k: v.to(dev, non_blocking=True) for k, v in hb.items()} for hb in host_batches]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device leg: inputs resident in HBM ------------------------------------------------------------
    use_graph = not args.no_graph
    if use_graph:
        runner.enable_cuda_graph(dev_batches[0])
    for i in range(args.warmup):
        runner.step(dev_batches[i % pool])
    barrier()
    n_before = ops.launch_count()
    runner.step(dev_batches[0])
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - n_before           # 0 in graph mode: replays do not pass through the C ABI
    if use_graph:
        launches_per_step = runner.launches_in_graph
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        runner.step(dev_batches[i % pool])
    if world > 1:
        runner.meter.all_reduce()
    ev1.record()
    barrier()
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1000.0)

    # ---------------- e2e leg: public API, host buffers, H2D + D2H inside the timed region ----------------------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
        d2h = 2 * B * 2 * 8
        host_out = torch.empty((2, B, 2), dtype=torch.int64).pin_memory()

        # Every step's inputs travel host -> device inside the timed region and its counts device -> host; with the CUDA
        # graph the copy of batch i+1 is started right after step i is launched (EpisodeRunner.prefetch), so it runs under
        # that step's kernels instead of in front of the next one.
        pipelined = use_graph and hasattr(runner, "prefetch")

        def e2e_run(nsteps):
            if pipelined:
                runner.prefetch(host_batches[0])
            for i in range(nsteps):
                if pipelined:
                    inter, union = runner.step_prefetched()
                    if i + 1 < nsteps:
                        runner.prefetch(host_batches[(i + 1) % pool])
                else:
                    inter, union = runner.step(host_batches[i % pool])  # pinned host tensors: H2D happens inside step()
                host_out[0].copy_(inter, non_blocking=True)
                host_out[1].copy_(union, non_blocking=True)
                torch.cuda.current_stream().synchronize()               # the caller reads the counts every step
                _ = int(host_out[0, 0, 1])
        e2e_run(2)
        barrier()
        t0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        e2e_run(args.steps)
        ee1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        e_ms = max(ee0.elapsed_time(ee1), wall_ms)               # host-synchronous loop: wall clock is the honest one
        te = torch.tensor([e_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * args.steps / (float(te.item()) / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}
        # same loop, but the caller also takes the uint8 segmentation image the reference __call__ hands back
        # (pipeline:534-545) off the device every step: B x 3 x S x S bytes D2H into pinned memory
        seg_host = torch.empty((B, 3, args.size, args.size), dtype=torch.uint8).pin_memory()

        def e2e_image_run(nsteps):
            if pipelined:
                runner.prefetch(host_batches[0])
            for i in range(nsteps):
                if pipelined:
                    inter, union = runner.step_prefetched()
                    if i + 1 < nsteps:
                        runner.prefetch(host_batches[(i + 1) % pool])
                else:
                    inter, union = runner.step(host_batches[i % pool])
                host_out[0].copy_(inter, non_blocking=True)
                host_out[1].copy_(union, non_blocking=True)
                seg_host.copy_(runner.last_seg_u8, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                _ = int(host_out[0, 0, 1]) + int(seg_host[0, 0, 0, 0])
        e2e_image_run(2)
        barrier()
        t0 = time.perf_counter()
        e2e_image_run(args.steps)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1000.0
        barrier()
        ti = torch.tensor([wall_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ti, op=dist.ReduceOp.MAX)
        e2e["with_uint8_image_d2h"] = {"value": world * B * args.steps / (float(ti.item()) / 1000.0), "unit": UNIT,
                                       "d2h_bytes_per_step": int(d2h + seg_host.numel())}

    # ---------------- loader-fed e2e leg: encoded image FILES on disk -> EpisodeLoader -> runner ------------------------------
    # The caller-side data format of the reference (evaluation_util/data/coco.py): JPEG images + PNG class masks, decoded by
    # PIL on a thread pool, shipped as raw bytes, resized / normalised / masked on the GPU, then the same graph step.  The
    # counts of step i are read back while step i+1 runs (one step in flight), so decoding overlaps the GPU.
    if e2e is not None and not args.no_loader:
        try:
            e2e["loader_fed"] = _loader_leg(args, runner, dev, rank, world, B, barrier, dist if world > 1 else None)
        except Exception as ex:                                   # e.g. a PIL build without a JPEG encoder
            e2e["loader_fed"] = {"unavailable": f"{type(ex).__name__}: {ex}"}

    # ---------------- roofline of the dominant kernel family ----------------------------------------------------------
    # Per-launch CUDA events cannot be recorded inside a graph replay, so the tensor-core launches are timed in an
    # instrumented pass of the same steps launched eagerly (same kernels, same shapes, same stream) right after.
    peak_tf, peak_gbs, peak_kind = _peaks()
    roofline = None
    kernels = None
    timer = None if args.no_kernel_timer else ops.KernelTimer()
    if timer is not None:
        runner_graph, runner._graph = getattr(runner, "_graph", None), None
        runner.step(dev_batches[0])
        barrier()
        ops.set_timer(timer)
        # Eager launches are issued by Python at ~30-60 us per op: for kernels shorter than that the GPU would drain its
        # queue and the event pair would time the host, not the kernel.  A 25 ms device-side sleep at the head of every
        # step lets the host run ahead (the launch queue holds the backlog), so the events bracket back-to-back GPU
        # work; the sleeps are timed and subtracted.
        sleep_cycles = int(25e-3 * torch.cuda.get_device_properties(dev).clock_rate * 1e3) if hasattr(
            torch.cuda.get_device_properties(dev), "clock_rate") else int(25e-3 * 1.9e9)
        sleeps = []
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for i in range(args.steps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(); torch.cuda._sleep(sleep_cycles); s1.record()
            sleeps.append((s0, s1))
            runner.step(dev_batches[i % pool])
        t1e.record()
        barrier()
        ops.set_timer(None)
        runner._graph = runner_graph
        ms_total_instr = t0e.elapsed_time(t1e) - sum(a.elapsed_time(b) for a, b in sleeps)
    if timer is not None:
        summ = timer.summary()
        # shares are taken over the GPU-busy time of the instrumented pass (sum of the per-launch event times): on a box with
        # slow host cores the eager pass has launch gaps, which say nothing about the kernels
        instr_wall_ms = ms_total_instr
        ms_total_instr = sum(v["ms"] for v in summ.values())
        kernels = {}
        for k, v in summ.items():
            rate = (v["flops"] / (v["ms"] * 1e9) if v["ms"] > 0 else None)          # TFLOP/s, or (memory kinds) TB/s
            d = {"launches": v["launches"], "ms": round(v["ms"], 3), "share_of_step": round(v["ms"] / ms_total_instr, 3)}
            if k in ops.MEM_KINDS: d["gb_per_s"] = round(rate * 1e3, 1) if rate else None   # algorithmic bytes
            else: d["tflops"] = round(rate, 1) if rate else None
            kernels[k] = d
        if args.layer_table and rank == 0:
            rows = sorted(timer.by_shape().items(), key=lambda kv: -kv[1]["ms"])
            with open(args.layer_table, "w") as f:
                f.write(f"# per-shape CUDA-event times over {args.steps} instrumented eager steps (B={B}/GPU); total {ms_total_instr:.1f} ms\n")
                f.write("kind\tshape\tlaunches\tms_total\tshare\tTFLOP/s (igemm, attn) | TB/s of algorithmic bytes (others)\n")
                for (kind, shape), v in rows:
                    tf = v["flops"] / (v["ms"] * 1e9) if v["ms"] > 0 else 0.0
                    f.write(f"{kind}\t{shape}\t{v['launches']}\t{v['ms']:.3f}\t{v['ms'] / ms_total_instr:.4f}\t"
                            + (f"{tf:.2f}" if kind in ops.MEM_KINDS else f"{tf:.1f}") + "\n")
        ig = summ.get("igemm")
        # the single dominant kernel: igemm_t128_kernel (channel-major tcgen05 implicit GEMM: every VAE conv with
        # Cout % 128 == 0 and the residual projections).  ops tags its launches in the instrumented pass.
        t128 = {"launches": 0, "ms": 0.0, "flops": 0.0}
        for (kind, shape), v in timer.by_shape().items():
            if kind == "igemm" and shape and shape.startswith("t128 "):
                t128["launches"] += v["launches"]; t128["ms"] += v["ms"]; t128["flops"] += v["flops"]
        if ig and ig["ms"] > 0:
            fam = ig["flops"] / (ig["ms"] * 1e9)
            dom = t128 if t128["ms"] > 0 else ig
            achieved = dom["flops"] / (dom["ms"] * 1e9)
            roofline = {"bound": "tensor",
                        "kernel": ("igemm_t128_kernel (tcgen05 implicit-GEMM conv / projection, channel-major accumulator)"
                                   if dom is t128 else "igemm_kernel<BLOCK_N> (tcgen05 implicit-GEMM conv/linear)"),
                        "achieved": round(achieved, 1), "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": round(achieved / peak_tf, 4), "traffic": _igemm_traffic(), "peak_source": f"{peak_kind} sustained bf16",
                        "launches": dom["launches"], "share_of_step": round(dom["ms"] / ms_total_instr, 3),
                        "avg_launch_ms": round(dom["ms"] / max(dom["launches"], 1), 4),
                        "algorithmic_tflop_per_launch": round(dom["flops"] / max(dom["launches"], 1) / 1e12, 4),
                        "timed_in": "instrumented eager pass of the same steps (events around every launch); shares over the "
                                    f"summed kernel time {ms_total_instr:.1f} ms (wall of that pass {instr_wall_ms:.1f} ms)",
                        "family_all_igemm_launches": {"kernels": "igemm_t128_kernel + igemm_kernel<16|128|160|256>",
                                                      "achieved": round(fam, 1), "frac": round(fam / peak_tf, 4),
                                                      "launches": ig["launches"],
                                                      "share_of_step": round(ig["ms"] / ms_total_instr, 3)}}

    # ---------------- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ---------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        del pipe
        secs = cpu_oracle_episode_time(args.size, args.nshot, cores)
        cpu = {"value": 1.0 / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"3 episodes ({args.nshot}-shot {args.size}x{args.size}, bsz=1, fp32 oracle port of the reference "
                         f"path) after a 64x64 warm-up: {secs:.1f}s per episode"}

    if rank == 0:
        cfg_id = next((k for k, v in CONFIGS.items() if (B, args.nshot, args.size) == v[:3]), None)
        flops_ep = CONFIGS[cfg_id][3] if cfg_id is not None else None
        # attention: achieved rate of the KV-fused kernel inside the step, as TFLOP/s and as the fraction of the tensor pipe's
        # cycles that rate occupies at the SM clock sampled during the run (4096 dense MAC / clk / SM) -- the live counterpart
        # of ncu's sm__pipe_tensor_cycles_active (profiles/r02_ncu_attn.tsv, captured stand-alone)
        attn = None
        hbm_families = None
        if kernels is not None:
            ka = kernels.get("attn")
            if ka and ka.get("tflops"):
                sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                pipe_peak_tf = TENSOR_FLOP_PER_CLK_PER_SM * sms * sm_mhz * 1e6 / 1e12
                attn = {"tflops_in_step": ka["tflops"], "ms_per_step": round(ka["ms"] / args.steps, 3),
                        "tensor_pipe_pct_at_sampled_clock": round(100.0 * ka["tflops"] / pipe_peak_tf, 1),
                        "frac_of_sustained_bf16_peak": round(ka["tflops"] / peak_tf, 3),
                        "ncu_tensor_pipe_active_pct": _ncu_attn_pct(),
                        "ncu_source": "profiles/r02_ncu_attn.tsv (B16 h5 4096x8192, stand-alone, --clock-control none)"}
            hbm_families = {k: {"gb_per_s": v["gb_per_s"], "frac_of_hbm_peak": round(v["gb_per_s"] / peak_gbs, 3),
                                "ms_per_step": round(v["ms"] / args.steps, 3)}
                            for k, v in kernels.items() if k in ops.MEM_KINDS and v.get("gb_per_s")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.operands, "data": "synthetic",
            "config": {"workload": f"{args.nshot}-shot {args.size}x{args.size} episodes, batch {B} per GPU, single-step "
                                   "SD-2.1 UNet (KV-bank attention) + VAE encode x3 / decode + rthres/IoU, random-init "
                                   f"weights (BASELINE config {cfg_id if cfg_id is not None else 'custom'})",
                       "episodes_per_step_per_gpu": B, "parallelism": f"dp{world}",
                       "launch": "one CUDA graph per step" if use_graph else "eager",
                       "l2_policy": "inputs + activations per step (>2 GB) exceed the 126 MB L2; 2 alternating batches",
                       "precision": ("fp32 evaluation mode: every activation fp32, GEMMs / convolutions on the tcgen05 kernels with "
                                     "split f16 operands (hi + lo, 3 partial products), norms / softmax / GELU / attention core in "
                                     "fp32 (latent rel-L2 3e-5 vs the fp32 oracle; tests/test_f32_mode_gpu.py)")
                       if args.operands == "f32" else f"{args.operands} tensor-core operands ("
                                    + ("the reference's own half mode; bf16, the format BASELINE config 2 names, runs at the same "
                                       "rate (--operands bf16) but measures 1.1e-2 .. 1.6e-2 latent rel-L2, above the 1e-2 bar"
                                       if args.operands == "f16" else "selected with --operands bf16; latent rel-L2 1.1e-2 .. 1.6e-2")
                                    + f"), fp32 accumulate / softmax / statistics, UNet residual stream {args.unet_stream}, "
                                    f"VAE stream {args.vae_stream}"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "cpu_baseline": cpu, "kernels": kernels, "attn": attn, "hbm_families": hbm_families,
            "frac_of_tensor_roofline_whole_path": round(value / world * flops_ep / (peak_tf * 1e12), 4) if flops_ep else None,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
