"""BASELINE config 4: training-shape forward + backward of the KV-fused attention and the GroupNorm(+SiLU) kernels
(7 supports, query batch 1).  CUDA-event timings, median of 5, > L2-size flush between iterations.
  python scripts/bench_config4.py > profiles/r02_config4_fwd_bwd.tsv"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402


def timeit(fn, flush, iters=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dt = torch.float16
    print("kernel\tshape\tfwd_ms\tbwd_ms\tfwd_rate\tbwd_rate")
    for (h, Lq) in [(5, 4096), (10, 1024), (20, 256), (20, 64)]:
        Lb = 7 * Lq
        C = h * 64
        mk = lambda L: torch.randn(1, L, C, device="cuda").to(dt)
        q, ks, vs, kb, vb, d_o = mk(Lq), mk(Lq), mk(Lq), mk(Lb), mk(Lb), mk(Lq)
        o, lse = ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125, return_lse=True)
        f = timeit(lambda: ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125, return_lse=True), flush)
        b = timeit(lambda: ops.attn_kvfused_backward(q, ks, vs, kb, vb, o, d_o, h, 0.125, lse=lse), flush)
        fl = 4.0 * h * Lq * (Lq + Lb) * 64
        print(f"attn_kvfused\tB1 h{h} Lq{Lq} Lk{Lq + Lb}\t{f:.3f}\t{b:.3f}\t{fl / f / 1e9:.1f} TFLOP/s\t{2.5 * fl / b / 1e9:.1f} TFLOP/s (2.5x fwd FLOPs)")
        del q, ks, vs, kb, vb, d_o, o
        torch.cuda.empty_cache()
    for (N, H, C) in [(7, 64, 320), (7, 32, 640), (7, 16, 1280), (7, 8, 1280), (1, 64, 320), (1, 64, 960), (1, 32, 1920), (1, 16, 2560)]:
        x = torch.randn(N, H, H, C, device="cuda").to(dt); dy = torch.randn(N, H, H, C, device="cuda").to(dt)
        g = torch.ones(C, device="cuda"); be = torch.zeros(C, device="cuda")
        f = timeit(lambda: ops.groupnorm(x, g, be, eps=1e-5, silu=True, out_dtype=dt), flush)
        b = timeit(lambda: ops.groupnorm_backward(x, dy, g, be, eps=1e-5, silu=True), flush)
        nb = x.numel() * 2
        print(f"groupnorm_silu\tN{N} {H}x{H} C{C}\t{f:.4f}\t{b:.4f}\t{2 * nb / f / 1e6:.0f} GB/s (in+out)\t{3 * nb / b / 1e6:.0f} GB/s (x+dy+dx)")


if __name__ == "__main__":
    main()
