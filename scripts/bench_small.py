"""Stand-alone timings of the small HBM-bound kernels of the step (LayerNorm, cross-attention, row softmax, concat) at the
config-2 shapes: CUDA events, L2 flushed between iterations, median of --iters.  python scripts/bench_small.py"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402

h = torch.float16


def timeit(fn, iters, flush):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=15)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, C in [(65536, 320), (16384, 640), (4096, 1280), (1024, 1280)]:
        x = torch.randn(M, C, device="cuda").to(h)
        g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
        t = timeit(lambda: ops.layernorm(x, g, b, out_dtype=h), a.iters, flush)
        print(f"layernorm M{M} C{C}: {t * 1e3:.1f} us  {2 * M * C * 2 / t * 1e-9:.2f} TB/s", flush=True)
    for M, C, hd in [(65536, 320, 5), (16384, 640, 10), (4096, 1280, 20)]:
        q = torch.randn(16, M // 16, C, device="cuda").to(h)
        k = torch.randn(16, 2, C, device="cuda").to(h); v = torch.randn(16, 2, C, device="cuda").to(h)
        t = timeit(lambda: ops.cross_attn(q, k, v, hd, 0.125), a.iters, flush)
        print(f"cross_attn M{M} C{C}: {t * 1e3:.1f} us  {2 * M * C * 2 / t * 1e-9:.2f} TB/s", flush=True)
    s = torch.randn(65536, 4096, device="cuda")
    t = timeit(lambda: ops.softmax_rows(s, 0.044, out_dtype=h), a.iters, flush)
    print(f"softmax_rows M65536 L4096 (fp32 in, 16-bit out): {t * 1e3:.1f} us  {65536 * 4096 * 6 / t * 1e-9:.2f} TB/s", flush=True)
    del s
    for N, H, Ca, Cb in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280), (16, 64, 320, 640)]:
        xa = torch.randn(N, H, H, Ca, device="cuda").to(h); xb = torch.randn(N, H, H, Cb, device="cuda").to(h)
        t = timeit(lambda: ops.concat_channels(xa, xb), a.iters, flush)
        print(f"concat N{N} {H}x{H} {Ca}+{Cb}: {t * 1e3:.1f} us  {2 * N * H * H * (Ca + Cb) * 2 / t * 1e-9:.2f} TB/s", flush=True)
    for N, HW, C in [(16, 4096, 320), (16, 1024, 640), (16, 256, 1280), (16, 4096, 640), (16, 1024, 1280)]:
        x = torch.randn(N, HW, 1, C, device="cuda").to(h)
        g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
        t = timeit(lambda: ops.groupnorm(x, g, b, eps=1e-5, silu=True, out_dtype=h), a.iters, flush)
        print(f"groupnorm (stats + apply) N{N} HW{HW} C{C}: {t * 1e3:.1f} us  {2 * N * HW * C * 2 / t * 1e-9:.2f} TB/s", flush=True)


if __name__ == "__main__":
    main()
