"""GroupNorm + SiLU folded into the consuming convolution (ops.conv2d_gn_in) against the two-kernel path
(ops.groupnorm -> ops.conv2d) on the VAE shapes of BASELINE config 2: bit-equality and CUDA-event timings.
  python scripts/bench_gnin.py [--iters N] [--out FILE.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402
from diffews_b200.weights import conv_weight_to_gemm  # noqa: E402


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
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    for (N, H, Ci, Co, res) in [(16, 512, 128, 128, False), (16, 512, 128, 128, True), (16, 256, 256, 256, False),
                                (16, 256, 256, 256, True), (16, 128, 512, 512, True), (16, 64, 512, 512, True),
                                (16, 512, 256, 128, False), (16, 256, 512, 256, False)]:
        g = torch.Generator(device="cuda").manual_seed(H + Ci)
        x0 = torch.randn(N, H, H, 64, device="cuda", generator=g).half()
        w0 = conv_weight_to_gemm(torch.randn(Ci, 64, 1, 1, device="cuda", generator=g) * 0.125).half()
        x = ops.conv2d(x0, w0, None, ksize=1, gn_stats=True)            # producer that emits the statistics
        del x0
        assert ops.conv_gn_in_supported(x, Co, 3), (N, H, Ci, Co)
        w = conv_weight_to_gemm(torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) * (9 * Ci) ** -0.5).half()
        b = torch.randn(Co, device="cuda", generator=g)
        gam = torch.randn(Ci, device="cuda", generator=g) + 1.0
        bet = torch.randn(Ci, device="cuda", generator=g)
        r = torch.randn(N, H, H, Co, device="cuda", generator=g).half() if res else None

        def fused():
            return ops.conv2d_gn_in(x, gam, bet, 1e-6, w, b, ksize=3, residual=r, gn_stats=True)

        def split():
            xn = ops.groupnorm(x, gam, bet, eps=1e-6, silu=True, out_dtype=torch.float16)
            return ops.conv2d(xn, w, b, ksize=3, residual=r, gn_stats=True)

        y1, y2 = fused(), split()
        same = bool(torch.equal(y1, y2))
        del y1, y2
        t1, t2 = timeit(fused, a.iters, flush), timeit(split, a.iters, flush)
        fl = 2.0 * N * H * H * Co * Ci * 9
        rows.append({"shape": f"N{N} {H}x{H} {Ci}->{Co} k3" + (" +res" if res else ""), "bit_identical": same,
                     "fused_ms": round(t1, 4), "norm_then_conv_ms": round(t2, 4), "speedup": round(t2 / t1, 3),
                     "fused_tflops_conv_only": round(fl / t1 * 1e-9, 1)})
        print(rows[-1], flush=True)
        del x, w, r
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
