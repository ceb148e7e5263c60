"""Stand-alone timing of the fused decoder head (dfw_seg_head_u8) against the three-launch path it replaces, at the shapes
of BASELINE configs 2 and 5.  CUDA events, inputs rotated over > 126 MB so nothing stays in L2.  Usage: python scripts/bench_seg_head.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffews_b200 import ops  # noqa: E402
from diffews_b200.weights import conv_weight_to_gemm  # noqa: E402


def main():
    out = {}
    g = torch.Generator().manual_seed(0)
    gam = (torch.randn(128, generator=g) * 0.3 + 1.0).cuda(); bet = (torch.randn(128, generator=g) * 0.2).cuda()
    w = torch.randn(3, 128, 3, 3, generator=g) * 0.03
    b = torch.randn(3, generator=g) * 0.1
    wb = ops.seg_head_prepare(w, torch.float16, "cuda")
    wg = conv_weight_to_gemm(w).cuda().half()
    for name, N, S in (("cfg2 B16 512x512", 16, 512), ("cfg5 B8 768x768", 8, 768)):
        xs = []
        for i in range(2):
            x0 = torch.randn(N, S, S, 64, device="cuda", dtype=torch.float16)
            w0 = (torch.randn(128, 64, device="cuda") * 0.2).half()
            x = ops.conv2d(x0, w0, None, ksize=1, gn_stats=True)
            assert ops.seg_head_supported(x)
            xs.append(x)
            del x0
        algo = N * S * S * (128 * 2 + 3)

        def fused(i):
            return ops.seg_head_u8(xs[i % 2], gam, bet, 1e-6, wb, b, want_f32=False, want_u8=True)

        def unfused(i):
            x = xs[i % 2]
            h = ops.groupnorm(x, gam, bet, eps=1e-6, silu=True, out_dtype=torch.float16)
            y = ops.conv2d(h, wg, b.cuda(), ksize=3, out_f32=True)
            return ops.seg_post(y.view(N, S * S, 3), S, S, want_f32=False, want_u8=True)
        for label, fn in (("fused dfw_seg_head_u8", fused), ("gn-apply + igemm 128->3 + seg_post", unfused)):
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(10):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{name}  {label:36s} {ms:7.3f} ms  {algo / ms / 1e9:6.2f} TB/s of algorithmic bytes", flush=True)
            out[f"{name} {label}"] = {"ms": ms, "tb_per_s_algorithmic": algo / ms / 1e9}
    with open(os.path.join(ROOT, "gpurun_out", "r02_seg_head_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
