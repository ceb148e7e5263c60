"""Token GEMMs of the step (x [M, K] @ w[N, K]^T) with the weight-stationary mainloop on / off (DFW_OPT_B_RESIDENT):
bit-equality of the two and CUDA-event timings, L2 flushed between iterations.  python scripts/bench_linear.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import _lib, ops  # noqa: E402

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


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, K, N, res) in [(65536, 320, 320, False), (65536, 320, 320, True), (65536, 320, 960, False), (65536, 512, 512, False),
                       (65536, 512, 512, True), (16384, 640, 640, False), (28672, 320, 320, True), (539, 1024, 320, False)]:
    x = torch.randn(M, K, device="cuda").to(h)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(h)
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").to(h) if res else None
    out = {}
    for mode in (1, 0):
        ops.set_option(_lib.OPT_B_RESIDENT, mode)
        y = ops.linear(x, w, b, residual=r)
        t = timeit(lambda: ops.linear(x, w, b, residual=r), 11, flush)
        out[mode] = (y, t)
    ops.set_option(_lib.OPT_B_RESIDENT, 0)
    same = bool(torch.equal(out[0][0], out[1][0]))
    fl = 2.0 * M * K * N
    print(f"linear M{M} K{K} N{N} res={res}: stationary {out[1][1] * 1e3:.1f} us ({fl / out[1][1] * 1e-9:.0f} TFLOP/s)  "
          f"ring {out[0][1] * 1e3:.1f} us ({fl / out[0][1] * 1e-9:.0f} TFLOP/s)  bit-identical {same}", flush=True)
