"""Board power, SM clock and energy per launch of the hot kernels run back to back for a few seconds each (NVML sampled
every 20 ms from a thread).  Shows which regime the step is in: python scripts/power_probe.py [--seconds 3]"""
import argparse
import os
import sys
import threading
import time

import pynvml
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402
from diffews_b200.weights import conv_weight_to_gemm  # noqa: E402

h = torch.float16


class Sampler(threading.Thread):
    def __init__(self, handle):
        super().__init__(daemon=True)
        self.h, self.run_flag, self.p, self.c = handle, True, [], []

    def run(self):
        while self.run_flag:
            self.p.append(pynvml.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            self.c.append(pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            time.sleep(0.02)


def probe(name, fn, seconds, handle, work=None, unit=""):
    fn(); torch.cuda.synchronize()
    s = Sampler(handle); s.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    s.run_flag = False; s.join()
    ms = e0.elapsed_time(e1) / n
    k = len(s.p) // 3                                  # skip the ramp
    pw = sum(s.p[k:]) / max(1, len(s.p[k:])); ck = sorted(s.c[k:])[len(s.c[k:]) // 2]
    rate = f"{work / ms * 1e-9:8.1f} {unit}" if work else ""
    print(f"{name:34s} {ms:8.3f} ms  {pw:6.0f} W  {ck:5d} MHz  {pw * ms:8.2f} mJ/launch  {rate}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    a = ap.parse_args()
    pynvml.nvmlInit()
    handle = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    print(f"power limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(handle) / 1000:.0f} W", flush=True)

    def conv(N, H, Ci, Co):
        x = torch.randn(N, H, H, Ci, device="cuda").to(h)
        w = conv_weight_to_gemm(torch.randn(Co, Ci, 3, 3, device="cuda") * (9 * Ci) ** -0.5).to(h)
        b = torch.randn(Co, device="cuda")
        return (lambda: ops.conv2d(x, w, b, ksize=3, gn_stats=True)), 2.0 * N * H * H * Co * Ci * 9
    for name, (N, H, Ci, Co) in [("t128 conv 512^2 128->128", (16, 512, 128, 128)), ("t128 conv 128^2 512->512", (16, 128, 512, 512)),
                                 ("igemm<160> conv 64^2 320->320", (16, 64, 320, 320))]:
        fn, fl = conv(N, H, Ci, Co)
        probe(name, fn, a.seconds, handle, fl, "TFLOP/s")
        del fn
        torch.cuda.empty_cache()
    x = torch.randn(16, 512, 512, 128, device="cuda").to(h)
    w0 = conv_weight_to_gemm(torch.randn(128, 128, 1, 1, device="cuda") * 0.09).to(h)
    y = ops.conv2d(x, w0, None, ksize=1, gn_stats=True)
    g = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
    probe("gn_apply 512^2 x 128 (from partials)", lambda: ops.groupnorm(y, g, b, eps=1e-6, silu=True, out_dtype=h), a.seconds, handle,
          y.numel() * 4, "GB/s")
    del x, y
    torch.cuda.empty_cache()
    B, hd, L = 16, 5, 4096
    C = hd * 64
    qkv = torch.randn(B, L, 3 * C, device="cuda").to(h); bank = torch.randn(B, L, 3 * C, device="cuda").to(h)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    kb, vb = bank[..., C:2 * C], bank[..., 2 * C:]
    probe("attention B16 h5 4096x8192", lambda: ops.attn_kvfused(q, k, v, kb, vb, hd, 0.125), a.seconds, handle,
          4.0 * B * hd * L * 2 * L * 64, "TFLOP/s")
    m = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    probe("torch.matmul bf16 8192^3 (cuBLAS)", lambda: torch.matmul(m, m), a.seconds, handle, 2.0 * 8192 ** 3, "TFLOP/s")


if __name__ == "__main__":
    main()
