"""Stand-alone timing of the KV-fused attention kernel at the shapes of the BASELINE configs (CUDA events, L2 flushed by
rotating over several input sets larger than the 126 MB L2).  Prints TFLOP/s per shape for the v3 (default) and v4 kernels and,
with --v2, the round-1 kernel.  Usage on a B200:  python scripts/bench_attn.py [--v2] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffews_b200 import _lib, ops  # noqa: E402

SHAPES = [  # (name, B, heads, Lq, Ls, Lb)
    ("cfg2 L0 query  B16 h5  4096x8192", 16, 5, 4096, 4096, 4096),
    ("cfg2 L0 support B16 h5 4096x4096", 16, 5, 4096, 4096, 0),
    ("cfg2 L1 query  B16 h10 1024x2048", 16, 10, 1024, 1024, 1024),
    ("cfg2 L2 query  B16 h20 256x512", 16, 20, 256, 256, 256),
    ("cfg2 L3 query  B16 h20 64x128", 16, 20, 64, 64, 64),
    ("cfg3 L0 query  B8 h5 4096x24576", 8, 5, 4096, 4096, 20480),
    ("cfg5 L0 query  B8 h5 9216x18432", 8, 5, 9216, 9216, 9216),
    ("cfg4 L0 query  B1 h5 4096x32768", 1, 5, 4096, 4096, 28672),
]


def time_shape(B, h, Lq, Ls, Lb, iters=20):
    C = h * 64
    nsets = max(2, int(300e6 // (B * (Ls + Lb) * 3 * C * 2)) + 1)
    nsets = min(nsets, 8)
    sets = []
    for i in range(nsets):
        qkv = torch.randn(B, Ls, 3 * C, device="cuda", dtype=torch.float16)
        bank = torch.randn(B, max(Lb, 1), 3 * C, device="cuda", dtype=torch.float16)
        sets.append((qkv, bank))

    def run(i):
        qkv, bank = sets[i % nsets]
        q, k, v = qkv[:, :Lq, :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        kb, vb = (bank[..., C:2 * C], bank[..., 2 * C:]) if Lb else (None, None)
        return ops.attn_kvfused(q, k, v, kb, vb, h, 0.125)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4.0 * B * h * Lq * (Ls + Lb) * 64
    return ms, flops / (ms * 1e9)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--v2", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--lib", default=None, help="alternative build of the library for dfw_attn_kvfused_fwd (A/B of compile-time options)")
    ap.add_argument("--only", type=int, default=None, help="index of the single shape to run")
    args = ap.parse_args()
    if args.lib:
        import ctypes as C
        alt = C.CDLL(os.path.join(ROOT, args.lib))
        fn = alt.dfw_attn_kvfused_fwd
        fn.restype, fn.argtypes = _lib.SIGNATURES["dfw_attn_kvfused_fwd"]
        _lib.lib.dfw_attn_kvfused_fwd = fn
    out = {}
    for ver in (["v4", "v3", "v2"] if args.v2 else ["v4", "v3"]):
        ops.set_option(_lib.OPT_ATTN_V2, int(ver == "v2"))
        ops.set_option(_lib.OPT_ATTN_V4, int(ver == "v4"))
        for name, B, h, Lq, Ls, Lb in (SHAPES if args.only is None else SHAPES[args.only:args.only + 1]):
            ms, tf = time_shape(B, h, Lq, Ls, Lb)
            print(f"{ver}  {name:38s} {ms:8.3f} ms  {tf:7.1f} TFLOP/s  ({100 * tf / 2250:.1f} % of 2250 nominal, "
                  f"{100 * tf / 1397.5:.1f} % of 1397.5 sustained)", flush=True)
            out[f"{ver} {name}"] = {"ms": ms, "tflops": tf}
    ops.set_option(_lib.OPT_ATTN_V2, 0)
    ops.set_option(_lib.OPT_ATTN_V4, 0)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
