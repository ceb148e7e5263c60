"""Standalone launches of the hot kernels at BASELINE config-2 shapes (B = 16, 512^2), timed with CUDA events.
Used (a) for quick kernel iteration and (b) as the short command profiled under `ncu --set full`.
  python scripts/profile_kernels.py [conv128|conv256|conv512|conv320|conv1280|lin|geglu|attn|gn|all] [--iters N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops  # noqa: E402
from diffews_b200.weights import conv_weight_to_gemm  # noqa: E402

bf16 = torch.float16          # the product default 16-bit format (name kept for brevity)


def timeit(fn, iters, flush):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # > L2-size write between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def conv_case(N, H, C_in, C_out, ks=3, out_f32=False, res=None, gn=False):
    x = torch.randn(N, H, H, C_in, device="cuda").to(bf16)
    w = conv_weight_to_gemm(torch.randn(C_out, C_in, ks, ks, device="cuda") * (C_in * ks * ks) ** -0.5).to(bf16)
    b = torch.randn(C_out, device="cuda")
    r = None
    if res is not None:
        r = torch.randn(N, H, H, C_out, device="cuda").to(res)
    flops = 2.0 * N * H * H * C_out * C_in * ks * ks
    return (lambda: ops.conv2d(x, w, b, ksize=ks, out_f32=out_f32, residual=r, gn_stats=gn)), flops, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["all"])
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = {
        "conv128": lambda: conv_case(16, 512, 128, 128),
        "conv128_f32res": lambda: conv_case(16, 512, 128, 128, out_f32=True, res=torch.float32),
        "conv128_h16res": lambda: conv_case(16, 512, 128, 128, res=bf16),
        "conv256_f32res": lambda: conv_case(16, 256, 256, 256, out_f32=True, res=torch.float32),
        "conv256": lambda: conv_case(16, 256, 256, 256),
        "conv512": lambda: conv_case(16, 128, 512, 512),
        "conv512s": lambda: conv_case(16, 64, 512, 512),
        "conv320": lambda: conv_case(16, 64, 320, 320),
        "conv640": lambda: conv_case(16, 32, 640, 640),
        "conv1280": lambda: conv_case(16, 16, 1280, 1280),
        "conv1280s": lambda: conv_case(16, 8, 1280, 1280),
        "conv128to3": lambda: conv_case(16, 512, 128, 3),
        "conv64k1": lambda: conv_case(16, 512, 64, 128, ks=1),
        "conv64k1_gn": lambda: conv_case(16, 512, 64, 128, ks=1, gn=True),
        "conv128_gn": lambda: conv_case(16, 512, 128, 128, gn=True),
    }

    def lin_res():
        x = torch.randn(16 * 4096, 320, device="cuda").to(bf16); w = (torch.randn(320, 320, device="cuda") * 0.05).to(bf16)
        r = torch.randn(16 * 4096, 320, device="cuda"); b = torch.randn(320, device="cuda")
        return (lambda: ops.linear(x, w, b, residual=r, out_f32=True)), 2.0 * 16 * 4096 * 320 * 320, None

    def lin():
        x = torch.randn(16 * 4096, 320, device="cuda").to(bf16); w = (torch.randn(960, 320, device="cuda") * 0.05).to(bf16)
        return (lambda: ops.linear(x, w)), 2.0 * 16 * 4096 * 320 * 960, None

    def geglu():
        x = torch.randn(16 * 4096, 320, device="cuda").to(bf16); w = (torch.randn(2560, 320, device="cuda") * 0.05).to(bf16)
        b = torch.randn(2560, device="cuda")
        return (lambda: ops.linear(x, w, b, geglu=True)), 2.0 * 16 * 4096 * 320 * 2560, None

    def attn():
        B, h, L = 16, 5, 4096
        qkv = torch.randn(B, L, 3 * 320, device="cuda").to(bf16)
        bank = torch.randn(B, L, 3 * 320, device="cuda").to(bf16)
        q, k, v = qkv[..., :320], qkv[..., 320:640], qkv[..., 640:]
        kb, vb = bank[..., 320:640], bank[..., 640:]
        return (lambda: ops.attn_kvfused(q, k, v, kb, vb, h, 0.125)), 4.0 * B * h * L * 2 * L * 64, None

    def attn_bwd():            # BASELINE config 4, 64x64 level: query batch 1 against its own + 7 supports' keys
        h, L = 5, 4096
        mk = lambda n: torch.randn(1, n, h * 64, device="cuda").to(bf16)
        q, ks, vs, kb, vb, d_o = mk(L), mk(L), mk(L), mk(7 * L), mk(7 * L), mk(L)
        o, lse = ops.attn_kvfused(q, ks, vs, kb, vb, h, 0.125, return_lse=True)
        return (lambda: ops.attn_kvfused_backward(q, ks, vs, kb, vb, o, d_o, h, 0.125, lse=lse)), 2.5 * 4.0 * h * L * 8 * L * 64, None

    def gn():
        x = torch.randn(16, 512 * 512, 128, device="cuda")
        g = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
        nbytes = x.numel() * (4 + 2)          # algorithmic: read fp32 once, write bf16 once
        return (lambda: ops.groupnorm(x, g, b, eps=1e-6, silu=True)), None, nbytes
    def gn16():
        x = torch.randn(16, 512 * 512, 128, device="cuda").half()
        g = torch.ones(128, device="cuda"); b = torch.zeros(128, device="cuda")
        nbytes = x.numel() * (2 + 2)          # algorithmic: read fp16 once, write fp16 once
        return (lambda: ops.groupnorm(x, g, b, eps=1e-6, silu=True, out_dtype=torch.float16)), None, nbytes
    def im2col():
        x = torch.randn(16, 3, 512, 512, device="cuda")
        nbytes = x.numel() * 4 + 16 * 512 * 512 * 64 * 2
        return (lambda: ops.im2col3x3_small(x, 64, torch.float16)), None, nbytes
    cases.update(im2col=im2col)

    def gnapply_small():
        x0 = torch.randn(16, 64, 64, 512, device="cuda").half()
        w = conv_weight_to_gemm(torch.randn(512, 512, 3, 3, device="cuda") * 0.01).half()
        y = ops.conv2d(x0, w, torch.zeros(512, device="cuda"), ksize=3, gn_stats=True)
        g = torch.ones(512, device="cuda"); b = torch.zeros(512, device="cuda")
        nbytes = y.numel() * 4
        return (lambda: ops.groupnorm(y, g, b, eps=1e-6, silu=True, out_dtype=torch.float16)), None, nbytes
    cases.update(gnapply_small=gnapply_small)

    def xattn(M=65536, C=320, h=5):
        lg = torch.randn(16, M // 16, 16 * ((2 * h + 15) // 16), device="cuda")
        U = torch.randn(2 * h, C, device="cuda"); b = torch.randn(C, device="cuda")
        r = torch.randn(16, M // 16, C, device="cuda").half()
        nbytes = lg.numel() * 4 + 2 * r.numel() * 2
        return (lambda: ops.cross_attn_collapsed(lg, U, b, r, h, 2, torch.float16)), None, nbytes
    cases.update(xattn=xattn, xattn1280=lambda: xattn(4096, 1280, 20))
    cases.update(lin=lin, lin_res=lin_res, geglu=geglu, attn=attn, attn_bwd=attn_bwd, gn=gn, gn16=gn16)
    which = list(cases) if args.which == ["all"] else args.which
    for name in which:
        fn, flops, nbytes = cases[name]()
        ms = timeit(fn, args.iters, flush)
        if flops:
            print(f"{name:10s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s")
        else:
            print(f"{name:10s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic bytes)")
        del fn
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
