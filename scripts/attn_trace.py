"""Cycle-level timeline of the KV-fused attention kernel (debug build, not the product library).

Build (here, no GPU needed):
  cd diffews_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
      -DDFW_BUILD -DDFW_ATTN_TRACE -shared -cudart static -o ../../scripts/microbench/libattn_trace.so common.cu attn.cu
Run on a B200:  python scripts/attn_trace.py
With -DDFW_ATTN_TRACE lane 0 of every warp of CTA (0,0,0) stores clock64() at the phase boundaries of each key tile:
softmax warps  0 before / 1 after the S-ready wait, 2 after the fast pass (exp + vote), 3 after the (rare) exact path,
4 after P is in TMEM, 5 after the p_full arrive; MMA warp  0 / 1 around the P_A-ready wait, 2 / 3 around the P_B-ready wait.
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffews_b200._lib import SIGNATURES  # noqa: E402


def main():
    lib = C.CDLL(os.path.join(ROOT, "scripts", "microbench", sys.argv[1] if len(sys.argv) > 1 else "libattn_trace.so"))
    fn = lib.dfw_attn_kvfused_fwd
    fn.restype, fn.argtypes = SIGNATURES["dfw_attn_kvfused_fwd"]
    lib.dfw_attn_trace_buffer.argtypes = [C.c_void_p]
    B, h, L = 16, 5, 4096
    Cc = h * 64
    qkv = torch.randn(B, L, 3 * Cc, device="cuda").half()
    bank = torch.randn(B, L, 3 * Cc, device="cuda").half()
    o = torch.empty(B, L, Cc, device="cuda", dtype=torch.float16)
    trace = torch.zeros(12 * 64 * 8, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        q, k, v = qkv[..., :Cc], qkv[..., Cc:2 * Cc], qkv[..., 2 * Cc:]
        kb, vb = bank[..., Cc:2 * Cc], bank[..., 2 * Cc:]
        rc = fn(q.data_ptr(), L * 3 * Cc, 3 * Cc, k.data_ptr(), v.data_ptr(), L * 3 * Cc, 3 * Cc, kb.data_ptr(), vb.data_ptr(),
                L * 3 * Cc, 3 * Cc, o.data_ptr(), L * Cc, Cc, B, h, L, L, L, 0.125, 1, st)
        assert rc == 0, rc
    run(); run()
    torch.cuda.synchronize()
    lib.dfw_attn_trace_buffer(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    t = trace.cpu().view(12, 64, 8)
    print("v3 kernel: phase durations in SM cycles, mean over key tiles 8..59 of CTA (0,0,0); one row per softmax warp")
    print("warp tile  wait_S  exp(fast)  slow  P->TMEM  arrive  period")
    for w in range(4, 12):
        a = t[w, 8:60].double()
        nxt = t[w, 9:61, 0].double()
        d = [(a[:, 1] - a[:, 0]).mean(), (a[:, 2] - a[:, 1]).mean(), (a[:, 3] - a[:, 2]).mean(), (a[:, 4] - a[:, 3]).mean(),
             (a[:, 5] - a[:, 4]).mean(), (nxt - a[:, 0]).mean()]
        print(f"{w:4d} {'AB'[(w - 4) // 4]:>4} " + " ".join(f"{float(x):8.0f}" for x in d))
    m = t[1, 8:60].double()
    mn = t[1, 9:61, 0].double()
    print("MMA warp: wait_P_A %.0f  issue(PV_A,S_B) %.0f  wait_P_B %.0f  issue(PV_B,S_A next) %.0f  period %.0f" % (
        float((m[:, 1] - m[:, 0]).mean()), float((m[:, 2] - m[:, 1]).mean()), float((m[:, 3] - m[:, 2]).mean()),
        float((mn - m[:, 3]).mean()), float((mn - m[:, 0]).mean())))
    print("MMA warp detail: kv_full wait %.0f  S_A issue %.0f  [wait P_A]  PV_A issue %.0f  S_B issue %.0f  [wait P_B]  PV_B issue %.0f  loop-back %.0f" % (
        float((m[:, 5] - m[:, 4]).mean()), float((m[:, 0] - m[:, 5]).mean()), float((m[:, 7] - m[:, 1]).mean()),
        float((m[:, 2] - m[:, 7]).mean()), float((m[:, 6] - m[:, 3]).mean()), float((t[1, 9:61, 4].double() - m[:, 6]).mean())))
    off = (t[8, 8:60, 1] - t[4, 8:60, 1]).double().mean()
    print("tile-B softmax starts %.0f cycles after tile-A softmax (same key tile)" % float(off))
    print("tensor-pipe work per key tile pair: 1024 cycles")


if __name__ == "__main__":
    main()
