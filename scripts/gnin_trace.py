"""Phase cycle counts of the GN_IN convolution (CTA 0), from a -DDFW_GNIN_TRACE=1 build of the library
(scripts/build_variant.sh igemm.cu X.so -DDFW_GNIN_TRACE=1; copy X.so over diffews_b200/libdiffews_b200.so on the box)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffews_b200 import ops, _lib  # noqa: E402
from diffews_b200.weights import conv_weight_to_gemm  # noqa: E402

lib = C.CDLL(_lib.LIB_PATH)
lib.dfw_debug_t128_trace.argtypes = [C.POINTER(C.c_longlong), C.c_int]
NAMES = ["mma_total", "mma_wait_patch", "mma_wait_w", "mma_wait_acc", "xf_total", "xf_load+math", "xf_wait_slot", "xf_write", "xf_chunks", "xf_fence"]

for (N, H, Ci, Co, res) in [(16, 512, 128, 128, False), (16, 256, 256, 256, True), (16, 128, 512, 512, True)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.randn(N, H, H, 64, device="cuda", generator=g).half()
    w0 = conv_weight_to_gemm(torch.randn(Ci, 64, 1, 1, device="cuda", generator=g) * 0.125).half()
    x = ops.conv2d(x0, w0, None, ksize=1, gn_stats=True)
    w = conv_weight_to_gemm(torch.randn(Co, Ci, 3, 3, device="cuda", generator=g) * (9 * Ci) ** -0.5).half()
    gam = torch.ones(Ci, device="cuda"); bet = torch.zeros(Ci, device="cuda")
    r = torch.randn(N, H, H, Co, device="cuda", generator=g).half() if res else None
    out = (C.c_longlong * 16)()
    for it in range(3):
        lib.dfw_debug_t128_trace(out, 1)
        ops.conv2d_gn_in(x, gam, bet, 1e-6, w, None, ksize=3, residual=r, gn_stats=True)
        lib.dfw_debug_t128_trace(out, 0)
    n = max(1, out[8])
    print(f"N{N} {H}x{H} {Ci}->{Co} res={res}: chunks {out[8]}; per chunk cycles: " +
          ", ".join(f"{NAMES[i]} {out[i] / n:.0f}" for i in (0,1,2,3,4,5,6,7,9)), flush=True)
