"""Measurement of the fused AdamW step (+ gradient-norm clip) at SD-2.1 UNet size: 866 M fp32 parameters in ~690 tensors.
Algorithmic bytes per parameter: clip 4 (grad read) ; step 16 read (p, g, m, v) + 12 written (p, m, v) + 2 (fp16 copy).
    python scripts/bench_optim.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from diffews_b200.optim import AdamW
    # SD-2.1-like size mix: conv 3x3 weights, linears, GEGLU projections, norms / biases
    shapes = []
    for c in (320, 640, 1280):
        reps = {320: 12, 640: 14, 1280: 30}[c]
        for _ in range(reps):
            shapes += [(c, c, 3, 3), (c,), (c,), (c,)]
        for _ in range({320: 5, 640: 5, 1280: 6}[c]):
            shapes += [(c, c)] * 4 + [(c, 1024)] * 2 + [(c, c)] * 2 + [(8 * c, c), (8 * c,), (c, 4 * c), (c,)] + [(c,)] * 6
    ps = [torch.randn(s, device="cuda") * 0.02 for s in shapes]
    n = sum(p.numel() for p in ps)
    for p in ps:
        p.grad = torch.randn_like(p) * 1e-3
    half = [torch.empty(p.numel(), dtype=torch.float16, device="cuda") for p in ps]
    opt = AdamW(ps, lr=1e-5, half_copies=half)

    def timed(fn, reps=5):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_clip = timed(lambda: opt.clip_grad_norm_(1.0))
    ms_step = timed(opt.step)
    peak = None
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    gbs_clip = n * 4 / ms_clip / 1e6
    gbs_step = n * 30 / ms_step / 1e6
    # torch's own fused optimizer beside it (library baseline, same tensors)
    ref = torch.optim.AdamW(ps, lr=1e-5, fused=True)
    ms_torch = timed(ref.step)
    # LayerNorm backward at the config-2/4 token counts (16 x 4096 tokens of 320 channels; 16 x 256 of 1280), fp16
    from diffews_b200 import ops
    ln = {}
    for M, C in ((65536, 320), (16384, 640), (4096, 1280)):
        x = torch.randn(M, C, device="cuda").half(); dy = torch.randn(M, C, device="cuda").half()
        gam = torch.ones(C, device="cuda")
        ms = timed(lambda: ops.layernorm_backward(x, dy, gam, 1e-5), reps=20)
        ln[f"M{M}_C{C}"] = {"ms": round(ms, 4), "gb_per_s": round(M * C * 6 / ms / 1e6, 1)}
    print(json.dumps({"layernorm_bwd_fp16 (6 B / element: x, dy read, dx written)": ln,
                      "what": "fused multi-tensor AdamW (+fp16 operand copy) and gradient-norm clip", "tensors": len(ps),
                      "parameters": n, "working_set_gb": round(n * 18 / 1e9, 2), "launches": {"clip": 2, "step": 1},
                      "ms_clip": round(ms_clip, 3), "ms_step": round(ms_step, 3),
                      "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "clip_achieved": round(gbs_clip, 1),
                                   "step_achieved": round(gbs_step, 1),
                                   "step_frac": round(gbs_step / peak, 4) if peak else None,
                                   "algorithmic_bytes_per_param": {"clip": 4, "step": 30}},
                      "torch_fused_adamw_ms_step": round(ms_torch, 3)}))


if __name__ == "__main__":
    main()
