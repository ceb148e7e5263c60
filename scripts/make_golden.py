"""Generates tests/golden/*.json.  The reference itself cannot be imported here (no diffusers/xformers), so:
  rthres_cases.json          torch-CPU evaluation of the exact expression of evaluation_util/main_oss.py:128-134
  oracle_small_episode.json  the oracle's output on a seeded reduced-width episode (drift guard)
Run from the repo root: python scripts/make_golden.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def rthres_eval(r, g, b, mx):
    img = torch.zeros(1, 3, 1, 2, dtype=torch.uint8)
    img[0, :, 0, 0] = torch.tensor([r, g, b], dtype=torch.uint8)
    img[0, 0, 0, 1] = mx
    pred = img.to(torch.float32).div(255)                  # to_tensor
    thr = pred.max() * 0.25
    return bool((pred.mean(dim=1) > thr)[0, 0, 0])


def main():
    cases = []
    gen = torch.Generator().manual_seed(0)
    # exact ties of the integer rule: 4(R+G+B) == 3*max
    ties_true, ties_false = 0, 0
    for mx in range(4, 256, 4):
        s = 3 * mx // 4
        for _ in range(40):
            r = int(torch.randint(0, min(s, mx) + 1, (1,), generator=gen))
            g = int(torch.randint(0, min(s - r, mx) + 1, (1,), generator=gen))
            b = s - r - g
            if b > mx or b < 0:
                continue
            v = rthres_eval(r, g, b, mx)
            if v and ties_true < 60:
                cases.append([r, g, b, mx, int(v)]); ties_true += 1
            elif not v and ties_false < 60:
                cases.append([r, g, b, mx, int(v)]); ties_false += 1
    for _ in range(200):                                   # random non-tie cases
        mx = int(torch.randint(1, 256, (1,), generator=gen))
        r, g, b = (int(x) for x in torch.randint(0, mx + 1, (3,), generator=gen))
        cases.append([r, g, b, mx, int(rthres_eval(r, g, b, mx))])
    cases.append([0, 0, 0, 0, 0])
    with open(os.path.join(OUT, "rthres_cases.json"), "w") as f:
        json.dump({"source": "torch CPU fp32: to_tensor(u8).mean(dim=1) > to_tensor(u8).max()*0.25 (main_oss.py:128-134)",
                   "ties_true": ties_true, "ties_false": ties_false, "cases": cases}, f)
    print("rthres cases:", len(cases), "ties true/false:", ties_true, ties_false)

    from diffews_b200.synthetic import make_batch, prompt_embedding
    from oracle import pipeline as op
    from oracle import sd21
    torch.set_num_threads(8)
    unet, vae = sd21.build_models(0, (64, 128, 256, 256), (1, 2, 4, 4), (64, 64, 128, 128))
    inter, union, mask, seg_u8, lat = op.evaluate_episode(unet, vae, prompt_embedding(), make_batch(0, 1, 64, 1))
    with open(os.path.join(OUT, "oracle_small_episode.json"), "w") as f:
        json.dump({"latent_abs_mean": float(lat.double().abs().mean()), "seg_u8_mean": float(seg_u8.double().mean()),
                   "inter": inter[:, 0].tolist(), "union": union[:, 0].tolist(), "mask_sum": float(mask.sum())}, f)
    print("oracle small episode:", inter[:, 0].tolist(), union[:, 0].tolist())


if __name__ == "__main__":
    main()
