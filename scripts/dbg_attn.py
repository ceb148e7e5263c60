import sys, os, torch
sys.path.insert(0, os.getcwd())
from diffews_b200 import ops, _lib
import ctypes as C
if len(sys.argv) > 1:
    lib = C.CDLL(os.path.join(os.getcwd(), sys.argv[1]))
    fn = lib.dfw_attn_kvfused_fwd
    fn.restype, fn.argtypes = _lib.SIGNATURES["dfw_attn_kvfused_fwd"]
    _lib.lib.dfw_attn_kvfused_fwd = fn
    ops.lib = _lib.lib
    print("using", sys.argv[1])
g = torch.Generator().manual_seed(5)
B, h, L = 1, 3, 512
C = h * 64
for scale in (1.0, 2.0, 3.0):
    g = torch.Generator().manual_seed(5)
    q = (torch.randn(B, L, C, generator=g) * scale).half()
    k = (torch.randn(B, L, C, generator=g) * scale).half()
    v = torch.randn(B, L, C, generator=g).half()
    kb = (torch.randn(B, 2 * L, C, generator=g) * scale).half()
    vb = torch.randn(B, 2 * L, C, generator=g).half()
    o = ops.attn_kvfused(q.cuda(), k.cuda(), v.cuda(), kb.cuda(), vb.cuda(), h, 0.125).float().cpu()
    hd = lambda t: t.float().view(B, -1, h, 64).transpose(1, 2)
    K = torch.cat([k, kb], 1); V = torch.cat([v, vb], 1)
    S = hd(q) @ hd(K).transpose(-1, -2) * 0.125 * 1.4426950408889634
    ref = (torch.softmax(hd(q) @ hd(K).transpose(-1, -2) * 0.125, -1) @ hd(V)).transpose(1, 2).reshape(B, L, C)
    bad = torch.isnan(o).view(B, L, h, 64).any(-1)[0]      # [L, h]
    print("scale", scale, "nan rows per head", bad.sum(0).tolist(), "rel", ((o - ref).norm() / ref.norm()).item())
    if bad.any():
        idx = bad.nonzero()[:6]
        for r, hh in idx.tolist():
            s = S[0, hh, r]                       # [Lk] log2-domain logits
            tiles = s.view(-1, 128)
            tmax = tiles.max(-1).values
            run = torch.cummax(tmax, 0).values
            print(" row", r, "head", hh, "tile maxima", [round(float(x), 1) for x in tmax], "min logit", float(s.min()))
    if scale == 3.0:
        oh = o.view(B, L, h, 64)[0]; rh = ref.view(B, L, h, 64)[0]
        err = ((oh - rh).norm(dim=-1) / rh.norm(dim=-1))        # [L, h]
        err = torch.nan_to_num(err, nan=9.9)
        flat = err.flatten().topk(8)
        for e, i in zip(flat.values.tolist(), flat.indices.tolist()):
            r, hh = divmod(i, h)
            s = S[0, hh, r]; tmax = s.view(-1, 128).max(-1).values
            # reconstruct: what if tiles before the first trigger were dropped / mis-scaled?
            pr = torch.softmax(s * 0.6931471805599453, -1)
            contrib = pr.view(-1, 128).sum(-1)
            print(f" row {r} (warp q{(r % 128) // 32} tile {'AB'[(r % 256) // 128]}) head {hh}: err {e:.3f}; tile maxima {[round(float(x), 1) for x in tmax]}; tile prob mass {[round(float(x), 3) for x in contrib]}")
        print(" rows with err > 1e-2:", int((err > 1e-2).sum()), "of", err.numel(), "; by warp quadrant:", [(int(((err > 1e-2).any(-1))[q * 32:(q + 1) * 32].sum())) for q in range(16)])
        # hypotheses for the worst rows
        Vh = hd(V)[0]                                          # [h, Lk, 64]
        for e, i in zip(flat.values.tolist()[:4], flat.indices.tolist()[:4]):
            r, hh = divmod(i, h)
            s = S[0, hh, r]
            w = torch.exp2(s - s.max())
            nt = s.numel() // 128
            full = (w[:, None] * Vh[hh]).sum(0) / w.sum()
            for drop in range(nt - 3, nt):
                w2 = w.clone(); w2[drop * 128:(drop + 1) * 128] = 0
                alt = (w2[:, None] * Vh[hh]).sum(0) / w2.sum()
                print(f"   row {r} head {hh}: |o - ref_without_tile{drop}| / |ref| = {float((oh[r, hh] - alt).norm() / full.norm()):.3f}", end=";")
            # numerator without last tile but denominator with it, and vice versa
            w2 = w.clone(); w2[(nt - 1) * 128:] = 0
            a1 = (w2[:, None] * Vh[hh]).sum(0) / w.sum(); a2 = (w[:, None] * Vh[hh]).sum(0) / w2.sum()
            print(f" num-only-drop {float((oh[r, hh] - a1).norm() / full.norm()):.3f} den-only-drop {float((oh[r, hh] - a2).norm() / full.norm()):.3f} |o|/|ref| {float(oh[r, hh].norm() / full.norm()):.3f}")
