"""GPU parity of the training-step tail (first pieces) against torch's own CPU implementation of the
calls the reference makes (train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1384, :1393, :1186-1194 + :1394):
F.mse_loss, torch.nn.utils.clip_grad_norm_, torch.optim.AdamW.  Floating point: the tolerance is 2e-6 relative on the
parameters after several steps (same formula, fp32, different summation / FMA contraction)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(320, 320, 3, 3), (1280,), (5, 7), (1,), (640, 2560), (65536 * 2 + 3,)]


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) * 0.05 for s in SHAPES]


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("max_norm", [None, 0.5, 1e6])
def test_adamw_and_clip_match_torch(lib_built, max_norm):
    from diffews_b200.optim import AdamW
    ref_p = [torch.nn.Parameter(p.clone()) for p in _params(0)]
    dev_p = [p.detach().clone().cuda() for p in ref_p]
    half = [torch.empty(p.numel(), dtype=torch.float16, device="cuda") for p in dev_p]
    kw = dict(lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    ref_opt = torch.optim.AdamW(ref_p, foreach=False, **kw)
    opt = AdamW(dev_p, half_copies=half, **kw)
    for it in range(4):
        grads = [g * (3.0 if it == 1 else 1.0) for g in _params(100 + it)]
        for p, g in zip(ref_p, grads):
            p.grad = g.clone()
        for p, g in zip(dev_p, grads):
            p.grad = g.clone().cuda()
        if max_norm is not None:
            # torch.nn.utils.clip_grad_norm_ formula.  torch's CPU float32 accumulation of the per-tensor norms is itself
            # ~1.5e-5 off on tensors of this size (measured against float64), so the norm is checked against the float64
            # value (tight) and against torch (loose), and the reference gradients are clipped with the accurately
            # evaluated coefficient  min(1, max_norm / (norm + 1e-6)).
            exact = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
            n_dev = opt.clip_grad_norm_(max_norm)
            assert abs(n_dev.item() - exact.item()) <= 1e-6 * exact.item()
            torch_n = torch.nn.utils.clip_grad_norm_(ref_p, max_norm=float("inf"))       # norm only, no scaling
            assert abs(n_dev.item() - torch_n.item()) <= 1e-4 * exact.item()
            coef = min(1.0, max_norm / (exact.float().item() + 1e-6))
            for p in ref_p:
                p.grad.mul_(torch.tensor(coef, dtype=torch.float32))
        ref_opt.step()
        opt.step()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(dev_p, ref_p)):
        # norm-wise relative error; the 1- and 35-element tensors get a looser bar (a single m + w (g - m) with
        # cancellation is a few ulp off under a different FMA contraction, which a norm over 10^5 elements averages out)
        tol = 2e-6 if a.numel() >= 1000 else 5e-5
        assert rel(a.cpu(), b.detach()) < tol, (i, rel(a.cpu(), b.detach()))
        st = ref_opt.state[b]
        assert rel(opt.state[i]["exp_avg"].cpu(), st["exp_avg"]) < tol, (i, "exp_avg")
        assert rel(opt.state[i]["exp_avg_sq"].cpu(), st["exp_avg_sq"]) < tol, (i, "exp_avg_sq")
        assert torch.equal(half[i].cpu(), a.cpu().reshape(-1).half())          # the fused 16-bit copy is exactly the rounding
    # the gradients were not modified by the clip (the coefficient is applied inside the step)
    assert torch.equal(dev_p[0].grad.cpu(), _params(103)[0])


def test_adamw_state_dict_round_trips_with_torch(lib_built):
    """state_dict() has torch.optim.AdamW's layout (the reference checkpoints the optimizer through accelerator.save_state,
    train...v3.py:1408-1414): torch can load ours, we can load torch's, and both continue identically."""
    from diffews_b200.optim import AdamW
    kw = dict(lr=2e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=5e-2)
    ref_p = [torch.nn.Parameter(p.clone()) for p in _params(3)]
    ref_opt = torch.optim.AdamW(ref_p, foreach=False, **kw)
    for it in range(2):
        for p, g in zip(ref_p, _params(300 + it)):
            p.grad = g.clone()
        ref_opt.step()
    dev_p = [p.detach().clone().cuda() for p in ref_p]
    opt = AdamW(dev_p, lr=1.0, betas=(0.5, 0.5), eps=1.0, weight_decay=0.0)           # overwritten by the loaded group
    opt.load_state_dict(ref_opt.state_dict())
    assert (opt.lr, opt.betas, opt.eps, opt.weight_decay) == (2e-3, (0.9, 0.99), 1e-8, 5e-2)
    assert opt.state[0]["step"] == 2
    grads = _params(400)
    for p, g in zip(ref_p, grads):
        p.grad = g.clone()
    for p, g in zip(dev_p, grads):
        p.grad = g.clone().cuda()
    ref_opt.step(); opt.step()
    for a, b in zip(dev_p, ref_p):
        assert rel(a.cpu(), b.detach()) < (2e-6 if a.numel() >= 1000 else 5e-5)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["params"] == list(range(len(dev_p)))
    fresh_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    fresh = torch.optim.AdamW(fresh_p, foreach=False, **kw)
    fresh.load_state_dict({"state": {k: {n: (v.cpu() if torch.is_tensor(v) else v) for n, v in st.items()}
                                     for k, st in sd["state"].items()}, "param_groups": sd["param_groups"]})
    assert float(fresh.state[fresh_p[0]]["step"]) == 3.0
    assert rel(fresh.state[fresh_p[0]]["exp_avg"], ref_opt.state[ref_p[0]]["exp_avg"]) < 2e-6


def test_adamw_found_inf_guard_and_stale_clip(lib_built):
    """The reference trains under an fp16 GradScaler: one inf gradient must not poison parameters or moments.  And the
    clip coefficient is only valid for the gradient values it was computed from."""
    from diffews_b200.optim import AdamW
    dev_p = [p.clone().cuda() for p in _params(5)]
    before = [p.clone() for p in dev_p]
    opt = AdamW(dev_p, lr=1e-2)
    for p, g in zip(dev_p, _params(500)):
        p.grad = g.clone().cuda()
    dev_p[2].grad[0, 0] = float("inf")
    n = opt.clip_grad_norm_(1.0)
    opt.step()                                         # device-side guard
    torch.cuda.synchronize()
    assert not torch.isfinite(n).item()
    for a, b in zip(dev_p, before):
        assert torch.equal(a, b)
    assert all(float(st["exp_avg"].abs().max()) == 0.0 for st in opt.state.values())
    opt.clip_grad_norm_(1.0)
    opt.step(skip_nonfinite=True)                      # host-visible skip: the step counter does not advance either
    assert opt.state[0]["step"] == 1
    dev_p[2].grad[0, 0] = 0.5
    opt.clip_grad_norm_(1.0)
    dev_p[0].grad.add_(1.0)                            # a gradient written after the clip
    with pytest.raises(RuntimeError, match="stale"):
        opt.step()
    opt.clip_grad_norm_(1.0)
    opt.step()
    torch.cuda.synchronize()
    assert not torch.equal(dev_p[0], before[0]) and all(torch.isfinite(p).all() for p in dev_p)


def test_adamw_is_deterministic(lib_built):
    from diffews_b200.optim import AdamW
    outs = []
    for _ in range(2):
        ps = [p.cuda() for p in _params(1)]
        opt = AdamW(ps, lr=1e-3)
        for it in range(2):
            for p, g in zip(ps, _params(50 + it)):
                p.grad = g.cuda()
            n = opt.clip_grad_norm_(1.0)
            opt.step()
        outs.append([p.cpu().clone() for p in ps] + [n.cpu().clone()])
    assert all(torch.equal(a, b) for a, b in zip(*outs))


@pytest.mark.parametrize("shape", [(1, 4, 64, 64), (7, 4, 96, 96), (3,), (2, 4, 17, 5)])
def test_mse_loss_matches_torch(lib_built, shape):
    from diffews_b200.optim import mse_loss
    g = torch.Generator().manual_seed(3)
    pred = torch.randn(shape, generator=g).requires_grad_(True)
    target = torch.randn(shape, generator=g)
    ref = torch.nn.functional.mse_loss(pred.float(), target.float(), reduction="mean")
    ref.backward()
    loss, dpred = mse_loss(pred.detach().cuda(), target.cuda())
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    assert rel(dpred.cpu(), pred.grad) < 1e-6


def test_optimizer_refuses_cpu_parameters():
    from diffews_b200.optim import AdamW
    with pytest.raises(TypeError):
        AdamW([torch.zeros(4)])


@pytest.mark.parametrize("M,C,dt", [(4096, 320, torch.float16), (1024, 640, torch.float32), (77, 1280, torch.float16),
                                    (3, 320, torch.bfloat16), (5000, 1280, torch.float32)])
def test_layernorm_backward_matches_autograd(lib_built, M, C, dt):
    """Backward of BasicTransformerBlock's LayerNorms (eps 1e-5) against torch autograd in fp32 on the same (rounded)
    inputs; run twice: bit-identical (fixed reduction order)."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    x = (torch.randn(M, C, generator=g) * 1.7 + 0.3).to(dt)
    dy = torch.randn(M, C, generator=g).to(dt)
    gamma = torch.randn(C, generator=g) * 0.5 + 1.0
    beta = torch.randn(C, generator=g) * 0.1
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy.float())
    outs = [ops.layernorm_backward(x.cuda(), dy.cuda(), gamma.cuda(), 1e-5) for _ in range(2)]
    torch.cuda.synchronize()
    dx, dg, db = outs[0]
    assert all(torch.equal(a, b) for a, b in zip(outs[0], outs[1]))
    tol = 1e-5 if dt == torch.float32 else (6e-3 if dt == torch.bfloat16 else 8e-4)
    assert dx.dtype == dt and rel(dx.cpu(), xr.grad) < tol, rel(dx.cpu(), xr.grad)
    assert rel(dg.cpu(), gr.grad) < 2e-5 and rel(db.cpu(), br.grad) < 2e-5


@pytest.mark.parametrize("M,F,dt", [(4096, 1280, torch.float16), (257, 2560, torch.float32), (64, 5120, torch.bfloat16)])
def test_geglu_backward_matches_autograd(lib_built, M, F, dt):
    """diffusers GEGLU (FeedForward of BasicTransformerBlock): y = v * gelu_erf(g) with (v | g) = h.chunk(2, -1)."""
    from diffews_b200 import ops
    g = torch.Generator().manual_seed(F)
    h = (torch.randn(M, 2 * F, generator=g) * 1.5).to(dt)
    dy = torch.randn(M, F, generator=g).to(dt)
    hr = h.float().requires_grad_(True)
    v, gate = hr.chunk(2, dim=-1)
    (v * torch.nn.functional.gelu(gate)).backward(dy.float())
    dh = ops.geglu_backward(h.cuda(), dy.cuda())
    tol = 1e-5 if dt == torch.float32 else (6e-3 if dt == torch.bfloat16 else 8e-4)
    assert dh.dtype == dt and rel(dh.cpu(), hr.grad) < tol, rel(dh.cpu(), hr.grad)


def test_linear_backward_matches_autograd(lib_built):
    """dx / dW / dbias of an nn.Linear on the tcgen05 GEMM (first correct path, diffews_b200/backward.py)."""
    from diffews_b200.backward import linear_backward
    g = torch.Generator().manual_seed(11)
    M, K, N = 200, 320, 640                                  # M not a multiple of 64: rows are zero-padded
    x = torch.randn(M, K, generator=g).half()
    w = (torch.randn(N, K, generator=g) * K ** -0.5).half()
    dy = torch.randn(M, N, generator=g).half()
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    b = torch.zeros(N, requires_grad=True)
    torch.nn.functional.linear(xr, wr, b).backward(dy.float())
    dx, dw, db = linear_backward(x.cuda(), w.cuda(), dy.cuda())
    assert rel(dx.cpu(), xr.grad) < 2e-3 and rel(dw.cpu(), wr.grad) < 1e-4 and rel(db.cpu(), b.grad) < 1e-5


def test_conv3x3_backward_matches_autograd(lib_built):
    from diffews_b200.backward import conv3x3_backward
    from diffews_b200.weights import conv_weight_to_gemm
    g = torch.Generator().manual_seed(12)
    N, H, Ci, Co = 2, 16, 64, 128
    x = torch.randn(N, Ci, H, H, generator=g).half()
    w = (torch.randn(Co, Ci, 3, 3, generator=g) * (9 * Ci) ** -0.5).half()
    dy = torch.randn(N, Co, H, H, generator=g).half()
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    b = torch.zeros(Co, requires_grad=True)
    torch.nn.functional.conv2d(xr, wr, b, padding=1).backward(dy.float())
    dx, dw, db = conv3x3_backward(x.permute(0, 2, 3, 1).contiguous().cuda(), conv_weight_to_gemm(w).cuda().half(),
                                  dy.permute(0, 2, 3, 1).contiguous().cuda())
    assert rel(dx.cpu().permute(0, 3, 1, 2), xr.grad) < 2e-3
    assert rel(dw.cpu(), conv_weight_to_gemm(wr.grad)) < 1e-4
    assert rel(db.cpu(), b.grad) < 1e-5
