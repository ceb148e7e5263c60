"""GPU parity of the gradient kernels (csrc/grad.cu) against torch autograd / plain torch ops on the same inputs.
ref: accelerator.backward(loss), train_tools/train_icl_multitask_nocrop_nearest_nshot_v3.py:1386."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from diffews_b200 import ops
    return ops


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _gemm_w(w):          # torch [Cout, Cin, kh, kw] -> [Cout, (kh*ks+kw)*Cin + ci]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("N,H,W,Cin,Cout,ks,stride", [
    (2, 16, 16, 64, 128, 3, 1),        # one co tile, one ci tile
    (2, 16, 16, 320, 320, 3, 1),       # ragged co (3 tiles) and ci (NB = 3) tiles
    (3, 8, 8, 640, 128, 3, 1),         # 64-pixel images
    (4, 4, 4, 128, 64, 3, 1),          # TN = 4: a chunk spans four images (no wrap across images at the borders)
    (1, 32, 32, 128, 256, 1, 1),       # 1x1
    (2, 16, 16, 128, 128, 3, 2),       # stride 2 (Downsample2D)
    (1, 24, 40, 64, 64, 3, 1),         # ragged pixel chunks
    (1, 64, 64, 320, 8, 3, 1),         # tiny Cout (first rows only)
])
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_conv_wgrad_matches_autograd(N, H, W, Cin, Cout, ks, stride, dt):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(dt)
    dy = (torch.randn(N, H // stride, W // stride, Cout, device="cuda", generator=g) * 0.1).to(dt)
    w = torch.zeros(Cout, Cin, ks, ks, device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=ks // 2)
    y.backward(dy.float().permute(0, 3, 1, 2))
    ref = _gemm_w(w.grad)
    dw = torch.full((Cout, ks * ks * Cin), 7.0, device="cuda")
    ops.conv_wgrad(x, dy, dw, ksize=ks, stride=stride)
    assert _rel(dw, ref) < 2e-5, _rel(dw, ref)
    # accumulate + scale
    ops.conv_wgrad(x, dy, dw, ksize=ks, stride=stride, scale=0.5, accumulate=True)
    assert _rel(dw, 1.5 * ref) < 2e-5
    # determinism
    dw2 = torch.empty_like(dw)
    ops.conv_wgrad(x, dy, dw2, ksize=ks, stride=stride)
    dw3 = torch.empty_like(dw)
    ops.conv_wgrad(x, dy, dw3, ksize=ks, stride=stride)
    assert torch.equal(dw2, dw3)


def test_linear_wgrad_ragged_rows():
    ops = _ops()
    x = torch.randn(2, 77, 1024, device="cuda").half()
    dy = torch.randn(2, 77, 320, device="cuda").half()
    dw = torch.empty(320, 1024, device="cuda")
    ops.linear_wgrad(x, dy, dw)
    ref = dy.float().reshape(-1, 320).t() @ x.float().reshape(-1, 1024)
    assert _rel(dw, ref) < 2e-5


def test_conv_dgrad_through_permuted_weights():
    """dx of a 3x3 / stride-1 conv = conv of dy with the rotated, transposed filter (dfw_weight_permute)."""
    ops = _ops()
    N, H, W, Cin, Cout = 2, 16, 16, 128, 192
    x = torch.randn(N, Cin, H, W, device="cuda", requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    dy = torch.randn(N, H, W, Cout, device="cuda").half()
    F.conv2d(x, w.half().float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    wr = ops.weight_permute(_gemm_w(w).half(), Cout, 9, Cin, [8 - t for t in range(9)])
    dx = ops.conv2d(dy, wr.view(Cin, 9 * Cout), ksize=3)
    assert _rel(dx.float().permute(0, 3, 1, 2), x.grad) < 2e-3
    # Linear: transpose
    wl = torch.randn(192, 128, device="cuda").half()
    wt = ops.weight_permute(wl, 192, 1, 128, [0])
    assert torch.equal(wt.view(128, 192), wl.t().contiguous())


def test_stride2_dgrad_as_phase_convolutions():
    ops = _ops()
    N, H, W, Cin, Cout = 2, 16, 16, 128, 128
    x = torch.randn(N, Cin, H, W, device="cuda", requires_grad=True)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
    dy = torch.randn(N, H // 2, W // 2, Cout, device="cuda").half()
    F.conv2d(x, w.half().float(), stride=2, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    wg = _gemm_w(w).half()
    k_of = {(0, 0): -1, (0, 1): 1, (1, 0): 2, (1, 1): 0}          # (output phase, tap a) -> forward kernel index
    w4 = torch.empty(4, Cin, 4 * Cout, device="cuda", dtype=torch.float16)
    for ph in range(2):
        for pw in range(2):
            tm = []
            for a in range(2):
                for b in range(2):
                    kh, kw = k_of[(ph, a)], k_of[(pw, b)]
                    tm.append(-1 if kh < 0 or kw < 0 else kh * 3 + kw)
            ops.weight_permute(wg, Cout, 9, Cin, tm, out=w4[ph * 2 + pw])
    dx = ops.upconv2x(dy, w4)
    assert _rel(dx.float().permute(0, 3, 1, 2), x.grad) < 2e-3


def test_colsum_downsum_split_geglu_pad():
    ops = _ops()
    x = torch.randn(4, 32, 32, 320, device="cuda").half()
    s = ops.colsum(x)
    assert _rel(s, x.float().sum((0, 1, 2))[None]) < 1e-5
    s4 = ops.colsum(x, groups=4)
    assert _rel(s4, x.float().sum((1, 2))) < 1e-5
    ops.colsum(x, groups=4, out=s4, scale=2.0, accumulate=True)
    assert _rel(s4, 3 * x.float().sum((1, 2))) < 1e-5
    xf = torch.randn(8, 1280, device="cuda")
    assert _rel(ops.colsum(xf), xf.sum(0, keepdim=True)) < 1e-5
    dy = torch.randn(2, 16, 16, 64, device="cuda").half()
    ref = F.avg_pool2d(dy.float().permute(0, 3, 1, 2), 2) * 4
    assert _rel(ops.downsum2x(dy).float().permute(0, 3, 1, 2), ref) < 1e-3
    y = torch.randn(2, 8, 8, 320 + 640, device="cuda").half()
    a, b = ops.split_channels(y, 320)
    assert torch.equal(a, y[..., :320]) and torch.equal(b, y[..., 320:])
    h = torch.randn(3, 100, 2 * 640, device="cuda").half()
    ref = h[..., :640].float() * F.gelu(h[..., 640:].float())
    assert _rel(ops.geglu(h), ref) < 1e-3
    g = torch.randn(2, 4, 8, 8, device="cuda")
    p = ops.nchw_to_nhwc_pad(g, 64, torch.float16, scale=2.0)
    assert p.shape == (2, 8, 8, 64) and float(p[..., 4:].abs().max()) == 0.0
    assert _rel(p[..., :4].float(), 2 * g.permute(0, 2, 3, 1)) < 1e-3
