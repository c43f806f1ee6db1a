"""GPU parity of the tcgen05 (3xTF32) convolution path vs the CPU oracle and vs the exact-fp32 CUDA-core path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()


TC_CASES = [
    # n, h, w, cin, cout, k, dil, alpha
    (1, 8, 16, 32, 32, 3, 1, 0.2),          # exactly one 8x16 tile, one K block per tap
    (1, 16, 32, 64, 64, 3, 1, 0.2),
    (2, 24, 40, 128, 128, 3, 1, 0.2),       # partial tiles in x, batch 2
    (1, 24, 48, 128, 96, 3, 4, 0.2),        # dilated (context net)
    (1, 24, 48, 96, 64, 3, 16, 0.2),
    (1, 12, 40, 136, 128, 3, 1, 0.2),       # K not a multiple of 32 (zero-filled channels)
    (1, 6, 20, 192, 192, 3, 1, 0.2),        # level-6 sized map, N=192
    (2, 20, 36, 16, 16, 3, 1, 0.2),         # half-empty K block, N=16
    (1, 16, 16, 128, 64, 1, 1, 0.1),        # 1x1 (DispNet conv_redir)
    (1, 96, 320, 128, 128, 3, 1, 0.2),      # the dominant MADNet layer shape
    (1, 16, 32, 72, 128, 3, 1, 0.2),        # estimator disp-1 at level 3: dgrad N=72 (BN=80)
    (1, 16, 32, 36, 128, 3, 1, 0.2),        # context-1 like: dgrad N=36 (BN=48)
    (1, 8, 16, 136, 128, 3, 1, 0.2),
]


@pytest.mark.parametrize('case', TC_CASES)
def test_conv_tc_forward_and_dgrad(case):
    from madstereo import ops
    from oracle import tf1_ops as T
    n, h, w, cin, cout, k, dil, alpha = case
    rng = np.random.default_rng(sum(case[:7]))
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
    b = rng.uniform(-0.1, 0.1, cout).astype(np.float32)
    xt = torch.tensor(x, requires_grad=True)
    ref = T.conv2d(xt, torch.tensor(wt), torch.tensor(b), stride=1, dilation=dil, alpha=alpha)
    out = ops.conv2d_tc(cu(x), cu(wt), cu(b), dil, alpha)
    torch.cuda.synchronize()
    exact = ops.conv2d(cu(x), cu(wt), cu(b), 1, dil, alpha)
    assert rel_linf(out.cpu().numpy(), ref.detach().numpy()) < 3e-5, 'fwd vs oracle'
    assert rel_linf(out.cpu().numpy(), exact.cpu().numpy()) < 3e-5, 'fwd vs fp32 path'
    pre = T.conv2d(xt, torch.tensor(wt), torch.tensor(b), stride=1, dilation=dil, alpha=None)
    g = rng.standard_normal(pre.shape).astype(np.float32)
    (gx,) = torch.autograd.grad(pre, xt, grad_outputs=torch.tensor(g))
    dx = ops.conv2d_dgrad_tc(cu(g), cu(wt), dil)
    torch.cuda.synchronize()
    assert rel_linf(dx.cpu().numpy(), gx.numpy()) < 3e-5, 'dgrad vs oracle'


WG_CASES = [
    # n, h, w, cin, cout, k, dil
    (1, 16, 32, 64, 32, 3, 1), (1, 24, 48, 128, 128, 3, 1), (2, 24, 40, 128, 96, 3, 4), (1, 16, 32, 72, 128, 3, 1),
    (1, 12, 40, 136, 128, 3, 1), (1, 32, 64, 64, 32, 3, 1), (2, 16, 20, 192, 192, 3, 1), (1, 16, 16, 128, 64, 1, 1),
    (1, 96, 320, 128, 128, 3, 1),
]


@pytest.mark.parametrize('case', WG_CASES)
def test_wgrad_tc(case):
    from madstereo import ops
    from oracle import tf1_ops as T
    n, h, w, cin, cout, k, dil = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    g = rng.standard_normal((n, h, w, cout)).astype(np.float32)
    wt = torch.zeros(k, k, cin, cout, requires_grad=True)
    bt = torch.zeros(cout, requires_grad=True)
    pre = T.conv2d(torch.tensor(x), wt, bt, stride=1, dilation=dil, alpha=None)
    gw, gb = torch.autograd.grad(pre, [wt, bt], grad_outputs=torch.tensor(g))
    dw, db = ops.conv2d_wgrad_tc(cu(x), cu(g), k, k, dil)
    torch.cuda.synchronize()
    assert rel_linf(dw.cpu().numpy(), gw.numpy()) < 5e-5, 'dw'
    assert rel_linf(db.cpu().numpy(), gb.numpy()) < 5e-5, 'db'
