"""Probe: does tcgen05.mma kind::f16 accept A = fp16 and B = bf16 in one instruction (instruction-descriptor formats
0 / 1)?  If yes, wgrad_bf can read the forward fp16 planes directly (MS_WGRAD_MIXED=1) instead of a bf16 re-split.
Run in its own process: an illegal instruction would poison the CUDA context."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import numpy as np, torch
from madstereo import ops
from oracle import tf1_ops as T
worst = 0.0
for case in [(1, 16, 32, 64, 64, 3, 1, 1), (2, 24, 40, 128, 128, 3, 1, 1), (1, 48, 160, 64, 96, 3, 2, 1)]:
    n, h, w, cin, cout, k, stride, dil = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = torch.tensor((rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32), requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    pre = T.conv2d(torch.tensor(x), wt, b, stride=stride, dilation=dil, alpha=None)
    g = rng.standard_normal(pre.shape).astype(np.float32)
    gw, gb = torch.autograd.grad(pre, (wt, b), grad_outputs=torch.tensor(g))
    dw, db = ops.conv2d_wgrad_bf(torch.tensor(x).cuda(), torch.tensor(g).cuda(), k, k, stride, dil, x_fmt=1)
    torch.cuda.synchronize()
    e = float(np.abs(dw.cpu().numpy() - gw.numpy()).max() / np.abs(gw.numpy()).max())
    eb = float(np.abs(db.cpu().numpy() - gb.numpy()).max() / np.abs(gb.numpy()).max())
    print('mixed f16 x bf16 wgrad', case, 'rel L-inf dw %.3e db %.3e' % (e, eb), flush=True)
    worst = max(worst, e, eb)
print('MIXED_OK' if worst < 1e-4 else 'MIXED_WRONG', worst)
