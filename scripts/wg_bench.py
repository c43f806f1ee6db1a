"""Timing of the tcgen05 weight-gradient kernel on the MADNet level-2 shapes (CUDA events, L2 flushed)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from madstereo import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for (n, h, w, ci, co) in [(1, 96, 320, 128, 128), (1, 96, 320, 128, 96), (1, 96, 320, 64, 32), (1, 48, 160, 128, 128)]:
    x = torch.randn(n, h, w, ci, device='cuda'); g = torch.randn(n, h, w, co, device='cuda')
    for _ in range(3): ops.conv2d_wgrad_tc(x, g, 3, 3, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.conv2d_wgrad_tc(x, g, 3, 3, 1); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); t = ts[len(ts) // 2] * 1e3
    print('wgrad %dx%d %d->%d : %7.1f us (incl. transpose, reduce, bias, allocs)  %6.1f TFLOP/s useful' % (h, w, ci, co, t, 2 * n * h * w * 9 * ci * co / t / 1e6))
