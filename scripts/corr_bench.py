"""Correlation kernel on HBM-sized shapes: DispNet corr (41 MB) and MADNet level-2 at 1920x1088 x batch 8 (288 MB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from madstereo import ops

def bench(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3

for name, (b, h, w, c, d, warp) in {'MADNet L2 1280x384 B1': (1, 96, 320, 32, 2, True), 'MADNet L2 1920x1088 B8': (8, 272, 480, 32, 2, True),
                                   'MADNet L3 1920x1088 B8': (8, 136, 240, 64, 2, True), 'DispNet corr 1280x384': (1, 96, 320, 128, 40, False)}.items():
    x = torch.randn(b, h, w, c, device='cuda'); y = torch.randn(b, h, w, c, device='cuda')
    u = (torch.rand(b, h, w, 1, device='cuda') * 4 - 2) if warp else None
    nd = 2 * d + 1
    if warp:
        t = bench(lambda: ops.cost_volume(x, y, d, 1, u=u))          # fused warp + corr + concat (writes left copy too)
        byts = b * h * w * ((2 * c + nd) * 4 + c * 4 + 8)           # + left copy written + u read/written
        t2 = bench(lambda: ops.correlation(x, y, d, 1, u=u))
        print('%-26s fused concat: %7.1f us  %6.0f GB/s | corr only: %7.1f us %6.0f GB/s (algorithmic B*h*w*(2C+%d)*4)' % (
            name, t * 1e6, byts / t / 1e9, t2 * 1e6, b * h * w * (2 * c + nd) * 4 / t2 / 1e9, nd))
    else:
        t2 = bench(lambda: ops.correlation(x, y, d, 1))
        print('%-26s corr only: %7.1f us %6.0f GB/s' % (name, t2 * 1e6, b * h * w * (2 * c + nd) * 4 / t2 / 1e9))
