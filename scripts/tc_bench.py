"""Micro-benchmark of the conv kernels on the dominant MADNet layer shapes (CUDA events, L2 flushed between reps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import numpy as np, torch
from madstereo import ops
from madstereo._lib import lib, check
from ctypes import c_void_p

def bench(fn, reps=20, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us

flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
shapes = [(1, 96, 320, 128, 128, 1), (1, 96, 320, 128, 128, 4), (1, 96, 320, 128, 96, 1), (1, 96, 320, 40, 128, 1),
          (1, 96, 320, 64, 32, 1), (2, 192, 640, 16, 16, 1), (2, 96, 320, 32, 32, 1), (1, 48, 160, 128, 128, 1), (2, 6, 20, 192, 192, 1)]
sel = int(sys.argv[1]) if len(sys.argv) > 1 else None
for i, (n, h, w, cin, cout, dil) in enumerate(shapes):
    if sel is not None and sel != i: continue
    x = torch.randn(n, h, w, cin, device='cuda'); wt = torch.randn(3, 3, cin, cout, device='cuda') * 0.05
    b = torch.zeros(cout, device='cuda')
    macs = n * h * w * 9 * cin * cout
    y = torch.empty(n, h, w, cout, device='cuda')
    ns = lib().ms_conv2d_tc_scratch(3, 3, cin, cout); scratch = torch.empty(ns, device='cuda')
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    def tc():
        check(lib().ms_conv2d_fwd_tc(c_void_p(x.data_ptr()), n, h, w, cin, cin, c_void_p(wt.data_ptr()), c_void_p(b.data_ptr()),
                                     c_void_p(y.data_ptr()), cout, cout, 3, 3, dil, 0.2, c_void_p(scratch.data_ptr()), ns, st), 'tc')
    t_tc = bench(tc, flush=flush)
    t_f32 = bench(lambda: ops.conv2d(x, wt, b, 1, dil, 0.2), flush=flush)
    print('n%d %3dx%3d %3d->%3d dil%-2d  tc(+prep) %7.1f us  %6.1f TFLOP/s useful | fp32 %7.1f us %5.1f TFLOP/s' % (
        n, h, w, cin, cout, dil, t_tc, 2 * macs / t_tc / 1e6, t_f32, 2 * macs / t_f32 / 1e6))
