"""Steady-state cost of one conv_bf launch inside a CUDA graph (warm L2, programmatic dependent launch, no host gaps):
50 back-to-back launches of the same layer captured into one graph, replayed 20 times.
  python scripts/chain_bench.py            # table over small / large maps
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from ctypes import c_void_p
from madstereo._lib import lib, check

L = lib()
dev = 'cuda'
P = lambda t: c_void_p(t.data_ptr() if t is not None else 0)
SHAPES = [(1, 6, 20, 192, 192, 3), (1, 12, 40, 128, 128, 3), (1, 24, 80, 128, 128, 3), (1, 48, 160, 128, 128, 3), (1, 96, 320, 128, 128, 3),
          (1, 96, 320, 64, 32, 3), (2, 96, 320, 32, 32, 3)]
NL = 50
for shape in SHAPES:
    n, h, w, cin, cout, k = shape
    x = torch.randn(n, h, w, cin, device=dev); wt = torch.randn(k, k, cin, cout, device=dev) * 0.05
    b = torch.zeros(cout, device=dev); y = torch.empty(n, h, w, cout, device=dev)
    pcs = (cin + 7) // 8 * 8; ypcs = (cout + 7) // 8 * 8
    xh = torch.empty(n * h * w * pcs, dtype=torch.bfloat16, device=dev); xl = torch.empty_like(xh)
    yh = torch.empty(n * h * w * ypcs, dtype=torch.bfloat16, device=dev); yl = torch.empty_like(yh)
    s0 = c_void_p(torch.cuda.current_stream().cuda_stream)
    check(L.ms_bf_split(P(x), n, h, w, cin, cin, P(xh), P(xl), pcs, 1, 0.0625, s0), 'split')
    halfs = L.ms_bf_weight_halfs(k * k, cout, cin)
    wt16 = torch.empty(halfs, dtype=torch.bfloat16, device=dev); job = torch.empty(256, dtype=torch.uint8, device=dev)
    check(L.ms_bf_prep_weights(P(wt), k * k, cin, cout, 0, 1, P(wt16), P(job), s0), 'prep')
    part = torch.empty(L.ms_conv2d_bf_part_floats(), device=dev)
    tick = torch.zeros(L.ms_conv2d_bf_ticket_words(), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    def launch(st):
        check(L.ms_conv2d_fwd_bf_planes(P(xh), P(xl), pcs, 1, 0.0625, n, h, w, cin, P(wt16), P(b), P(y), cout, cout, P(yh), P(yl), ypcs,
                                        k, k, 1, 1, 0.2, P(part), P(tick), c_void_p(st.cuda_stream)), 'fwd')
    with torch.cuda.stream(side):
        for _ in range(3): launch(side)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(NL): launch(side)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (20 * NL)
    macs = n * h * w * k * k * cin * cout
    print('n%d %3dx%3d %3d->%3d k%d | %6.2f us per launch in a graph | %6.1f useful TFLOP/s' % (n, h, w, cin, cout, k, us, 2 * macs / us / 1e6), flush=True)
