"""Plane-level micro-benchmark of the split-bf16 tcgen05 kernels (conv_bf forward, wgrad_bf) on MADNet / DispNet layer
shapes: operands pre-split as in the engine's steady state, CUDA events, L2 flushed between repetitions.
  python scripts/bf_bench.py            # table
  python scripts/bf_bench.py one <i>    # 3 launches of shape i (the ncu target)
"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from ctypes import c_void_p
from madstereo._lib import lib, check

L = lib()
dev = 'cuda'
P = lambda t: c_void_p(t.data_ptr() if t is not None else 0)
st = lambda: c_void_p(torch.cuda.current_stream().cuda_stream)

SHAPES = [  # n, h, w, cin, cout, k, stride, dil
    (1, 96, 320, 128, 128, 3, 1, 1), (1, 96, 320, 128, 128, 3, 1, 4), (1, 96, 320, 128, 96, 3, 1, 8), (1, 96, 320, 96, 64, 3, 1, 16),
    (1, 96, 320, 38, 128, 3, 1, 1), (1, 96, 320, 64, 32, 3, 1, 1), (1, 48, 160, 128, 128, 3, 1, 1), (1, 24, 80, 128, 128, 3, 1, 1),
    (1, 6, 20, 197, 128, 3, 1, 1), (2, 96, 320, 32, 32, 3, 1, 1), (2, 192, 640, 16, 32, 3, 2, 1), (2, 48, 160, 64, 96, 3, 2, 1),
    (1, 96, 320, 128, 256, 5, 2, 1), (1, 24, 80, 256, 512, 3, 2, 1), (1, 12, 40, 512, 512, 3, 1, 1),
    # 15..: DispNet conv2 (5x5 stride 2) and a stride-1 problem of the same output size; DispNet stem; stride-2 3x3 pair
    (2, 192, 640, 64, 128, 5, 2, 1), (2, 96, 320, 64, 128, 5, 1, 1), (2, 384, 1280, 3, 64, 7, 2, 1),
    (2, 192, 640, 64, 128, 3, 2, 1), (2, 96, 320, 64, 128, 3, 1, 1),
]


def planes(t, c, fmt):
    pcs = (c + 7) // 8 * 8
    n, h, w, _ = t.shape
    hi = torch.empty(n * h * w * pcs, dtype=torch.bfloat16, device=dev); lo = torch.empty_like(hi)
    check(L.ms_bf_split(P(t), n, h, w, c, c, P(hi), P(lo), pcs, fmt, 0.0625, st()), 'split')
    return hi, lo, pcs


def timeit(fn, flush, reps=15):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def setup(shape):
    n, h, w, cin, cout, k, s, d = shape
    oh, ow = (h + s - 1) // s, (w + s - 1) // s
    x = torch.randn(n, h, w, cin, device=dev); wt = torch.randn(k, k, cin, cout, device=dev) * 0.05
    b = torch.zeros(cout, device=dev); y = torch.empty(n, oh, ow, cout, device=dev)
    g = torch.randn(n, oh, ow, cout, device=dev)
    xh, xl, xpcs = planes(x, cin, 1); xbh, xbl, _ = planes(x, cin, 0); gh, gl, gpcs = planes(g, cout, 0)
    ypcs = (cout + 7) // 8 * 8
    yh = torch.empty(n * oh * ow * ypcs, dtype=torch.bfloat16, device=dev); yl = torch.empty_like(yh)
    halfs = L.ms_bf_weight_halfs(k * k, cout, cin)
    wt16 = torch.empty(halfs, dtype=torch.bfloat16, device=dev)
    job = torch.empty(256, dtype=torch.uint8, device=dev)
    check(L.ms_bf_prep_weights(P(wt), k * k, cin, cout, 0, 1, P(wt16), P(job), st()), 'prep')
    part = torch.empty(L.ms_conv2d_bf_part_floats(), device=dev)
    tick = torch.zeros(L.ms_conv2d_bf_ticket_words(), dtype=torch.int32, device=dev)
    nws = L.ms_conv2d_wgrad_bf_workspace(k, k, cin, cout)
    ws = torch.empty(nws, device=dev); dw = torch.empty(k, k, cin, cout, device=dev); db = torch.empty(cout, device=dev)
    def fwd():
        check(L.ms_conv2d_fwd_bf_planes(P(xh), P(xl), xpcs, 1, 0.0625, n, h, w, cin, P(wt16), P(b), P(y), cout, cout, P(yh), P(yl), ypcs,
                                        k, k, s, d, 0.2, P(part), P(tick), st()), 'fwd')
    def wgrad():
        check(L.ms_conv2d_wgrad_bf_planes(P(xbh), P(xbl), xpcs, n, h, w, cin, P(gh), P(gl), gpcs, oh, ow, cout, P(dw), P(db),
                                          k, k, s, d, P(ws), nws, st()), 'wgrad')
    keep = (x, wt, b, y, g, xh, xl, xbh, xbl, gh, gl, yh, yl, wt16, job, part, tick, ws, dw, db)
    return fwd, wgrad, n * oh * ow * k * k * cin * cout, keep


if len(sys.argv) > 2 and sys.argv[1] == 'prof':
    # MS_BF_PROF=1 python scripts/bf_bench.py prof <i>: per-CTA clock64 breakdown of the forward kernel on shape i
    import ctypes, numpy as np
    fwd, wgrad, macs, keep = setup(SHAPES[int(sys.argv[2])])
    for _ in range(5):
        fwd()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (8 * 8192))()
    n = L.ms_debug_bf_prof(buf, 8192)
    a = np.frombuffer(buf, dtype=np.uint64)[:n * 8].reshape(n, 8).astype(np.int64)
    names = ['setup', 'first data', 'main loop (issue)', 'drain to accumulator barrier', 'epilogue', 'exit sync']
    d = np.stack([a[:, 1] - a[:, 0], a[:, 2] - a[:, 1], a[:, 3] - a[:, 2], a[:, 4] - a[:, 3], a[:, 5] - a[:, 4], a[:, 6] - a[:, 5]], 1)
    print('shape', SHAPES[int(sys.argv[2])], 'CTAs', n)
    for i, nm in enumerate(names):
        print('  %-30s mean %8.0f  min %8.0f  max %8.0f cycles' % (nm, d[:, i].mean(), d[:, i].min(), d[:, i].max()))
    print('  %-30s mean %8.0f cycles (of the main loop)' % ('MMA thread waiting on data', a[:, 7].mean()))
    print('  %-30s mean %8.0f cycles' % ('CTA lifetime', (a[:, 6] - a[:, 0]).mean()))
    sys.exit(0)

if len(sys.argv) > 2 and sys.argv[1] == 'one':
    fwd, wgrad, macs, keep = setup(SHAPES[int(sys.argv[2])])
    for _ in range(3):
        fwd(); wgrad()
    torch.cuda.synchronize()
    sys.exit(0)

flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rows = []
sel = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 and sys.argv[1] == 'sel' else None
for si, shape in enumerate(SHAPES):
    if sel is not None and si not in sel:
        continue
    fwd, wgrad, macs, keep = setup(shape)
    tf = timeit(fwd, flush); tw = timeit(wgrad, flush)
    rows.append({'shape': shape, 'fwd_us': tf, 'fwd_tflops': 2 * macs / tf / 1e6, 'wgrad_us': tw, 'wgrad_tflops': 2 * macs / tw / 1e6})
    print('n%d %3dx%3d %3d->%3d k%d s%d dil%-2d | fwd %7.1f us %6.1f useful TFLOP/s | wgrad(+reduce) %7.1f us %6.1f TFLOP/s' % (
        *shape, tf, 2 * macs / tf / 1e6, tw, 2 * macs / tw / 1e6), flush=True)
    del keep
out = os.path.join(ROOT, 'gpurun_out', 'bf_bench.json')
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, 'w'), indent=1)
