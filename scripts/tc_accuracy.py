"""Diagnostic: error of the fp32 CUDA-core conv and the tcgen05 3xTF32 conv against an fp64 CPU result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import numpy as np, torch
from madstereo import ops
from oracle import tf1_ops as T

def cu(x): return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).cuda()
for (n, h, w, cin, cout, k, dil) in [(1, 48, 64, 128, 128, 3, 1), (1, 48, 64, 32, 32, 3, 1), (1, 48, 64, 128, 128, 3, 4)]:
    rng = np.random.default_rng(0)
    for kind in ('gauss', 'positive'):
        x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
        wt = (rng.standard_normal((k, k, cin, cout)) / np.sqrt(k * k * cin)).astype(np.float32)
        if kind == 'positive':
            x = np.abs(x); wt = np.abs(wt)
        b = np.zeros(cout, np.float32)
        ref = T.conv2d(torch.tensor(x).double(), torch.tensor(wt).double(), torch.tensor(b).double(), 1, dil, None).numpy()
        a = ops.conv2d(cu(x), cu(wt), cu(b), 1, dil, 1.0).cpu().numpy().astype(np.float64)
        t = ops.conv2d_tc(cu(x), cu(wt), cu(b), dil, 1.0).cpu().numpy().astype(np.float64)
        sc = np.abs(ref).max()
        for name, y in (('fp32', a), ('tc', t)):
            e = y - ref
            print('%s cin=%d dil=%d %-5s max|e|/max|ref| %.2e  rms(e)/rms(ref) %.2e  mean(e*sign(ref))/mean|ref| %.2e' % (
                kind, cin, dil, name, np.abs(e).max() / sc, np.sqrt((e ** 2).mean()) / np.sqrt((ref ** 2).mean()),
                (e * np.sign(ref)).mean() / np.abs(ref).mean()))
