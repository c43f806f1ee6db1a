"""tcgen05.mma cost table (diagnosis): cycles per MMA (M = 128, K = 16, bf16) for operand layouts x N x accumulator rotation.
  python scripts/mma_probe.py > gpurun_out/mma_probe.log
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from ctypes import c_void_p
from madstereo._lib import lib, check

L = lib()
out = torch.zeros(2 * 148, dtype=torch.int64, device='cuda')
st = c_void_p(torch.cuda.current_stream().cuda_stream)
ITERS = 512
print('cycles per MMA, %d back-to-back MMAs per CTA; columns: layout(A,B) N accs rot | 1 CTA: issue, retire | 148 CTAs: issue, retire (mean)' % ITERS)
for a_mn, b_mn in ((0, 0), (1, 0), (0, 1), (1, 1)):
    for n in (32, 64, 128, 256):
        for n_acc, rot in ((1, 1), (min(4, 512 // n), 1), (1, 0)):
            row = []
            for ctas in (1, 148):
                for _ in range(2):
                    check(L.ms_debug_mma_probe(a_mn, b_mn, n, n_acc, rot, ITERS, ctas, c_void_p(out.data_ptr()), st), 'probe')
                torch.cuda.synchronize()
                v = out[:2 * ctas].view(ctas, 2).double().mean(0) / ITERS
                row += [float(v[0]), float(v[1])]
            print('A %s B %s  N %3d  accs %d rot %d | %7.1f %7.1f | %7.1f %7.1f' % ('MN' if a_mn else 'K ', 'MN' if b_mn else 'K ', n, n_acc, rot, *row), flush=True)
