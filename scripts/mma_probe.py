"""tcgen05.mma cost table (diagnosis): cycles per MMA (M = 128, K = 16, bf16) for issue scheme x operand layouts x N.
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
print('cycles per tcgen05.mma (M 128, K 16, bf16), %d back-to-back MMAs per CTA, 148 CTAs; floor = N/2 cycles' % ITERS)
print('issue scheme            layout(A,B)  N   accs | issue loop | until retired')
for uni in (0, 1):
    for a_mn, b_mn in ((0, 0), (1, 1), (1, 0), (0, 1)):
        for n in (32, 64, 128, 256):
            for n_acc in (1, min(4, 512 // n)):
                for _ in range(2):
                    check(L.ms_debug_mma_probe(a_mn, b_mn, n, n_acc, 1, ITERS, uni, 148, c_void_p(out.data_ptr()), st), 'probe')
                torch.cuda.synchronize()
                v = out.view(148, 2).double().mean(0) / ITERS
                print('%-22s  A %s B %s  %3d   %d   | %8.1f   | %8.1f' % ('warp-uniform + elect' if uni else 'lane-0 loop', 'MN' if a_mn else 'K ',
                                                                        'MN' if b_mn else 'K ', n, n_acc, float(v[0]), float(v[1])), flush=True)
