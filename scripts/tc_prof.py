import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
os.environ['MS_TC_DEBUG'] = '8'
import torch
from madstereo import ops
from madstereo._lib import lib
L = ctypes.CDLL(os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200', 'libmadstereo.so'))
n, h, w, cin, cout = 1, 96, 320, 128, 128
x = torch.randn(n, h, w, cin, device='cuda'); wt = torch.randn(3, 3, cin, cout, device='cuda') * 0.05; b = torch.zeros(cout, device='cuda')
for _ in range(3): ops.conv2d_tc(x, wt, b, 1, 0.2)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 32)()
L.ms_debug_tc_prof(buf, 1)
R = 10
for _ in range(R): ops.conv2d_tc(x, wt, b, 1, 0.2)
torch.cuda.synchronize()
L.ms_debug_tc_prof(buf, 1)
names = ['producer wait bempty', 'mma wait ready', 'mma issue+commit', 'split wait pfull', 'split wait free', 'split wait bfull', 'split work', 'split fence', 'epilogue', 'mainloop total (splitter view)', 'count', 'epilogue: wait accum']
cnt = max(buf[10], 1)
for i, nm in enumerate(names): print('%-32s %10.0f cycles / tile' % (nm, buf[i] / cnt))
