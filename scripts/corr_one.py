"""ncu target: the MADNet level-2 cost volume at 1920x1088 x 8 frames (288 MB class).
  python scripts/corr_one.py         # plain variant (separate [B,h,w,5] output)
  python scripts/corr_one.py fused   # the variant Engine::forward launches: left copy + corr + u channel into the concat buffer"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from ctypes import c_void_p
from madstereo import ops
from madstereo._lib import lib, check
b, h, w, c, d = 8, 272, 480, 32, 2
x = torch.randn(b, h, w, c, device='cuda'); y = torch.randn(b, h, w, c, device='cuda')
u = torch.rand(b, h, w, 1, device='cuda') * 4 - 2
if len(sys.argv) > 1 and sys.argv[1] == 'fused':
    nd = 2 * d + 1
    ocs = (c + nd + 1 + 3) // 4 * 4
    cost = torch.zeros(b, h, w, ocs, device='cuda')
    cost[..., c + nd] = u[..., 0]
    uu = cost[..., c + nd:]
    P = lambda t: c_void_p(t.data_ptr())
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(4):
        check(lib().ms_corr_fwd(P(x), c, P(y), c, P(uu), ocs, P(cost), ocs, b, h, w, c, d, 1, 1, 1, st), 'corr_fwd')
else:
    for _ in range(4):
        out = ops.correlation(x, y, d, 1, u=u)
torch.cuda.synchronize()
