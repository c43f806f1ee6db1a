"""3 launches of the DispNet correlation (forward + backward, banded mma.sync kernels) at 1280x384 -- the ncu target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from madstereo import ops
b, h, w, c, d = 1, 96, 320, 128, 40
x = torch.randn(b, h, w, c, device='cuda'); y = torch.randn(b, h, w, c, device='cuda')
g = torch.randn(b, h, w, 2 * d + 1, device='cuda') * 1e-3
out = torch.empty(b, h, w, 2 * d + 1, device='cuda')
for _ in range(3):
    ops.correlation_wide(x, y, d, 64.0, out=out)
    ops.correlation_bwd(x, y, g, d, 1)
torch.cuda.synchronize()
