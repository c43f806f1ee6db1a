import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from madstereo import ops
b, h, w, c, d = 8, 272, 480, 32, 2
x = torch.randn(b, h, w, c, device='cuda'); y = torch.randn(b, h, w, c, device='cuda')
u = torch.rand(b, h, w, 1, device='cuda') * 4 - 2
for _ in range(4):
    out = ops.correlation(x, y, d, 1, u=u)
torch.cuda.synchronize()
