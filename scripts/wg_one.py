import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
import torch
from madstereo import ops
torch.manual_seed(0)
n, h, w, cin, cout = 1, 16, 32, 64, 32
def run(x, g, name):
    dw, db = ops.conv2d_wgrad_tc(x, g, 3, 3, 1)
    torch.cuda.synchronize()
    dw2, db2 = ops.conv2d_wgrad(x, g, 3, 3, 1, 1)
    print(name, 'max err %.4g of %.4g' % (float((dw - dw2).abs().max()), float(dw2.abs().max())))
    print('  tc  tap4 [ci0..3][co0..3]:', dw[1, 1, :3, :4].flatten().tolist())
    print('  ref tap4 [ci0..3][co0..3]:', dw2[1, 1, :3, :4].flatten().tolist())
    return dw, dw2
ones_x = torch.ones(n, h, w, cin, device='cuda'); ones_g = torch.ones(n, h, w, cout, device='cuda')
xr = torch.randn(n, h, w, cin, device='cuda'); gr = torch.randn(n, h, w, cout, device='cuda')
run(ones_x, ones_g, 'ones/ones')
# B path: x ones, g depends only on co
gco = torch.arange(cout, device='cuda', dtype=torch.float32).view(1, 1, 1, cout).expand(n, h, w, cout).contiguous()
run(ones_x, gco, 'x=1, g=co index')
# B path pixel dependence: g = pixel x coordinate
gpx = torch.arange(w, device='cuda', dtype=torch.float32).view(1, 1, w, 1).expand(n, h, w, cout).contiguous()
run(ones_x, gpx, 'x=1, g=px')
xci = torch.arange(cin, device='cuda', dtype=torch.float32).view(1, 1, 1, cin).expand(n, h, w, cin).contiguous()
run(xci, ones_g, 'x=ci index, g=1')
xpx = torch.arange(w, device='cuda', dtype=torch.float32).view(1, 1, w, 1).expand(n, h, w, cin).contiguous()
dw, dw2 = run(xpx, gpx, 'x=px, g=px')
run(xr, gr, 'random')
