"""FPS of the other BASELINE.json configurations on one GPU (not the headline bench line): MADNet forward only (NONE),
MADNet FULL back-propagation, DispNet forward only and DispNet FULL, all at 1280x384 with synthetic pairs."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200')
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)
import torch
import Nets
from madstereo.adaptation import OnlineAdaptation
from madstereo.synthetic import make_pair, init_params

H, W = 384, 1280
dev = torch.device('cuda', 0)
pairs = [make_pair(H, W, seed=i)[:2] for i in range(4)]
dpairs = [(torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)) for l, r in pairs]
out = {}
for name, mode in (('MADNet', 'NONE'), ('MADNet', 'FULL'), ('MADNet', 'MAD'), ('Dispnet', 'NONE'), ('Dispnet', 'FULL')):
    args = dict(left_img=dpairs[0][0], right_img=dpairs[0][1], split_layers=[None], sequence=True, train_portion='BEGIN',
                bulkhead=(mode == 'MAD'), is_training=False)
    if name == 'MADNet':
        args.update(warping=True, context_net=True, radius_d=2, stride=1)
    else:
        args.update(correlation=True)
    net = Nets.get_stereo_net(name, args)
    kw = {}
    if mode == 'MAD':
        kw = dict(train_config=json.load(open(os.path.join(PKG, 'block_config', 'MadNet_full.json'))), sample_mode='SEQUENTIAL')
    ad = OnlineAdaptation(net, mode=mode, lr=1e-4, **kw)
    ad.load_weights(init_params(net.engine.layers, seed=42))
    steps = 30
    for i in range(6):
        ad.step(*dpairs[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        ad.step(*dpairs[i % 4])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out['%s_%s' % (name, mode)] = {'ms_per_frame': ms, 'fps': 1e3 / ms}
    print('%-8s %-5s %8.3f ms/frame  %7.1f FPS' % (name, mode, ms, 1e3 / ms), file=sys.stderr)
    del ad, net
    torch.cuda.empty_cache()
print(json.dumps(out))
