"""CPU oracle — DispNet-C forward restated from Nets/DispNet.py (TEST INFRASTRUCTURE).

Pinned against the reference's own Nets/DispNet.py executed over oracle/tf1_shim.py (tests/golden/reference_graph_dispnet_64x128.npz);
TF's conv-padding / resize kernels themselves stay unpinned (see oracle/tf1_ops.py).

Follows /root/reference/Nets/DispNet.py:39-43 (_make_disp), :45-57 (_upsampling_block), :59-73 (_preprocess_inputs),
:75-152 (_build_network, correlation=True branch).  Activations: sharedLayers.conv2d default leaky 0.1
(Nets/sharedLayers.py:54), variables `<scope>/weights`, `<scope>/bias`.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import tf1_ops as T

MAX_DISP = 40          # DispNet.py:7
ALPHA = 0.1

# (scope, kh, cin, cout, stride)   encoder, in graph-construction order
ENCODER = [
    ('conv1', 7, 3, 64, 2), ('conv2', 5, 64, 128, 2), ('conv_redir', 1, 128, 64, 1),
    ('conv3', 5, 2 * MAX_DISP + 1 + 64, 256, 2), ('conv3/1', 3, 256, 256, 1), ('conv4', 3, 256, 512, 2),
    ('conv4/1', 3, 512, 512, 1), ('conv5', 3, 512, 512, 2), ('conv5/1', 3, 512, 512, 1),
    ('conv6', 3, 512, 1024, 2), ('conv6/1', 3, 1024, 1024, 1),
]
# (name, bottom layer, skip layer, in, out, skip channels)
UPS = [('up5', 'conv6/1', 'conv5/1', 1024, 512, 512), ('up4', 'up5/concat', 'conv4/1', 512, 256, 512),
       ('up3', 'up4/concat', 'conv3/1', 256, 128, 256), ('up2', 'up3/concat', 'conv2a', 128, 64, 128),
       ('up1', 'up2/concat', 'conv1a', 64, 32, 64)]


def param_shapes():
    p = OrderedDict()
    for scope, k, ci, co, _ in ENCODER:
        p['model/%s/weights' % scope] = (k, k, ci, co)
        p['model/%s/bias' % scope] = (co,)
    for name, _, _, cin, cout, skip in UPS:
        p['model/%s/deconv/weights' % name] = (4, 4, cout, cin); p['model/%s/deconv/bias' % name] = (cout,)
        p['model/%s/predict/weights' % name] = (3, 3, cin, 1); p['model/%s/predict/bias' % name] = (1,)
        p['model/%s/up_predict/weights' % name] = (4, 4, 1, 1); p['model/%s/up_predict/bias' % name] = (1,)
        p['model/%s/concat/weights' % name] = (3, 3, cout + skip + 1, cout); p['model/%s/concat/bias' % name] = (cout,)
    p['model/prediction/weights'] = (3, 3, 32, 1)
    p['model/prediction/bias'] = (1,)
    return p


def init_params(seed=7, bias_range=0.05):
    """Seeded xavier weights with a gain that keeps activations O(1) through the 26-layer net, small biases,
    positive head biases so the relu'd side outputs are alive (test-vector conditioning only)."""
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shp in param_shapes().items():
        if len(shp) == 4:
            w = T.xavier_uniform(rng, shp)
            head = '/predict/' in name or name.startswith('model/prediction') or '/up_predict/' in name
            out[name] = w * np.float32(0.7 if head else 1.6)
        else:
            b = rng.uniform(-bias_range, bias_range, size=shp).astype('float32')
            if '/predict/' in name or name.startswith('model/prediction'):
                b = b + np.float32(0.3)
            out[name] = b
    return out


class DispNetOracle:
    def __init__(self, params, dtype=torch.float32):
        self.p = OrderedDict((k, (v if torch.is_tensor(v) else torch.tensor(np.asarray(v))).to(dtype))
                             for k, v in params.items())
        self.dtype = dtype

    def requires_grad_(self, names=None):
        for k, v in self.p.items():
            v.requires_grad_(names is None or k in names)

    def _conv(self, x, scope, stride=1, alpha=ALPHA):
        return T.conv2d(x, self.p['model/%s/weights' % scope], self.p['model/%s/bias' % scope], stride=stride, alpha=alpha)

    def _deconv(self, x, scope, alpha=ALPHA):
        return T.conv2d_transpose(x, self.p['model/%s/weights' % scope], self.p['model/%s/bias' % scope], 2, alpha)

    def make_disp(self, op, hp, wp, h, w):
        """DispNet._make_disp (DispNet.py:39-43)."""
        scale = wp / op.shape[2]
        return T.crop_or_pad(T.resize_bilinear(torch.relu(op * scale), hp, wp), h, w)

    def forward(self, left, right):
        left = torch.as_tensor(left).to(self.dtype)
        right = torch.as_tensor(right).to(self.dtype)
        h, w = left.shape[1], left.shape[2]
        lp = T.pad_reflect_to_multiple(left / 255.0 - (100.0 / 255), 64)
        rp = T.pad_reflect_to_multiple(right / 255.0 - (100.0 / 255), 64)
        hp, wp = lp.shape[1], lp.shape[2]
        L = OrderedDict()
        L['conv1a'] = self._conv(lp, 'conv1', 2); L['conv1b'] = self._conv(rp, 'conv1', 2)
        L['conv2a'] = self._conv(L['conv1a'], 'conv2', 2); L['conv2b'] = self._conv(L['conv1b'], 'conv2', 2)
        L['conv_redir'] = self._conv(L['conv2a'], 'conv_redir', 1)
        L['corr'] = T.correlation(L['conv2a'], L['conv2b'], MAX_DISP)
        x = torch.cat([L['corr'], L['conv_redir']], -1)
        for scope, _, _, _, s in ENCODER[3:]:
            x = self._conv(x, scope, s)
            L[scope] = x
        disps = []
        for name, bottom, skip, cin, cout, _ in UPS:
            b = L[bottom]
            L[name + '/deconv'] = self._deconv(b, name + '/deconv')
            L[name + '/predict'] = self._conv(b, name + '/predict', 1, None)
            disps.append(self.make_disp(L[name + '/predict'], hp, wp, h, w))
            L[name + '/up_predict'] = self._deconv(L[name + '/predict'], name + '/up_predict', None)
            cat = torch.cat([L[skip], L[name + '/deconv'], L[name + '/up_predict']], -1)
            L[name + '/concat'] = self._conv(cat, name + '/concat', 1, None)
        L['prediction'] = self._conv(L['up1/concat'], 'prediction', 1, None)
        disps.append(self.make_disp(L['prediction'], hp, wp, h, w))
        resc = T.crop_or_pad(T.resize_bilinear(L['prediction'], hp, wp) * 2, h, w)      # DispNet.py:149-151, no relu
        L['rescaled_prediction'] = resc
        disps.append(resc)
        return disps, L


class DispNetAdapter:
    """NONE / FULL adaptation (MAD asserts for DispNet in the reference, Stereo_Online_Adaptation.py:97)."""

    def __init__(self, params, mode='FULL', lr=1e-4, mu=0.9, dtype=torch.float32):
        assert mode in ('NONE', 'FULL')
        self.mode, self.lr, self.mu = mode, lr, mu
        self.net = DispNetOracle(params, dtype=dtype)
        self.momentum = {k: torch.zeros_like(v) for k, v in self.net.p.items()}

    def step(self, left, right, module=None):
        net = self.net
        names = list(net.p.keys()) if self.mode == 'FULL' else []
        net.requires_grad_(set(names))
        lt = torch.as_tensor(left).to(net.dtype); rt = torch.as_tensor(right).to(net.dtype)
        with torch.set_grad_enabled(bool(names)):
            disps, layers = net.forward(lt, rt)
            loss = T.reprojection_loss(disps[-1], lt, rt)
        out = {'full_loss': float(loss.detach()), 'train_loss': float(loss.detach()),
               'disparities': [d.detach().numpy() for d in disps], 'grads': {}}
        if names:
            gl = torch.autograd.grad(loss, [net.p[n] for n in names], allow_unused=True)
            with torch.no_grad():
                for n, g in zip(names, gl):
                    if g is None:
                        continue
                    out['grads'][n] = g.numpy().copy()
                    w, m = T.momentum_update(net.p[n], g, self.momentum[n], self.lr, self.mu)
                    net.p[n].copy_(w)
                    self.momentum[n] = m
        net.requires_grad_(set())
        return out
