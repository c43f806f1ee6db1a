"""Generates tests/golden/madnet_64x128.npz from the CPU oracle (seeded weights + synthetic pair).

The reference (TF 1.12) cannot run here and ships no fixtures, so these vectors pin the ORACLE (parity
unpinned w.r.t. TensorFlow itself).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200'))
from madstereo.synthetic import make_pair  # noqa: E402
from oracle.adaptation import OracleAdapter  # noqa: E402
from oracle.madnet import MadNetOracle, init_params  # noqa: E402
from oracle import tf1_ops as T  # noqa: E402


def main():
    torch.set_num_threads(1)   # fixed summation order
    h, w = 64, 128
    left, right, _ = make_pair(h, w, seed=3)
    params = init_params(seed=42)
    net = MadNetOracle(params)
    disps, layers = net.forward(left, right)
    out = {'left': left.astype(np.float16), 'right': right.astype(np.float16)}
    # inputs are stored as fp16-rounded values: regenerate exactly by casting back to fp32
    left = out['left'].astype(np.float32); right = out['right'].astype(np.float32)
    disps, layers = net.forward(left, right)
    for i, d in enumerate(disps):
        out['disp%d' % i] = d.numpy()
    for k in ('left/conv4', 'right/conv12', 'corr6', 'corr2', 'fgc-volume-filtering-4/disp3', 'context5', 'final_disp'):
        out['layer:' + k] = layers[k].numpy()
    out['full_loss'] = np.float32(T.reprojection_loss(disps[-1], torch.tensor(left), torch.tensor(right)))
    for mode in ('MAD', 'FULL'):
        for k in (range(5) if mode == 'MAD' else [0]):
            ad = OracleAdapter(params, mode=mode, lr=1e-4)
            o = ad.step(left, right, k)
            out['%s%d:train_loss' % (mode, k)] = np.float32(o['train_loss'])
            for n, g in o['grads'].items():
                if n.endswith('biases'):
                    out['%s%d:grad:%s' % (mode, k, n)] = g
                else:   # weights: keep norm + a fixed slice to stay small
                    out['%s%d:gnorm:%s' % (mode, k, n)] = np.float32(np.sqrt((g.astype(np.float64) ** 2).sum()))
                    out['%s%d:gslice:%s' % (mode, k, n)] = g.reshape(-1)[:64].copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'madnet_64x128.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
