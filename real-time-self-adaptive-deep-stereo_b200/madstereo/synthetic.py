"""Synthetic stereo pairs and seeded weights (no datasets / checkpoints are reachable offline).

Left image: smooth multi-octave random texture in [0,255]; right image: the left image resampled
with a known smooth positive disparity field plus N(0,1) noise, so that the reprojection loss of
Losses/loss_factory.py:353-395 is meaningful (well below the 0.5 reset threshold of
Stereo_Online_Adaptation.py:242) and gradients are non-degenerate.  Pure numpy, deterministic.
"""
import numpy as np


def _upsample_bilinear(a, h, w):
    ih, iw = a.shape[:2]
    ys = np.linspace(0, ih - 1, h)
    xs = np.linspace(0, iw - 1, w)
    y0 = np.floor(ys).astype(int); y1 = np.minimum(y0 + 1, ih - 1); fy = (ys - y0)[:, None, None]
    x0 = np.floor(xs).astype(int); x1 = np.minimum(x0 + 1, iw - 1); fx = (xs - x0)[None, :, None]
    t = a[y0][:, x0] * (1 - fx) + a[y0][:, x1] * fx
    b = a[y1][:, x0] * (1 - fx) + a[y1][:, x1] * fx
    return t * (1 - fy) + b * fy


def make_pair(h, w, seed=0, max_disp_frac=0.08, batch=1):
    """Returns (left, right, gt_disp) float32 arrays [B,h,w,3], [B,h,w,3], [B,h,w,1]."""
    lefts, rights, gts = [], [], []
    for b in range(batch):
        rng = np.random.default_rng(seed + 1000 * b)
        tex = np.zeros((h, w, 3))
        amp = 0.0
        for o, a in ((4, 1.0), (16, 0.6), (64, 0.4), (256, 0.25)):
            gh, gw = max(2, h // (256 // o) + 2), max(2, w // (256 // o) + 2)
            tex += a * _upsample_bilinear(rng.uniform(0, 255, (gh, gw, 3)), h, w)
            amp += a
        left = tex / amp
        dgrid = rng.uniform(0.2, 1.0, (max(2, h // 64 + 2), max(2, w // 64 + 2), 1))
        disp = _upsample_bilinear(dgrid, h, w) * max_disp_frac * w
        xs = np.arange(w)[None, :, None] - disp                  # right(x - d) = left(x)  => sample left at x + d
        # build right by forward resampling of left: right[x] = left[x + d(x)] (approximation, smooth d)
        src = np.clip(np.arange(w)[None, :, None] + disp, 0, w - 1)
        x0 = np.floor(src).astype(int); x1 = np.minimum(x0 + 1, w - 1); f = src - x0
        rows = np.arange(h)[:, None]
        right = left[rows, x0[..., 0]] * (1 - f) + left[rows, x1[..., 0]] * f
        right = np.clip(right + rng.normal(0, 1.0, right.shape), 0, 255)
        lefts.append(left); rights.append(right); gts.append(disp)
    f32 = lambda x: np.ascontiguousarray(np.stack(x), dtype=np.float32)
    return f32(lefts), f32(rights), f32(gts)


def make_noise_pair(h, w, seed=0, batch=1):
    """i.i.d. uniform(0,255) images: adversarial numerics for kernel parity."""
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, 255, (batch, h, w, 3)).astype(np.float32),
            rng.uniform(0, 255, (batch, h, w, 3)).astype(np.float32))


def init_params(layers, seed=42, bias_range=0.1):
    """Seeded stand-in for a pretrained checkpoint (none is reachable offline): xavier-uniform weights
    (tf.contrib.layers.xavier_initializer, Nets/sharedLayers.py:4) rescaled so that activations stay O(1) on
    0..255 inputs and every disparity head is alive, small non-zero biases.
    `layers`: iterable of engine LayerInfo (scope, bias_name, kh, kw, cin, cout, transposed).
    """
    rng = np.random.default_rng(seed)
    out = {}
    for l in layers:
        shape = (l.kh, l.kw, l.cout, l.cin) if l.transposed else (l.kh, l.kw, l.cin, l.cout)
        limit = np.sqrt(6.0 / (l.kh * l.kw * (l.cin + l.cout)))
        w = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        head = l.scope.endswith('disp-6') or l.scope.endswith('context-7')
        if l.scope.endswith('gc-read-pyramid/conv1'):
            w *= np.float32(1.0 / 48.0)
        elif head:
            w *= np.float32(0.5)
        else:
            w *= np.float32(1.3)
        b = rng.uniform(-bias_range, bias_range, size=(l.cout,)).astype(np.float32)
        if head:
            b -= np.float32(0.05 if l.scope.endswith('context-7') else 0.35)
        out[l.scope + '/weights'] = w
        out[l.scope + '/' + l.bias_name] = b
    return out
