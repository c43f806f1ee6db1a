"""Eager stand-in for the slice of the TensorFlow 1.x API the reference's hot path uses (TEST INFRASTRUCTURE).

Purpose: execute the reference's OWN Python graph code (Nets/MadNet.py, Nets/DispNet.py, Nets/sharedLayers.py,
Nets/Stereo_net.py, Losses/loss_factory.py, Data_utils/preprocessing.py -- imported unmodified from /root/reference by
oracle/run_reference_graph.py, in the build container only) on torch CPU tensors, so that the wiring of the networks,
the warps, the losses, the train-op variable lists and the variable naming are the reference's own and not a restatement.
TensorFlow itself cannot be installed here; what this module supplies instead is

  * graph-mode bookkeeping restated from TF 1.x behaviour: variable scopes (names are NOT uniquified), name scopes
    (uniquified per parent: re-entering `gc-read-pyramid` gives `gc-read-pyramid_1`), op names `<name scope>/<OpType>[_k]`,
    collections, and `tf.get_collection(scope=...)` as the regex *prefix* match TF implements;
  * kernels: elementwise / reductions / gather / gather_nd / slice / pad / avg_pool as plain torch calls; the three kernels
    whose TF semantics are subtle -- SAME-padded conv2d / atrous_conv2d / conv2d_transpose, legacy `resize_images`
    (align_corners=False) and `resize_image_with_crop_or_pad` -- are shared with oracle/tf1_ops.py.

So vectors produced through this shim pin the oracle to the reference's graph code, NOT to TensorFlow's C++ kernels:
for those three kernel families parity stays unpinned (stated in DESIGN.md section 2).
Gradients come from torch autograd through the executed graph (`tf.stop_gradient` = detach).
"""
import builtins as _bi
import contextlib
import re
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import tf1_ops as T

float32 = 'float32'
int32 = 'int32'
uint8 = 'uint8'
uint16 = 'uint16'
float64 = 'float64'
int64 = 'int64'
_DT = {'float32': torch.float32, 'int32': torch.int32, 'float64': torch.float64, 'int64': torch.int64,
       'uint8': torch.uint8, 'uint16': torch.int32}


# ------------------------------------------------------------------------------------------------
# graph-mode bookkeeping
# ------------------------------------------------------------------------------------------------
class _Graph:
    def __init__(self):
        self.var_scope = []                 # variable-scope components (not uniquified)
        self.name_scope = ''                # current name scope, '' or 'a/b'
        self.used = {}                      # name scope -> {child name: count}
        self.collections = {}
        self.variables = {}                 # full variable name -> TT
        self.params = None                  # dict full-name (without ':0') -> array: values for tf.get_variable
        self.created = []                   # (name, shape) of every variable created, in creation order

    def unique(self, base):
        scope = self.name_scope
        d = self.used.setdefault(scope, {})
        n = d.get(base, 0)
        d[base] = n + 1
        leaf = base if n == 0 else '%s_%d' % (base, n)
        return (scope + '/' + leaf) if scope else leaf


_G = _Graph()


def reset_graph(params=None):
    global _G
    _G = _Graph()
    _G.params = params
    return _G


def graph():
    return _G


class GraphKeys:
    TRAINABLE_VARIABLES = 'trainable_variables'
    WEIGHTS = 'weights'
    GLOBAL_VARIABLES = 'variables'


@contextlib.contextmanager
def name_scope(name):
    old = _G.name_scope
    _G.name_scope = _G.unique(name)
    try:
        yield _G.name_scope
    finally:
        _G.name_scope = old


@contextlib.contextmanager
def variable_scope(name, reuse=None, **_kw):
    _G.var_scope.append(name)
    old = _G.name_scope
    _G.name_scope = _G.unique(name)
    try:
        yield name
    finally:
        _G.name_scope = old
        _G.var_scope.pop()


def get_collection(key, scope=None):
    items = list(_G.collections.get(key, []))
    if scope is None:
        return items
    rx = re.compile(scope)
    return [v for v in items if rx.match(v.name)]       # tf.get_collection: re.match == prefix match


def add_to_collection(key, value):
    _G.collections.setdefault(key, []).append(value)


# ------------------------------------------------------------------------------------------------
# tensor wrapper
# ------------------------------------------------------------------------------------------------
class _Dim:
    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, _Dim) else o)

    def __hash__(self):
        return hash(self.value)

    def __repr__(self):
        return str(self.value)


class _Shape:
    def __init__(self, dims):
        self._d = [int(d) for d in dims]

    def as_list(self):
        return list(self._d)

    def __getitem__(self, i):
        if isinstance(i, _bi.slice):
            return _Shape(self._d[i])
        return _Dim(self._d[i])

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter([_Dim(d) for d in self._d])

    def __str__(self):
        return '(' + ', '.join(str(d) for d in self._d) + (',)' if len(self._d) == 1 else ')')

    __repr__ = __str__


def _raw(x):
    if isinstance(x, TT):
        return x.t
    if isinstance(x, _Dim):
        return x.value
    return x


def _wrap(t, op):
    return TT(t, _G.unique(op))


class TT:
    """A 'tensor': torch value + TF-style op name."""
    __array_priority__ = 1000

    def __init__(self, t, name):
        self.t = t
        self.name = name

    # -- shape API used by the reference
    def get_shape(self):
        return _Shape(self.t.shape)

    @property
    def shape(self):
        return _Shape(self.t.shape)

    def set_shape(self, _s):
        return None

    @property
    def dtype(self):
        return self.t.dtype

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __int__(self):
        return int(self.t)

    __index__ = __int__

    def __float__(self):
        return float(self.t)

    def __len__(self):
        return self.t.shape[0]

    def __iter__(self):
        return iter([self[i] for i in _bi.range(self.t.shape[0])])

    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            idx = tuple(_bi.slice(_ival(s.start), _ival(s.stop), _ival(s.step)) if isinstance(s, _bi.slice) else _ival(s) for s in idx)
        elif isinstance(idx, _bi.slice):
            idx = _bi.slice(_ival(idx.start), _ival(idx.stop), _ival(idx.step))
        else:
            idx = _ival(idx)
        return _wrap(self.t[idx], 'strided_slice')

    # -- arithmetic
    def _bin(self, o, f, op, rev=False):
        a, b = self.t, _raw(o)
        if isinstance(b, (list, tuple, np.ndarray)):
            b = torch.as_tensor(np.asarray(b), dtype=a.dtype)
        if isinstance(b, float) and not a.dtype.is_floating_point:
            a = a.to(torch.float32)
        return _wrap(f(b, a) if rev else f(a, b), op)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, 'add')
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, 'add', True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, 'sub')
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, 'sub', True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, 'mul')
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, 'mul', True)
    def __truediv__(self, o): return self._bin(o, _div, 'truediv')
    def __rtruediv__(self, o): return self._bin(o, _div, 'truediv', True)
    def __floordiv__(self, o): return self._bin(o, _floordiv, 'floordiv')
    def __rfloordiv__(self, o): return self._bin(o, _floordiv, 'floordiv', True)
    def __mod__(self, o): return self._bin(o, lambda a, b: a % b, 'mod')
    def __pow__(self, o): return self._bin(o, lambda a, b: a ** b, 'pow')
    def __neg__(self): return _wrap(-self.t, 'Neg')
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b, 'Greater')
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b, 'Less')


def _ival(v):
    if v is None:
        return None
    if isinstance(v, TT):
        return int(v.t)
    if isinstance(v, _Dim):
        return int(v.value)
    return v


def _div(a, b):
    a = a if torch.is_tensor(a) else torch.as_tensor(a)
    b = b if torch.is_tensor(b) else torch.as_tensor(b)
    if not a.dtype.is_floating_point and not b.dtype.is_floating_point:
        a = a.to(torch.float64)         # python3 '/' on int tensors: true division (TF: float64)
    return a / b


def _floordiv(a, b):
    return torch.div(a if torch.is_tensor(a) else torch.as_tensor(a), b, rounding_mode='floor')


def _t(x, like=None):
    """torch value of a TT / number / list (lists may contain TTs)."""
    if isinstance(x, TT):
        return x.t
    if isinstance(x, (list, tuple)):
        return torch.stack([_t(v).to(torch.float32) if not isinstance(v, (int, np.integer)) else torch.tensor(v) for v in x])
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x)
    return torch.as_tensor(x)


# ------------------------------------------------------------------------------------------------
# variables
# ------------------------------------------------------------------------------------------------
def constant_initializer(value=0.0):
    return ('constant', value)


def get_variable(name, shape, initializer=None, **_kw):
    full = '/'.join(_G.var_scope + [name])
    key = full + ':0'
    if key in _G.variables:
        return _G.variables[key]
    shape = [int(_ival(s)) for s in (shape if isinstance(shape, (list, tuple)) else [shape])]
    if _G.params is not None and full in _G.params:
        val = torch.as_tensor(np.asarray(_G.params[full]), dtype=torch.float32).clone()
        assert list(val.shape) == shape, (full, list(val.shape), shape)
    else:
        raise KeyError('tf1_shim: no value supplied for variable %s %s' % (full, shape))
    val.requires_grad_(True)
    v = TT(val, key)
    _G.variables[key] = v
    _G.created.append((full, tuple(shape)))
    add_to_collection(GraphKeys.TRAINABLE_VARIABLES, v)
    add_to_collection(GraphKeys.GLOBAL_VARIABLES, v)
    return v


# ------------------------------------------------------------------------------------------------
# ops
# ------------------------------------------------------------------------------------------------
def shape(x):
    return TT(torch.tensor(list(_t(x).shape), dtype=torch.int32), _G.unique('Shape'))


def cast(x, dtype=None, **kw):
    dtype = kw.get('dtype', dtype)
    dt = _DT[dtype] if isinstance(dtype, str) else dtype
    t = _t(x)
    if not dt.is_floating_point and t.dtype.is_floating_point:
        t = torch.trunc(t)              # tf.cast float -> int truncates toward zero
    return _wrap(t.to(dt), 'Cast')


def to_int32(x):
    return cast(x, 'int32')


def constant(value, dtype=None, **_kw):
    t = torch.as_tensor(np.asarray(_raw(value)))
    if dtype is not None:
        t = t.to(_DT[dtype] if isinstance(dtype, str) else dtype)
    elif t.dtype == torch.float64:
        t = t.to(torch.float32)
    return _wrap(t, 'Const')


def zeros(shape_, dtype='float32'):
    return _wrap(torch.zeros([int(_ival(s)) for s in shape_], dtype=_DT[dtype] if isinstance(dtype, str) else dtype), 'zeros')


def ones(shape=None, dtype='float32', **_kw):
    s = _t(shape) if isinstance(shape, TT) else shape
    dims = [int(v) for v in (s.tolist() if torch.is_tensor(s) else [_ival(v) for v in s])]
    return _wrap(torch.ones(dims, dtype=_DT[dtype] if isinstance(dtype, str) else dtype), 'ones')


def zeros_like(x, dtype=None):
    t = torch.zeros_like(_t(x))
    return _wrap(t.to(_DT[dtype]) if isinstance(dtype, str) else t, 'zeros_like')


def ones_like(x, dtype=None):
    t = torch.ones_like(_t(x))
    return _wrap(t.to(_DT[dtype]) if isinstance(dtype, str) else t, 'ones_like')


def maximum(a, b): return _wrap(torch.maximum(_t(a), _t(b)), 'Maximum')
def abs(x): return _wrap(torch.abs(_t(x)), 'Abs')                    # noqa: A001
def square(x): return _wrap(_t(x) ** 2, 'Square')
def sqrt(x): return _wrap(torch.sqrt(_t(x)), 'Sqrt')
def floor(x): return _wrap(torch.floor(_t(x)), 'Floor')
def equal(a, b): return _wrap(_t(a) == _t(b), 'Equal')
def greater(a, b): return _wrap(_t(a) > _t(b), 'Greater')
def stop_gradient(x): return _wrap(_t(x).detach(), 'StopGradient')
def identity(x): return _wrap(_t(x), 'Identity')


def clip_by_value(x, lo, hi):
    t = _t(x)
    return _wrap(torch.minimum(torch.maximum(t, _t(lo).to(t.dtype)), _t(hi).to(t.dtype)), 'clip_by_value')


def where(cond, x=None, y=None):
    return _wrap(torch.where(_t(cond), _t(x), _t(y)), 'Select')


def floordiv(a, b): return _wrap(_floordiv(_t(a), _t(b)), 'floordiv')


def concat(values, axis=-1, **_kw):
    ts = [_t(v) if not isinstance(v, np.ndarray) else torch.as_tensor(v) for v in values]
    dt = torch.result_type(ts[0], ts[1]) if len(ts) > 1 else ts[0].dtype
    if any(t.dtype == torch.float64 for t in ts) and any(t.dtype == torch.float32 for t in ts):
        dt = torch.float32              # numpy constants mixed with float32 tensors (MadNet._build_indeces)
    return _wrap(torch.cat([t.to(dt) for t in ts], dim=axis), 'concat')


def split(x, sizes, axis=0, **_kw):
    parts = torch.split(_t(x), [int(s) for s in sizes] if isinstance(sizes, (list, tuple)) else int(sizes), dim=axis)
    return [_wrap(p, 'split') for p in parts]


def stack(values, axis=0):
    return _wrap(torch.stack([_t(v) if isinstance(v, TT) else torch.as_tensor(_ival(v)) for v in values], dim=axis), 'stack')


def reshape(x, shape_):
    s = _t(shape_) if isinstance(shape_, TT) else shape_
    dims = [int(v) for v in (s.tolist() if torch.is_tensor(s) else [_ival(v) for v in s])]
    return _wrap(_t(x).reshape(dims), 'Reshape')


def transpose(x, perm=None): return _wrap(_t(x).permute(*perm), 'transpose')
def expand_dims(x, axis): return _wrap(_t(x).unsqueeze(axis), 'ExpandDims')
def tile(x, multiples): return _wrap(_t(x).repeat(*[int(_ival(m)) for m in multiples]), 'Tile')
def matmul(a, b): return _wrap(_t(a) @ _t(b), 'MatMul')
def add_n(values): return _wrap(sum(_t(v) for v in values), 'AddN')


def range(*args, **kw):                                              # noqa: A001
    dt = kw.get('dtype')
    vals = [float(_t(a)) if isinstance(a, TT) else a for a in args]
    delta = kw.get('delta')
    if delta is not None:
        vals = [0, vals[0], float(_t(delta))]
    t = torch.arange(*vals)
    if dt is not None:
        t = t.to(_DT[dt] if isinstance(dt, str) else dt)
    elif len(args) == 1 and isinstance(args[0], TT):
        t = t.to(args[0].t.dtype)
    return _wrap(t, 'range')


def slice(x, begin, size):                                           # noqa: A001
    t = _t(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        b, s = int(_ival(b)), int(_ival(s))
        idx.append(_bi.slice(b, None if s == -1 else b + s))
    return _wrap(t[tuple(idx)], 'Slice')


def pad(x, paddings, mode='CONSTANT', **_kw):
    t = _t(x)
    p = [[int(_ival(a)), int(_ival(b))] for a, b in paddings]
    assert t.dim() == 4 and p[0] == [0, 0] and p[3] == [0, 0]
    nchw = t.permute(0, 3, 1, 2)
    out = F.pad(nchw, (p[2][0], p[2][1], p[1][0], p[1][1]), mode='reflect' if mode.upper() == 'REFLECT' else 'constant')
    return _wrap(out.permute(0, 2, 3, 1), 'Pad' if mode.upper() == 'CONSTANT' else 'MirrorPad')


def _reduce(f, name):
    def op(x, axis=None, keepdims=False, **kw):
        keepdims = kw.get('keep_dims', keepdims)
        if isinstance(x, (list, tuple)):                             # tf.reduce_sum(list of scalars)
            t = torch.stack([_t(v) for v in x])
        else:
            t = _t(x)
        if axis is None:
            return _wrap(f(t), name)
        return _wrap(f(t, dim=axis, keepdim=keepdims), name)
    return op


reduce_mean = _reduce(torch.mean, 'Mean')
reduce_sum = _reduce(torch.sum, 'Sum')


def gather(params, indices):
    return _wrap(_t(params)[_t(indices).long()], 'Gather')


def gather_nd(params, indices):
    p, i = _t(params), _t(indices).long()
    return _wrap(p[tuple(i[..., k] for k in _bi.range(i.shape[-1]))], 'GatherNd')


def Print(x, *_a, **_k):                                             # noqa: N802
    return x


def placeholder(dtype, shape=None, name=None):
    raise NotImplementedError('tf1_shim: placeholders are not used on the sequence path')


# ---- namespaces -----------------------------------------------------------------------------------
def _conv2d(x, w, strides=None, padding='SAME', **_kw):
    assert padding == 'SAME' and strides[0] == 1 and strides[3] == 1 and strides[1] == strides[2]
    return _wrap(T.conv2d(_t(x), _t(w), None, stride=int(strides[1]), dilation=1, alpha=None), 'Conv2D')


def _atrous_conv2d(x, w, rate=1, padding='SAME', **_kw):
    assert padding == 'SAME'
    return _wrap(T.conv2d(_t(x), _t(w), None, stride=1, dilation=int(rate), alpha=None), 'convolution')


def _conv2d_transpose(x, w, output_shape, strides=None, padding='SAME', **_kw):
    assert padding == 'SAME' and strides[1] == strides[2]
    out = T.conv2d_transpose(_t(x), _t(w), torch.zeros(_t(w).shape[2]), stride=int(strides[1]), alpha=None)
    want = [int(_ival(v)) for v in output_shape]
    assert list(out.shape) == want, (list(out.shape), want)
    return _wrap(out, 'conv2d_transpose')


def _bias_add(x, b): return _wrap(_t(x) + _t(b), 'BiasAdd')
def _relu(x): return _wrap(torch.relu(_t(x)), 'Relu')
def _leaky_relu(x, alpha=0.2): return _wrap(torch.maximum(alpha * _t(x), _t(x)), 'LeakyRelu')


def _avg_pool(x, ksize, strides, padding='VALID', **_kw):
    assert padding == 'VALID' and list(strides) == [1, 1, 1, 1]
    out = F.avg_pool2d(_t(x).permute(0, 3, 1, 2), (int(ksize[1]), int(ksize[2])), stride=1)
    return _wrap(out.permute(0, 2, 3, 1), 'AvgPool')


nn = types.SimpleNamespace(conv2d=_conv2d, atrous_conv2d=_atrous_conv2d, conv2d_transpose=_conv2d_transpose,
                           bias_add=_bias_add, relu=_relu, leaky_relu=_leaky_relu, avg_pool=_avg_pool)


class _ResizeMethod:
    BILINEAR = 0


def _resize_images(x, size, method=0, align_corners=False, **_kw):
    assert method == 0 and not align_corners
    s = _t(size) if isinstance(size, TT) else size
    oh, ow = [int(v) for v in (s.tolist() if torch.is_tensor(s) else [_ival(v) for v in s])]
    return _wrap(T.resize_bilinear(_t(x), oh, ow), 'ResizeBilinear')


def _crop_or_pad(x, th, tw):
    return _wrap(T.crop_or_pad(_t(x), int(_ival(th)), int(_ival(tw))), 'resize_image_with_crop_or_pad')


image = types.SimpleNamespace(resize_images=_resize_images, resize_bilinear=_resize_images, ResizeMethod=_ResizeMethod,
                              resize_image_with_crop_or_pad=_crop_or_pad)
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=lambda *a, **k: ('xavier',)))
summary = types.SimpleNamespace(scalar=lambda *a, **k: None, image=lambda *a, **k: None)
layers = types.SimpleNamespace()


def as_module():
    """A module object exposing this file's public names, to be installed as sys.modules['tensorflow']."""
    import sys
    me = sys.modules[__name__]
    m = types.ModuleType('tensorflow')
    for k in dir(me):
        if not k.startswith('_'):
            setattr(m, k, getattr(me, k))
    m.__version__ = '1.12.0-shim'
    return m
