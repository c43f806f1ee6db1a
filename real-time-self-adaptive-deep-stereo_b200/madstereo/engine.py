"""Python handle on the C++ engine (csrc/engine.cu) — torch CUDA tensors are memory containers only.

The engine replaces the single `sess.run(tf_fetches)` of the reference inner loop
(Stereo_Online_Adaptation.py:194-208): forward of the whole net, full-resolution reprojection loss, and the
selected train op (MAD module or FULL) with its momentum update, all as hand-written CUDA.
"""
import ctypes
from collections import OrderedDict, namedtuple
from ctypes import byref, c_float, c_int, c_size_t, c_void_p

import numpy as np
import torch

from ._lib import MadStereoError, check, lib

LayerInfo = namedtuple('LayerInfo', 'index name scope bias_name kh kw cin cout stride dilation transposed alpha')

MODE_NONE, MODE_MAD, MODE_FULL = 0, 1, 2


class _NullCtx(object):
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL_CTX = _NullCtx()


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


class StereoEngine:
    def __init__(self, net_name, batch, height, width, radius_d=2, stride=1, warping=True, device=None):
        if not torch.cuda.is_available():
            raise MadStereoError('a CUDA device is required: libmadstereo has no CPU fallback')
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self._lib = lib()
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.net_name, self.B, self.H, self.W = net_name, batch, height, width
        with self._on_device():
            self._h = self._lib.ms_engine_create(net_name.encode(), batch, height, width, radius_d, stride,
                                                 1 if warping else 0)
        if not self._h:
            raise MadStereoError('ms_engine_create: %s' % self._lib.ms_last_error().decode())
        self.layers = []
        n = self._lib.ms_engine_num_layers(self._h)
        for i in range(n):
            nm, sc, bn = (ctypes.create_string_buffer(128) for _ in range(3))
            dims = (c_int * 7)()
            alpha = c_float()
            check(self._lib.ms_engine_layer_info(self._h, i, nm, 128, sc, 128, bn, 128, dims, byref(alpha)), 'layer_info')
            self.layers.append(LayerInfo(i, nm.value.decode(), sc.value.decode(), bn.value.decode(), *list(dims),
                                         alpha.value))
        self.layer_by_name = {l.name: l for l in self.layers}
        self.n_groups = 0
        self.bound = False
        self._keep = []

    def _on_device(self):
        """Context that makes this engine's GPU current -- a no-op object when it already is (the per-frame calls of a
        one-process-per-GPU run: entering torch.cuda.device() costs ~10 us of host time on the frame's critical path)."""
        if torch.cuda.current_device() == self._dev_index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                self._lib.ms_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- construction ----------------------------------------------------------------------
    def set_groups(self, groups):
        """groups: list (one per MAD module) of lists of layer indices. Must be called before bind()."""
        arr = (c_int * len(self.layers))(*([-1] * len(self.layers)))
        for g, idxs in enumerate(groups):
            for i in idxs:
                if arr[i] != -1 and arr[i] != g:
                    raise MadStereoError('layer %s appears in two MAD groups: unsupported' % self.layers[i].name)
                arr[i] = g
        check(self._lib.ms_engine_set_groups(self._h, arr, len(self.layers), len(groups)), 'set_groups')
        self.n_groups = len(groups)

    def bind(self):
        npar, nws = c_size_t(), c_size_t()
        check(self._lib.ms_engine_sizes(self._h, byref(npar), byref(nws)), 'sizes')
        self.n_params = npar.value
        dev = self.device
        self.weights = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.momentum = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
        self.workspace = torch.empty(nws.value, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.ms_engine_bind(self._h, _ptr(self.weights), _ptr(self.grads), _ptr(self.momentum),
                                           _ptr(self.workspace), nws.value, _stream()), 'bind')
        self.bound = True
        self.param_offsets = {}
        for l in self.layers:
            wo, bo = c_size_t(), c_size_t()
            check(self._lib.ms_engine_param_offsets(self._h, l.index, byref(wo), byref(bo)), 'param_offsets')
            self.param_offsets[l.index] = (wo.value, bo.value)
        self.group_ranges = []
        for g in range(self.n_groups):
            b, e = c_size_t(), c_size_t()
            check(self._lib.ms_engine_group_range(self._h, g, byref(b), byref(e)), 'group_range')
            self.group_ranges.append((b.value, e.value))

    def weight_shape(self, l):
        return (l.kh, l.kw, l.cout, l.cin) if l.transposed else (l.kh, l.kw, l.cin, l.cout)

    def param_views(self, arena=None):
        """OrderedDict TF-variable-name -> torch view (HWIO weights / [cout] bias) into an arena."""
        arena = self.weights if arena is None else arena
        out = OrderedDict()
        for l in self.layers:
            wo, bo = self.param_offsets[l.index]
            shp = self.weight_shape(l)
            out[l.scope + '/weights'] = arena[wo:wo + int(np.prod(shp))].view(*shp)
            out[l.scope + '/' + l.bias_name] = arena[bo:bo + l.cout]
        return out

    def load_params(self, params, strict=True):
        """strict=False: parameters absent from `params` keep their current values (tf.train.Saver with a partial
        var_list, Data_utils/weights_utils.py:27-37)."""
        views = self.param_views()
        for k, v in views.items():
            if k not in params:
                if strict:
                    raise MadStereoError('missing parameter %s' % k)
                continue
            v.copy_(torch.as_tensor(np.asarray(params[k]), dtype=torch.float32).to(self.device).reshape(v.shape))
        self.weights_changed()

    def export_params(self, arena=None):
        return OrderedDict((k, v.detach().cpu().numpy().copy()) for k, v in self.param_views(arena).items())

    # ---- per-frame calls ---------------------------------------------------------------------
    def set_input(self, left, right):
        """left/right: [B,H,W,3] float32 torch tensors (CUDA, or pinned/pageable CPU) or numpy arrays."""
        if self._is_u8(left) and self._is_u8(right):
            l, r = self._as_u8(left), self._as_u8(right)
            self._keep = [l, r]
            with self._on_device():
                check(self._lib.ms_engine_set_input_u8(self._h, _ptr(l), _ptr(r), _stream()), 'set_input_u8')
            return
        l, r = self._as_f32(left), self._as_f32(right)
        self._keep = [l, r]
        with self._on_device():
            check(self._lib.ms_engine_set_input(self._h, _ptr(l), _ptr(r), _stream()), 'set_input')

    @staticmethod
    def _is_u8(x):
        return (isinstance(x, np.ndarray) and x.dtype == np.uint8) or (torch.is_tensor(x) and x.dtype == torch.uint8)

    def _as_u8(self, x, c=3):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if not x.is_contiguous():
            x = x.contiguous()
        if tuple(x.shape) != (self.B, self.H, self.W, c):
            raise MadStereoError('input shape %s != %s' % (tuple(x.shape), (self.B, self.H, self.W, c)))
        return x

    def set_gt(self, gt):
        g = self._as_f32(gt, 1)
        self._keep_gt = g
        with self._on_device():
            check(self._lib.ms_engine_set_gt(self._h, _ptr(g), _stream()), 'set_gt')

    def set_proxy(self, proxy):
        """Proxy disparities [B,H,W,1] of the continual-adaptation variant (Stereo_Continual_Adaptation.py:54)."""
        g = self._as_f32(proxy, 1)
        self._keep_proxy = g
        with self._on_device():
            check(self._lib.ms_engine_set_proxy(self._h, _ptr(g), _stream()), 'set_proxy')

    def set_loss(self, kind, weight_full=0.01, weight_module=0.1):
        """kind: 'reprojection' (SSIM + L1, Stereo_Online_Adaptation.py) or 'proxy' (masked L1 to proxy labels)."""
        k = {'reprojection': 0, 'proxy': 1}[kind]
        check(self._lib.ms_engine_set_loss(self._h, k, weight_full, weight_module), 'set_loss')

    def _as_f32(self, x, c=3):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.to(torch.float32).contiguous()
        if tuple(x.shape) != (self.B, self.H, self.W, c):
            raise MadStereoError('input shape %s != %s' % (tuple(x.shape), (self.B, self.H, self.W, c)))
        return x

    def forward(self, disp_mask=0b100000):
        with self._on_device():
            check(self._lib.ms_engine_forward(self._h, disp_mask, _stream()), 'forward')

    def loss(self, which, with_grad, slot, grad_scale=1.0):
        with self._on_device():
            check(self._lib.ms_engine_loss(self._h, which, 1 if with_grad else 0, slot, grad_scale, _stream()), 'loss')

    def backward(self, mode, group=0):
        with self._on_device():
            check(self._lib.ms_engine_backward(self._h, mode, group, _stream()), 'backward')

    def update(self, group, lr, mu=0.9, grad_scale=1.0):
        with self._on_device():
            check(self._lib.ms_engine_update(self._h, group, lr, mu, grad_scale, _stream()), 'update')

    def run(self, mode, group=0, disp_mask=0, with_update=True, lr=1e-4, mu=0.9, grad_scale=1.0):
        """One whole frame (forward, full-res loss, train op) — replayed as a single CUDA graph."""
        with self._on_device():
            check(self._lib.ms_engine_run(self._h, mode, group, disp_mask, int(with_update), lr, mu, grad_scale,
                                          _stream()), 'run')

    # ---- data parallel: gradient exchange over NVLink peer memory fused with the update (csrc/dp.cu) ----------
    def dp_setup(self, rank, world, process_group=None):
        """Allocate this rank's exchange buffer, all-gather the CUDA IPC handles through torch.distributed (plumbing
        only) and map every peer.  Afterwards run(..., with_update=2) performs all-reduce + momentum update in-graph."""
        import torch.distributed as dist
        mine = (ctypes.c_ubyte * 128)()
        with self._on_device():
            check(self._lib.ms_engine_dp_create(self._h, rank, world, mine), 'dp_create')
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(mine), group=process_group)
        blob = (ctypes.c_ubyte * (128 * world)).from_buffer_copy(b''.join(gathered))
        with self._on_device():
            check(self._lib.ms_engine_dp_connect(self._h, blob), 'dp_connect')
        dist.barrier(group=process_group)

    def dp_error(self):
        out = ctypes.c_uint(0)
        check(self._lib.ms_engine_dp_error(self._h, byref(out)), 'dp_error')
        return int(out.value)

    def weights_changed(self):
        """Call after writing the weight arena from outside (checkpoint load / reset)."""
        check(self._lib.ms_engine_weights_changed(self._h), 'weights_changed')

    def metrics(self):
        with self._on_device():
            check(self._lib.ms_engine_metrics(self._h, _stream()), 'metrics')

    def read_scalars(self):
        out = (c_float * 4)()
        with self._on_device():
            check(self._lib.ms_engine_read_scalars(self._h, out, _stream()), 'read_scalars')
        return list(out)

    # ---- profiling -----------------------------------------------------------------------------
    CATEGORIES = ('conv_fwd', 'conv_dgrad', 'conv_wgrad', 'corr_fwd', 'corr_bwd', 'loss', 'other')

    def profile(self, enable):
        """False/0 off; True/1 eager events; 2 = event nodes inside the replayed CUDA graph (in-graph kernel times)."""
        check(self._lib.ms_engine_profile(self._h, int(enable)), 'profile')

    def profile_event_overhead_us(self):
        return 1e3 * float(self._lib.ms_engine_profile_event_overhead_ms(self._h))

    def profile_layers(self):
        n = len(self.layers)
        ms = (ctypes.c_double * (3 * n))(); calls = (ctypes.c_longlong * (3 * n))()
        with self._on_device():
            check(self._lib.ms_engine_profile_layers(self._h, ms, calls), 'profile_layers')
        out = {}
        for d, name in enumerate(('fwd', 'dgrad', 'wgrad')):
            for i, l in enumerate(self.layers):
                if calls[d * n + i]:
                    out.setdefault(l.name, {})[name] = {'ms': ms[d * n + i], 'calls': calls[d * n + i]}
        return out

    def profile_read(self):
        ms, macs, byts = ((ctypes.c_double * 7)() for _ in range(3))
        calls = (ctypes.c_longlong * 7)()
        with self._on_device():
            check(self._lib.ms_engine_profile_read(self._h, ms, macs, byts, calls), 'profile_read')
        return {c: {'ms': ms[i], 'macs': macs[i], 'bytes': byts[i], 'calls': calls[i]}
                for i, c in enumerate(self.CATEGORIES)}

    def launch_count(self):
        return int(self._lib.ms_launch_count())

    # ---- introspection -------------------------------------------------------------------------
    def tensor_names(self):
        n = self._lib.ms_engine_num_tensors(self._h)
        out = []
        for i in range(n):
            b = ctypes.create_string_buffer(128)
            check(self._lib.ms_engine_tensor_name(self._h, i, b, 128), 'tensor_name')
            out.append(b.value.decode())
        return out

    def tensor_info(self, name):
        p = c_void_p()
        dims = (c_int * 5)()
        check(self._lib.ms_engine_tensor(self._h, name.encode(), byref(p), dims), 'tensor')
        return p.value, tuple(dims)

    def tensor(self, name):
        """torch view [n,h,w,c] (strided) of an engine tensor inside the workspace."""
        p, (n, h, w, c, cs) = self.tensor_info(name)
        off = (p - self.workspace.data_ptr()) // 4
        return self.workspace.as_strided((n, h, w, c), (h * w * cs, w * cs, cs, 1), off)
