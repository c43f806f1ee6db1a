"""Parity metrics shared by the GPU tests, bench.py's self-check and __graft_entry__.smoke() (TEST INFRASTRUCTURE).

`rel_linf` is the north-star disparity metric (relative L-inf).  For gradients a single leaky-ReLU / relu / floor()
mask that lands on the other side of its kink under a different fp32 summation order perturbs one pixel's
contribution to a weight gradient; L-inf of the whole tensor then moves by that pixel's share even though every
backward term is right.  `grad_report` therefore measures every trained tensor three ways against the fp64 oracle:

  * rel_l2      = |g_gpu - g64|_2 / |g64|_2                          -- what a kink flip barely moves
  * noise_l2    = |g32  - g64|_2 / |g64|_2                           -- the fp32 oracle's own distance to fp64
  * rel_linf    = |g_gpu - g64|_inf / |g64|_inf                      -- secondary bound

and `assert_grads` requires rel_l2 <= max(tol_l2, 2 * noise_l2) and rel_linf <= max(tol_linf, 2 * noise_linf): the GPU
may not be further from the fp64 truth than twice what plain fp32 CPU arithmetic already is (or the stated floor).
"""
import numpy as np


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def grad_report(got, g32, g64):
    """dicts name -> array.  Returns {name: (rel_l2, noise_l2, rel_linf, noise_linf)} over the names of g64."""
    rep = {}
    for n, t in g64.items():
        rep[n] = (rel_l2(got[n], t), rel_l2(g32[n], t), rel_linf(got[n], t), rel_linf(g32[n], t))
    return rep


def assert_grads(rep, tol_l2, tol_linf, what=''):
    """Per tensor: error <= max(floor, 2 x that tensor's fp32-oracle noise, the step's noise level).  The step's noise level
    (the largest fp32-vs-fp64 distance over the trained tensors) covers tensors with a handful of elements -- e.g. the single
    bias of a disparity head -- whose own noise figure is ONE sample of the kink-flip distribution that moves every tensor
    of the step, so twice that sample is not a bound."""
    worst = (0.0, None)
    n2_step = max(v[1] for v in rep.values()); ni_step = max(v[3] for v in rep.values())
    for n, (l2, n2, li, ni) in rep.items():
        assert l2 <= max(tol_l2, 2.0 * n2, n2_step), '%s %s: rel L2 %.3e (fp32-oracle noise %.3e, step noise %.3e, floor %.1e)' % (what, n, l2, n2, n2_step, tol_l2)
        # L-inf is a single element: its bound is twice the tensor's or the step's L-inf noise (the visit-h run of this
        # suite showed a +/-4 % margin deciding pass / fail when only the summation order of small layers changed)
        assert li <= max(tol_linf, 2.0 * ni, 2.0 * ni_step), '%s %s: rel Linf %.3e (fp32-oracle noise %.3e, step noise %.3e, floor %.1e)' % (what, n, li, ni, ni_step, tol_linf)
        if l2 > worst[0]:
            worst = (l2, n)
    return worst


def summarize(rep):
    l2 = max(v[0] for v in rep.values()); n2 = max(v[1] for v in rep.values())
    li = max(v[2] for v in rep.values()); ni = max(v[3] for v in rep.values())
    return {'max_rel_l2': l2, 'max_noise_l2': n2, 'max_rel_linf': li, 'max_noise_linf': ni}
