"""Host-side data-parallel logic of OnlineAdaptation on CPU: world_size 2, gloo, a fake engine.

Checks what SURVEY §8(e) requires of the N>1 path: every rank adapts the SAME module (rank 0 samples, broadcast),
the module's gradient range is summed across ranks and scaled by 1/N in the update, the loss used for the reward /
reset decision is the global mean, and the replicas stay bit-identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class FakeEngine:
    """Implements the engine surface OnlineAdaptation uses, on CPU tensors."""

    def __init__(self, rank):
        self.rank = rank
        self.device = torch.device('cpu')
        self.layers = list(range(6))

    def set_groups(self, groups):
        self.groups = groups

    def bind(self):
        self.n_params = 40
        self.weights = torch.arange(40, dtype=torch.float32) * 0.01
        self.grads = torch.zeros(40)
        self.momentum = torch.zeros(40)
        self.group_ranges = [(0, 8), (8, 16), (16, 24), (24, 32), (32, 40)]
        self.calls = []

    def load_params(self, p):
        pass

    def weights_changed(self):
        pass

    def set_input(self, l, r):
        self.frame = float(l)

    def run(self, mode, group=0, disp_mask=0, with_update=True, lr=1e-4, mu=0.9, grad_scale=1.0):
        assert not with_update                       # N>1: the update must come after the all-reduce
        lo, hi = self.group_ranges[group]
        self.grads[lo:hi] = (self.rank + 1) * (group + 1) * torch.ones(hi - lo)
        self.calls.append(('run', group))

    def update(self, group, lr, mu=0.9, grad_scale=1.0):
        lo, hi = self.group_ranges[group]
        self.momentum[lo:hi] = mu * self.momentum[lo:hi] + self.grads[lo:hi] * grad_scale
        self.weights[lo:hi] -= lr * self.momentum[lo:hi]
        self.calls.append(('update', group, grad_scale))

    def read_scalars(self):
        return [0.1 * (self.rank + 1) + 0.01 * self.frame, 0.0, 0.0, 0.0]


class FakeVar:
    def __init__(self, i):
        self.name = 'v%d:0' % i
        self.idx = i


class FakeNet:
    def __init__(self, rank):
        self.engine = FakeEngine(rank)
        self.bulkhead = True
        self._vars = {('g%d' % i): [FakeVar(i)] for i in range(5)}

    def get_disparities(self):
        return list(range(6))

    def get_variables(self, name):
        return self._vars[name]

    def layer_index_of_variable(self, v):
        return v.idx


class FakePeerEngine(FakeEngine):
    """The fused path: run(with_update=2) = backward + in-graph peer-memory all-reduce + update (csrc/dp.cu); here the
    exchange is emulated with gloo so that the host logic around it (sampler stream, loss handling) runs on CPU."""

    def dp_setup(self, rank, world, pg=None):
        self.world = world

    def dp_error(self):
        return 0

    def run(self, mode, group=0, disp_mask=0, with_update=True, lr=1e-4, mu=0.9, grad_scale=1.0):
        assert int(with_update) == 2
        lo, hi = self.group_ranges[group]
        self.grads[lo:hi] = (self.rank + 1) * (group + 1) * torch.ones(hi - lo)
        dist.all_reduce(self.grads[lo:hi])
        self.momentum[lo:hi] = mu * self.momentum[lo:hi] + self.grads[lo:hi] / self.world
        self.weights[lo:hi] -= lr * self.momentum[lo:hi]
        self.calls.append(('update', group, 1.0 / self.world))
        t = torch.tensor([0.1 * (self.rank + 1) + 0.01 * self.frame], dtype=torch.float64)
        dist.all_reduce(t)
        self._mean_loss = float(t) / self.world

    def update(self, *a, **k):
        raise AssertionError('fused path must not call update()')

    def read_scalars(self):
        return [self._mean_loss, 0.0, 0.0, 0.0]


def _worker(rank, world, port, out, peer=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    from conftest import PKG  # noqa: F401  (adds the package dir to sys.path)
    from madstereo.adaptation import OnlineAdaptation
    np.random.seed(100 + rank)                        # different host RNG per rank: sampling must still agree
    net = FakeNet(rank)
    if peer:
        net.engine = FakePeerEngine(rank)
    ad = OnlineAdaptation(net, mode='MAD', train_config=[['g%d' % i] for i in range(5)], lr=0.5,
                          sample_mode='PROBABILITY', num_blocks=1, ssim_th=10.0)
    hist = []
    for t in range(6):
        o = ad.step(float(t), float(t))
        hist.append((o['blocks'], o['loss']))
    out.put((rank, hist, net.engine.weights.clone().numpy(), [c for c in net.engine.calls if c[0] == 'update'],
             ad.sample_distribution.copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('peer', [False, True])
def test_dp_two_ranks_gloo(peer):
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, peer)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, h0, w0, u0, d0), (_, h1, w1, u1, d1) = res
    assert [b for b, _ in h0] == [b for b, _ in h1]                     # same module on every rank, every frame
    for (_, l0), (_, l1) in zip(h0, h1):
        assert abs(l0 - l1) < 1e-12                                    # the globally averaged loss
    assert abs(h0[0][1] - (0.1 * 1.5 + 0.0)) < 1e-6                    # mean of the two ranks' losses at frame 0
    assert np.array_equal(w0, w1)                                       # replicas stay identical
    assert all(c[2] == 0.5 for c in u0)                                 # 1/N folded into the update
    assert np.allclose(d0, d1)
