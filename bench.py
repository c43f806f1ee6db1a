#!/usr/bin/env python
"""bench.py — stereo FPS (forward + MAD back-prop + momentum update) @1280x384 on N B200s.

One "step" = one pass of the hot path over one stereo pair per GPU: what a single sess.run(fetches) does in
the reference inner loop (Stereo_Online_Adaptation.py:194-208) in MAD mode with the deterministic SEQUENTIAL
sampler (uniform 1/5 module mix): forward of the whole MADNet, full-resolution reprojection loss, module loss,
module backward, NCCL all-reduce of the module's gradient range (N>1), momentum update, loss read-back.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA, libmadstereo)
  python bench.py --impl reference ...                           # the reference path on host cores (CPU oracle)

Under torchrun (N>1) every rank runs one frame per step (weak scaling); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W = 384, 1280
METRIC = 'stereo FPS (fwd+MAD backprop) @1280x384'
L2_MB = 126


WORKLOAD = ('MADNet MAD adaptation (block_config/MadNet_full.json, SEQUENTIAL sampler => uniform mix of the 5 modules), '
            '1280x384, 1 frame per GPU per step')


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        d = json.load(open(path))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'src': 'fallback'}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) >= 7:
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def usable_cores():
    """Host threads the CPU arm can really use: affinity mask, clipped by the cgroup CPU quota (a container that
    sees 128 CPUs but owns a small quota collapses under 128 oversubscribed threads) and by 32 (the oracle's small
    convolutions stop scaling long before that)."""
    n = len(os.sched_getaffinity(0))
    try:
        q, per = open('/sys/fs/cgroup/cpu.max').read().split()
        if q != 'max':
            n = min(n, max(1, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return max(1, min(n, 32))


def corr_large_shapes(dev):
    """Correlation forward on HBM-sized shapes: no allocation inside the timed region, L2 flushed
    between repetitions, CUDA events on the launching stream.  Algorithmic bytes = B*h*w*(2C+nd)*4."""
    import torch
    from ctypes import c_void_p
    from madstereo._lib import lib, check
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    for name, (b, h, w, c, d, warp) in {'madnet_L2_1920x1088_B8': (8, 272, 480, 32, 2, True),
                                       'dispnet_1280x384': (1, 96, 320, 128, 40, False)}.items():
        nd = 2 * d + 1
        x = torch.randn(b, h, w, c, device=dev); y = torch.randn(b, h, w, c, device=dev)
        u = (torch.rand(b, h, w, 1, device=dev) * 4 - 2) if warp else None
        o = torch.empty(b, h, w, nd, device=dev)
        def run():
            check(lib().ms_corr_fwd(c_void_p(x.data_ptr()), c, c_void_p(y.data_ptr()), c,
                                    c_void_p(u.data_ptr() if warp else 0), 1, c_void_p(o.data_ptr()), nd,
                                    b, h, w, c, d, 1, 0, 0, st), 'ms_corr_fwd')
        for _ in range(3):
            run()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        byts = b * h * w * (2 * c + nd) * 4
        out[name] = {'us': ts[len(ts) // 2] * 1e3, 'bytes': byts, 'gbs': byts / (ts[len(ts) // 2] * 1e-3) / 1e9}
    return out


def conv_dominant_layer(dev):
    """The layer shape that dominates the step (128->128 3x3 at 96x320: estimator-2 / context net)
    timed alone through the operator-level C ABI (weight preparation kernel included), L2 flushed between repetitions."""
    import torch
    from ctypes import c_void_p
    from madstereo._lib import lib, check
    n, h, w, cin, cout = 1, 96, 320, 128, 128
    x = torch.randn(n, h, w, cin, device=dev); wt = torch.randn(3, 3, cin, cout, device=dev) * 0.05
    b = torch.zeros(cout, device=dev); y = torch.empty(n, h, w, cout, device=dev)
    ns = lib().ms_conv2d_tc_scratch(3, 3, cin, cout); scratch = torch.empty(ns, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    def run():
        check(lib().ms_conv2d_fwd_tc(c_void_p(x.data_ptr()), n, h, w, cin, cin, c_void_p(wt.data_ptr()), c_void_p(b.data_ptr()),
                                     c_void_p(y.data_ptr()), cout, cout, 3, 3, 1, 0.2, c_void_p(scratch.data_ptr()), ns, st), 'tc')
    for _ in range(3):
        run()
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    us = ts[len(ts) // 2] * 1e3
    flops = 2.0 * n * h * w * 9 * cin * cout
    return {'shape': '128->128 3x3 @96x320', 'us': us, 'useful_tflops': flops / us / 1e6,
            'traffic': 16362752, 'traffic_src': 'dram__bytes_read+write per launch, profiles/r1_ncu_conv_tc_ts_final_128x128_96x320.txt '
                                                '(algorithmic: 15.7 MB input + 0.6 MB weights; the 15.7 MB output stays in L2)',
            'note': 'conv_tc_ts_kernel + weight-prep kernel; each useful FLOP costs 3 tf32 tensor FLOPs (3xTF32), '
                    'tf32 dense peak = bf16 peak / 2'}


def make_inputs(n_pairs, rank):
    from madstereo.synthetic import make_pair
    return [make_pair(H, W, seed=100 * rank + i)[:2] for i in range(n_pairs)]


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference path restated on host cores (oracle port)
# ---------------------------------------------------------------------------------------------------
def cpu_mad_fps(steps, warmup, sample_note=False):
    import torch
    from oracle.adaptation import OracleAdapter
    from oracle.madnet import init_params
    cores = usable_cores()
    torch.set_num_threads(cores)
    (left, right), = make_inputs(1, 0)
    ad = OracleAdapter(init_params(seed=42), mode='MAD', lr=1e-4)
    k = 0
    for _ in range(warmup):
        ad.step(left, right, k % 5); k += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        ad.step(left, right, k % 5); k += 1
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    fps, ms, cores = cpu_mad_fps(steps, warmup)
    sample = ('%d MAD steps (SEQUENTIAL modules) on one 1280x384 pair after %d warm-up, torch-CPU fp32 oracle '
              'restatement of the TF1 graph (TF1 itself cannot run here)' % (steps, warmup))
    line = {'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'arm': 'reference path on the host cores (oracle port; TF1 cannot run here), '
                                                     'one frame stream whatever N'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------
# this framework
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import Nets
    from madstereo.adaptation import OnlineAdaptation
    from madstereo.synthetic import init_params

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    # keep fd 1 clean for the single JSON line: NCCL / the banner of the reference-style constructors print to stdout
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d' % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    n_pairs = 4
    pairs = make_inputs(n_pairs, rank)
    dev_pairs = [(torch.from_numpy(l).to(dev), torch.from_numpy(r).to(dev)) for l, r in pairs]
    host_pairs = [(torch.from_numpy(l).pin_memory(), torch.from_numpy(r).pin_memory()) for l, r in pairs]

    sys.stdout, real_stdout = sys.stderr, sys.stdout        # keep the reference-style banner off stdout
    net = Nets.get_stereo_net('MADNet', dict(left_img=dev_pairs[0][0], right_img=dev_pairs[0][1], split_layers=[None],
                                             sequence=True, train_portion='BEGIN', bulkhead=True, warping=True,
                                             context_net=True, radius_d=2, stride=1, is_training=False))
    cfg = json.load(open(os.path.join(PKG, 'block_config', 'MadNet_full.json')))
    ad = OnlineAdaptation(net, mode='MAD', train_config=cfg, lr=1e-4, sample_mode='SEQUENTIAL', num_blocks=1)
    ad.load_weights(init_params(net.engine.layers, seed=42))
    sys.stdout = real_stdout
    eng = net.engine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, steps, warmup, pipelined=False):
        # pipelined: the host->device copy of frame i+1 is issued (side stream) while frame i computes; every frame's
        # copy is still inside the timed region
        def one(i):
            if pipelined:
                ad.step(*src[i % n_pairs], prefetch=src[(i + 1) % n_pairs])
            else:
                ad.step(*src[i % n_pairs])
        for i in range(warmup):
            one(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record()
        for i in range(steps):
            one(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    clk = ClockSampler(local) if rank == 0 else None
    if clk:
        clk.start()
    ms_dev, launches = timed(dev_pairs, args.steps, args.warmup)
    ms_e2e_serial, _ = timed(host_pairs, args.steps, 1)
    try:
        ms_e2e, _ = timed(host_pairs, args.steps, 2, pipelined=True)
        e2e_mode = 'pipelined: H2D of frame i+1 on a side stream while frame i computes (OnlineAdaptation.step(prefetch=...))'
    except Exception as ex:                        # keep the serial number rather than no number
        print('pipelined e2e failed: %r' % (ex,), file=sys.stderr)
        ms_e2e, e2e_mode = ms_e2e_serial, 'serial (pipelined path failed: %r)' % (ex,)
    if clk:
        clk.stop_flag = True
        clk.join(timeout=2)

    # ---- kernel-level profile (separate instrumented steps: events around every kernel group)
    prof = None
    eng.profile(True)                         # every rank takes part: the DP step contains collectives
    for i in range(10):
        ad.step(*dev_pairs[i % n_pairs])
    torch.cuda.synchronize()
    prof = eng.profile_read()
    eng.profile(False)
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    corr_large = corr_large_shapes(dev)
    dom = conv_dominant_layer(dev)
    dom['frac_of_bf16_peak'] = dom['useful_tflops'] / pk['bf16_tflops']
    dom['tensor_flops_frac_of_tf32_peak'] = 3.0 * dom['useful_tflops'] / (pk['bf16_tflops'] / 2.0)
    fps = world * args.steps / (ms_dev / 1e3)
    fps_e2e = world * args.steps / (ms_e2e / 1e3)
    conv_ms = sum(prof[c]['ms'] for c in ('conv_fwd', 'conv_dgrad', 'conv_wgrad'))
    conv_macs = sum(prof[c]['macs'] for c in ('conv_fwd', 'conv_dgrad', 'conv_wgrad'))
    conv_calls = sum(prof[c]['calls'] for c in ('conv_fwd', 'conv_dgrad', 'conv_wgrad'))
    conv_tflops = 2.0 * conv_macs / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    corr_gbs = prof['corr_fwd']['bytes'] / (prof['corr_fwd']['ms'] / 1e3) / 1e9 if prof['corr_fwd']['ms'] > 0 else 0.0
    prof_total = sum(v['ms'] for v in prof.values())
    line = {
        'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'parallelism': 'dp%d' % world, 'global_batch': world,
                   'l2': 'no flush: per-step activation+gradient working set ~0.5 GB >> %d MB L2; inputs rotate over '
                         '%d distinct pairs' % (L2_MB, n_pairs),
                   'step': 'set_input, forward, full-res loss, module loss+backward, all-reduce(N>1), momentum update, '
                           'loss D2H (host reward policy needs it every frame)'},
        'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': 2 * H * W * 3 * 4, 'd2h_bytes_per_step': 16,
                'mode': e2e_mode, 'serial_value': world * args.steps / (ms_e2e_serial / 1e3)},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'kernel': 'conv stack: conv_tc_ts/conv_tc kernels (tcgen05 3xTF32 implicit GEMM, stride-1 fwd+dgrad) + wgrad_tc_kernel (tcgen05 stride-1 wgrad) + conv_gemm/conv_wgrad (fp32 CUDA-core: stride-2, cin=3 and cout=1 layers); useful FLOPs = 2*MACs, the 3x tf32 passes are not counted',
                     'achieved': conv_tflops, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                     'frac': conv_tflops / pk['bf16_tflops_sustained'], 'peak_src': pk['src'] + ' bf16 sustained (kernels timed inside a long step)',
                     'traffic': None, 'avg_launch_us': 1e3 * conv_ms / max(conv_calls, 1),
                     'share_of_step': conv_ms / prof_total if prof_total else None,
                     'dominant_layer': dom},
        'corr_kernel': {'bound': 'hbm', 'achieved': corr_gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                        'frac': corr_gbs / pk['hbm_gbs'],
                        'note': 'in-step number: all 5 MADNet levels at 1280x384 (<=8.5 MB each: L2-resident, launch-latency bound); '
                                'algorithmic bytes B*h*w*(2C+5)*4',
                        'large': {k: dict(v, frac=v['gbs'] / pk['hbm_gbs']) for k, v in corr_large.items()},
                        'traffic': {'madnet_L2_1920x1088_B8': 289069312,
                                    'src': 'dram__bytes_read.sum + dram__bytes_write.sum, profiles/r1_ncu_corr_fwd4_L2_1920x1088_B8.txt'}},
        'profile_ms_per_step': {k: v['ms'] / 10.0 for k, v in prof.items()},
        'clocks': clk.summary() if clk else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        fps_cpu, ms_cpu, cores = cpu_mad_fps(3, 1)
        line['cpu_baseline'] = {'value': fps_cpu, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                                'sample': '3 MAD steps (modules 1,2,3 of the SEQUENTIAL cycle) on one 1280x384 pair after '
                                          '1 warm-up; torch-CPU fp32 oracle restatement (TF1 cannot run here)'}
    sys.stdout.flush()
    os.dup2(saved_stdout_fd, 1)
    print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
