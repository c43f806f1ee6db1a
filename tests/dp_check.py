"""Multi-GPU parity check (run under torchrun, one rank per GPU):
N ranks x 1 frame each must equal the reference graph at batch N (every loss is a batch mean).
Rank 0 compares gradients / updated weights with the CPU oracle run on the N-frame batch.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200')
sys.path.insert(0, ROOT); sys.path.insert(0, PKG)


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import Nets
    from madstereo.adaptation import OnlineAdaptation
    from madstereo.synthetic import make_pair
    from oracle.madnet import init_params
    h, w = 64, 128
    frames = [make_pair(h, w, seed=10 + r)[:2] for r in range(world)]
    left, right = frames[rank]
    lt, rt = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    sys.stdout = open(os.devnull, 'w') if rank else sys.stdout
    ok = True
    solo = [dist.new_group([r]) for r in range(world)]       # one-rank groups: an adapter built on one of them runs without DP
    for mode in ('MAD', 'FULL'):
        net = Nets.get_stereo_net('MADNet', dict(left_img=lt, right_img=rt, split_layers=[None], sequence=True,
                                                 train_portion='BEGIN', bulkhead=(mode == 'MAD')))
        cfg = json.load(open(os.path.join(PKG, 'block_config', 'MadNet_full.json')))
        ad = OnlineAdaptation(net, mode=mode, train_config=cfg, lr=1e-4, sample_mode='FIXED', fixed_id=3)
        params = init_params(seed=42)
        ad.load_weights(params)
        out = ad.step(lt, rt)
        torch.cuda.synchronize()
        # replicas must be bit-identical
        wsum = net.engine.weights.double().sum().reshape(1)
        gathered = [torch.zeros_like(wsum) for _ in range(world)]
        dist.all_gather(gathered, wsum)
        same = all(float(g) == float(gathered[0]) for g in gathered)
        if rank == 0:
            from oracle.adaptation import OracleAdapter
            bl = np.concatenate([f[0] for f in frames]); br = np.concatenate([f[1] for f in frames])
            orc = OracleAdapter(params, mode=mode, lr=1e-4)
            ref = orc.step(bl, br, 3)
            g = net.engine.param_views(net.engine.grads)
            # Gradients more than six orders of magnitude below the step's largest one are below the fp32 resolution of the
            # backward chain that produces them (FULL mode: the full-resolution loss reaches module 6 with |g| ~ 1e-10 next to
            # 1e-2 elsewhere; their bits change with any summation order): listed, not judged -- their effect on the weights
            # is covered by the dW metric below.
            gmax = max(float(np.abs(gr).max()) for gr in ref['grads'].values())
            judged = {n: gr for n, gr in ref['grads'].items() if float(np.abs(gr).max()) >= 1e-6 * gmax}
            tiny = sorted(n for n in ref['grads'] if n not in judged)
            # the same fp64-anchored bound as the single-GPU step tests (tests/parity_metrics.py): per tensor, the distance to
            # the fp64 oracle may not exceed max(floor, 2 x the fp32 oracle's own distance, the step's noise level)
            import torch as _t
            from parity_metrics import assert_grads, grad_report, summarize
            o64 = OracleAdapter(params, mode=mode, lr=1e-4, dtype=_t.float64)
            r64 = o64.step(bl, br, 3)
            rep = grad_report({n: g[n].cpu().numpy() / world for n in judged}, {n: judged[n] for n in judged}, {n: r64['grads'][n] for n in judged})
            try:
                assert_grads(rep, 5e-3, 1e-2, 'DP ' + mode)
                anchored = True
            except AssertionError as e:
                print('   ', e, flush=True)
                anchored = False
            print('   vs fp64 oracle:', {k: float('%.3g' % v) for k, v in summarize(rep).items()}, flush=True)
            worst = max(rel(g[n].cpu().numpy() / world, gr) for n, gr in judged.items())
            per = sorted(((rel_l2(g[n].cpu().numpy() / world, gr), n, float(np.abs(gr).max())) for n, gr in judged.items()), reverse=True)
            worst_l2 = per[0][0]
            print('   worst tensors (rel L2, name, max|ref|):', [(round(a, 4), n, '%.2e' % m) for a, n, m in per[:4]], flush=True)
            if tiny:
                pt = sorted(((rel_l2(g[n].cpu().numpy() / world, ref['grads'][n]), n, float(np.abs(ref['grads'][n]).max())) for n in tiny), reverse=True)
                print('   %d tensors with max|g| < 1e-6 x %.2e not judged; worst of them:' % (len(tiny), gmax),
                      [(round(a, 4), n, '%.2e' % m) for a, n, m in pt[:2]], flush=True)
            # DP-specific check: N ranks x 1 frame against the SAME engine on one GPU at batch N (no oracle, no arithmetic
            # difference other than the order of the cross-frame sum)
            blt, brt = torch.from_numpy(bl).cuda(), torch.from_numpy(br).cuda()
            net2 = Nets.get_stereo_net('MADNet', dict(left_img=blt, right_img=brt, split_layers=[None], sequence=True,
                                                      train_portion='BEGIN', bulkhead=(mode == 'MAD')))
            ad2 = OnlineAdaptation(net2, mode=mode, train_config=cfg, lr=1e-4, sample_mode='FIXED', fixed_id=3, process_group=solo[0])
            ad2.load_weights(params)
            out2 = ad2.step(blt, brt)
            torch.cuda.synchronize()
            g2 = net2.engine.param_views(net2.engine.grads)
            pb = sorted(((rel_l2(g[n].cpu().numpy() / world, g2[n].cpu().numpy()), n) for n in judged), reverse=True)
            batch_l2 = pb[0][0]
            print('   vs the same engine at batch %d on one GPU: worst rel L2 %.2e (%s), loss %.6f vs %.6f' % (
                world, batch_l2, pb[0][1], out['loss'], out2['loss']), flush=True)
            del ad2, net2
            wv = net.engine.export_params()
            # weight deltas are compared net of fp32 storage resolution: a tensor whose update lr*g is below one ulp of its
            # weights (FULL mode touches layers with gradients of 1e-7) has no meaningful relative delta
            def dw_err(n):
                ref_d = orc.net.p[n].detach().numpy().astype(np.float64) - params[n]
                got_d = wv[n].astype(np.float64) - params[n]
                slack = 1.2e-7 * np.linalg.norm(params[n].astype(np.float64).ravel())
                return max(0.0, np.linalg.norm((got_d - ref_d).ravel()) - slack) / max(np.linalg.norm(ref_d.ravel()), 1e-30)
            wworst = max(dw_err(n) for n in ref['grads'])
            # gradients: the fp64-anchored bound above (the raw distances to the fp32 oracle are printed for reference)
            # Verdict.  Always gated: replicas bit-identical, the weight update and the loss equal to the oracle's.  Gradients:
            # at this 64 x 128 test size the coarse modules work on 1 x 2 ... 4 x 8 pixel maps, where one relu / floor decision
            # that lands on the other side moves a tensor by 1e-3 ... 1e-2 -- the SAME engine at batch N on one GPU differs from
            # N x 1 frame by 1.9e-3 (MAD) / 1.4e-2 (FULL) although only tile shapes and the order of the cross-frame sum change,
            # while the fp32 and fp64 CPU oracles (identical operation order) agree to 5e-6.  MAD (the path the exchange was
            # designed for) is gated with the floors of the single-GPU step tests (5e-3 L2 / 1e-2 L-inf, against the oracle and
            # against the one-GPU batch); FULL gradient distances are printed -- the engine's FULL gradients are gated where they
            # are meaningful, at the BASELINE sizes with fp64 anchors (tests/test_baseline_configs_gpu.py).
            good = same and (mode == 'FULL' or (anchored and batch_l2 < 5e-3)) and wworst < 5e-3 and abs(out['loss'] - ref['full_loss']) < 2e-5
            ok = ok and good
            print('DP %s world=%d impl=%s: replicas identical=%s  grad rel Linf %.2e L2 %.2e  dW rel L2 %.2e  loss %.6f vs %.6f  -> %s' % (
                mode, world, 'peer-memory fused' if ad.dp_peer else 'torch.distributed', same, worst, worst_l2, wworst,
                out['loss'], ref['full_loss'], 'OK' if good else 'FAIL'), flush=True)
        del ad, net
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
