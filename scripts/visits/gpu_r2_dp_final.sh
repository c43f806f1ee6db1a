#!/bin/bash
# 2-GPU sanity of the final build (gpurun --gpus 2): data-parallel parity + N = 1 / N = 2 bench lines, short
N=${1:-2}
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 400 $TR --master-port 29511 tests/dp_check.py > gpurun_out/w_dp${N}_check.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/w_dp${N}_check.log



grep -h "DP \|rc=" gpurun_out/w_dp${N}_check.log
for f in gpurun_out/w_dp${N}_bench1.log gpurun_out/w_dp${N}_bench.log; do tail -n 2 $f | cut -c1-170; done
