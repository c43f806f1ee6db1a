#!/bin/bash
# N-GPU visit (gpurun --gpus N): data-parallel parity (tests/dp_check.py) + weak-scaling bench, peer-memory exchange vs NCCL
N=${1:-2}
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29511 tests/dp_check.py > gpurun_out/dp${N}_check.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/dp${N}_check.log
MS_DP_IMPL=nccl timeout -s KILL 600 $TR --master-port 29512 tests/dp_check.py > gpurun_out/dp${N}_check_nccl.log 2>&1
echo "dp_check(nccl) rc=$?" >> gpurun_out/dp${N}_check_nccl.log
timeout -s KILL 600 python bench.py --gpus 1 --steps 60 --warmup 10 --no-cpu-baseline --no-corr-shapes --no-parity-check > gpurun_out/dp${N}_bench1.log 2>&1
timeout -s KILL 600 $TR --master-port 29513 bench.py --gpus $N --steps 60 --warmup 10 --no-corr-shapes > gpurun_out/dp${N}_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/dp${N}_bench.log
MS_DP_IMPL=nccl timeout -s KILL 600 $TR --master-port 29514 bench.py --gpus $N --steps 60 --warmup 10 --no-corr-shapes --no-parity-check > gpurun_out/dp${N}_bench_nccl.log 2>&1
grep -h "DP \|rc=" gpurun_out/dp${N}_check.log gpurun_out/dp${N}_check_nccl.log
for f in gpurun_out/dp${N}_bench1.log gpurun_out/dp${N}_bench.log gpurun_out/dp${N}_bench_nccl.log; do tail -1 $f | cut -c1-160; done
