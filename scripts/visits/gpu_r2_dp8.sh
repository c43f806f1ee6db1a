#!/bin/bash
# 8-GPU visit: data-parallel parity at N=8 + weak scaling N=1,2,4,8 (peer-memory fused exchange; NCCL path for comparison at 8)
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29511 tests/dp_check.py > gpurun_out/dp8_check.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/dp8_check.log
B="bench.py --steps 100 --warmup 10 --no-corr-shapes --no-cpu-baseline --no-parity-check"
timeout -s KILL 600 python $B --gpus 1 > gpurun_out/dp8_bench1.log 2>&1
timeout -s KILL 600 $TR --nproc-per-node 2 --master-port 29512 $B --gpus 2 > gpurun_out/dp8_bench2.log 2>&1
timeout -s KILL 600 $TR --nproc-per-node 4 --master-port 29513 $B --gpus 4 > gpurun_out/dp8_bench4.log 2>&1
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29514 $B --gpus 8 > gpurun_out/dp8_bench8.log 2>&1
echo "bench8 rc=$?" >> gpurun_out/dp8_bench8.log
MS_DP_IMPL=nccl timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29515 $B --gpus 8 > gpurun_out/dp8_bench8_nccl.log 2>&1
timeout -s KILL 600 $TR --nproc-per-node 8 --master-port 29516 bench.py --config 5 --steps 20 --warmup 5 --gpus 8 --no-corr-shapes --no-cpu-baseline --no-parity-check > gpurun_out/dp8_bench_cfg5.log 2>&1
grep -h "DP \|rc=" gpurun_out/dp8_check.log
for f in gpurun_out/dp8_bench1.log gpurun_out/dp8_bench2.log gpurun_out/dp8_bench4.log gpurun_out/dp8_bench8.log gpurun_out/dp8_bench8_nccl.log gpurun_out/dp8_bench_cfg5.log; do tail -n 2 $f | cut -c1-160; done
