#!/bin/bash
# visit k: wgrad side-stream overlap -- validation + timing; 50-step drift table
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/k_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/k_conv_bf.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/k_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/k_suite.log
B="python bench.py --steps 50 --warmup 10 --no-corr-shapes --no-cpu-baseline"
timeout -s KILL 600 $B > gpurun_out/k_bench_default.log 2>&1
MS_WGRAD_OVERLAP=0 timeout -s KILL 600 $B > gpurun_out/k_bench_nooverlap.log 2>&1
timeout -s KILL 900 python bench.py --config 2 --steps 20 --warmup 5 --no-corr-shapes --no-cpu-baseline > gpurun_out/k_bench_cfg2.log 2>&1
timeout -s KILL 900 python bench.py --config 5 --batch 8 --steps 5 --warmup 3 --no-corr-shapes --no-cpu-baseline > gpurun_out/k_bench_cfg5_b8.log 2>&1
timeout -s KILL 900 python scripts/drift_50.py 50 > gpurun_out/k_drift.log 2>&1
for f in gpurun_out/k_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-170; done
tail -n 3 gpurun_out/k_conv_bf.log gpurun_out/k_suite.log; tail -n 4 gpurun_out/k_drift.log
