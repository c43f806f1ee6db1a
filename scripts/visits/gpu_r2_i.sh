#!/bin/bash
# visit i: validation (default settings) + experiments: programmatic dependent launch, split-K cap, DispNet stem gradients
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/i_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/i_conv_bf.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/i_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/i_suite.log
B="python bench.py --steps 50 --warmup 10 --no-corr-shapes --no-cpu-baseline"
timeout -s KILL 600 $B > gpurun_out/i_bench_default.log 2>&1
MS_PDL=1 timeout -s KILL 600 $B > gpurun_out/i_bench_pdl.log 2>&1
MS_BF_KSPLIT_MAX=16 timeout -s KILL 600 $B > gpurun_out/i_bench_k16.log 2>&1
MS_BF_KSPLIT_MAX=4 timeout -s KILL 600 $B > gpurun_out/i_bench_k4.log 2>&1
MS_PDL=1 MS_BF_KSPLIT_MAX=16 timeout -s KILL 600 $B > gpurun_out/i_bench_pdl_k16.log 2>&1
MS_PDL=1 timeout -s KILL 900 python -m pytest tests/test_madnet_gpu.py tests/test_conv_bf_gpu.py -q --timeout 300 > gpurun_out/i_suite_pdl.log 2>&1
echo "pdl suite rc=$?" >> gpurun_out/i_suite_pdl.log
MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --config 4 --steps 20 --warmup 5 --no-corr-shapes > gpurun_out/i_bench_cfg4.log 2>&1
for f in gpurun_out/i_bench_*.log; do echo $f; tail -1 $f | cut -c1-170; done
tail -3 gpurun_out/i_conv_bf.log gpurun_out/i_suite.log gpurun_out/i_suite_pdl.log
