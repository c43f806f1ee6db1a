#!/bin/bash
# visit h: DispNet stem / conv_transpose on the tensor-core path, split-K heuristic, prefetch order -- suite + benches
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/h_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/h_conv_bf.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/h_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/h_suite.log
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/h_bf_bench.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --steps 50 --warmup 10 > gpurun_out/h_bench_cfg3.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --config 4 --steps 20 --warmup 5 --no-corr-shapes > gpurun_out/h_bench_cfg4.log 2>&1
timeout -s KILL 900 python bench.py --config 2 --steps 20 --warmup 5 --no-corr-shapes --no-cpu-baseline > gpurun_out/h_bench_cfg2.log 2>&1
MS_BF_SPLIT_CYCLES=1 timeout -s KILL 900 python bench.py --steps 50 --warmup 10 --no-corr-shapes --no-cpu-baseline --no-parity-check > gpurun_out/h_bench_cfg3_oldsplit.log 2>&1
tail -3 gpurun_out/h_conv_bf.log gpurun_out/h_suite.log
for f in gpurun_out/h_bench_cfg*.log; do tail -1 $f | cut -c1-200; done
