#!/bin/bash
# round-2 first visit: conv_bf op tests, then the full GPU suite, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/a_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/a_conv_bf.log
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_conv_bf_gpu.py > gpurun_out/a_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/a_suite.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/a_bench.log
MS_CONV_IMPL=tf32 timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_tf32.log 2>&1
tail -3 gpurun_out/a_conv_bf.log gpurun_out/a_suite.log gpurun_out/a_bench.log
