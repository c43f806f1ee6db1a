#!/bin/bash
# visit q: tcgen05.mma cost table (operand layouts x N), stem forward v2, row groups only for >= 5 rows
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 300 python scripts/mma_probe.py > gpurun_out/mma_probe.log 2>&1
cat gpurun_out/mma_probe.log | tail -n 50
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -k "stem" > gpurun_out/q_stem.log 2>&1
echo "rc=$?" >> gpurun_out/q_stem.log
tail -n 3 gpurun_out/q_stem.log
timeout -s KILL 300 python bench.py --config 4 --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/q_bench_cfg4.log 2>&1
echo "cfg4: $(tail -n 1 gpurun_out/q_bench_cfg4.log | cut -c1-200)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/q_layers4.log 2>&1
timeout -s KILL 300 python bench.py --config 2 --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/q_bench_cfg2.log 2>&1
echo "cfg2: $(tail -n 1 gpurun_out/q_bench_cfg2.log | cut -c1-200)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 2 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/q_layers2.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -k "dispnet or Dispnet" > gpurun_out/q_engine_tests.log 2>&1
echo "rc=$?" >> gpurun_out/q_engine_tests.log
tail -n 3 gpurun_out/q_engine_tests.log
