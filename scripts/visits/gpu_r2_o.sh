#!/bin/bash
# visit o: why does wgrad_bf spend ~190 cycles per MMA? (kill switches) + single-channel head wgrad kernel
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -k "conv2d_fwd_dgrad_wgrad" > gpurun_out/o_ops.log 2>&1
echo "rc=$?" >> gpurun_out/o_ops.log
tail -n 4 gpurun_out/o_ops.log
for d in 0 1 2 4 8 3; do
  echo "MS_WB_DEBUG=$d" >> gpurun_out/o_wb_debug.log
  MS_WB_DEBUG=$d timeout -s KILL 200 python scripts/bf_bench.py sel 0 15 18 2>&1 | tail -n 3 >> gpurun_out/o_wb_debug.log
done
cat gpurun_out/o_wb_debug.log
timeout -s KILL 300 python bench.py --config 4 --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/o_bench_cfg4.log 2>&1
echo "cfg4: $(tail -n 1 gpurun_out/o_bench_cfg4.log | cut -c1-200)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/o_layers4.log 2>&1
timeout -s KILL 300 python bench.py --config 3 --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/o_bench_cfg3.log 2>&1
echo "cfg3: $(tail -n 1 gpurun_out/o_bench_cfg3.log | cut -c1-200)"
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -k "dispnet or Dispnet or madnet_mad_step or full_step" > gpurun_out/o_engine_tests.log 2>&1
echo "rc=$?" >> gpurun_out/o_engine_tests.log
tail -n 4 gpurun_out/o_engine_tests.log
