#!/bin/bash
# visit p: wgrad_bf row groups for tall filters, DispNet stem on direct kernels
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -k "stem" > gpurun_out/p_stem.log 2>&1
echo "rc=$?" >> gpurun_out/p_stem.log
tail -n 4 gpurun_out/p_stem.log
rm -f gpurun_out/conv_bf_errors.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 300 > gpurun_out/p_convbf.log 2>&1
echo "rc=$?" >> gpurun_out/p_convbf.log
tail -n 4 gpurun_out/p_convbf.log
for gq in 0 1; do
  echo "MS_WB_GROUPS=$gq" >> gpurun_out/p_wb_groups.log
  MS_WB_GROUPS=$gq timeout -s KILL 200 python scripts/bf_bench.py sel 12 15 16 17 2>&1 | tail -n 4 >> gpurun_out/p_wb_groups.log
done
cat gpurun_out/p_wb_groups.log
timeout -s KILL 300 python bench.py --config 4 --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/p_bench_cfg4.log 2>&1
echo "cfg4: $(tail -n 1 gpurun_out/p_bench_cfg4.log | cut -c1-200)"
MS_STEM=0 MS_WB_GROUPS=0 timeout -s KILL 300 python bench.py --config 4 --steps 30 --warmup 5 --no-corr-shapes --no-parity-check > gpurun_out/p_bench_cfg4_off.log 2>&1
echo "cfg4 (stem + groups off): $(tail -n 1 gpurun_out/p_bench_cfg4_off.log | cut -c1-200)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/p_layers4.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -k "dispnet or Dispnet" > gpurun_out/p_engine_tests.log 2>&1
echo "rc=$?" >> gpurun_out/p_engine_tests.log
tail -n 4 gpurun_out/p_engine_tests.log
