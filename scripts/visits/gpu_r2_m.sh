#!/bin/bash
# visit m: conv2d_transpose gradients on tcgen05 (DispNet), tile-shape search A/B
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 300 > gpurun_out/m_convbf.log 2>&1
echo "rc=$?" >> gpurun_out/m_convbf.log
tail -n 4 gpurun_out/m_convbf.log
for ts in 0 1; do
  for cfg in 3 2 4; do
    MS_BF_TILE_SEARCH=$ts timeout -s KILL 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-corr-shapes > gpurun_out/m_bench_cfg${cfg}_ts${ts}.log 2>&1
    echo "cfg$cfg ts$ts: $(tail -n 1 gpurun_out/m_bench_cfg${cfg}_ts${ts}.log | cut -c1-160)"
  done
  MS_BF_TILE_SEARCH=$ts timeout -s KILL 300 python bench.py --config 5 --batch 8 --steps 10 --warmup 3 --no-corr-shapes > gpurun_out/m_bench_cfg5_ts${ts}.log 2>&1
  echo "cfg5 ts$ts: $(tail -n 1 gpurun_out/m_bench_cfg5_ts${ts}.log | cut -c1-160)"
done
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/m_layers4.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 3 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/m_layers3.log 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/m_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/m_suite.log
tail -n 4 gpurun_out/m_suite.log
