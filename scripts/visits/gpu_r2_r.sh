#!/bin/bash
# visit r: warp-uniform tcgen05 / TMA issue loops (elected lane) in conv_bf and wgrad_bf
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 300 python scripts/mma_probe.py > gpurun_out/mma_probe.log 2>&1
tail -n 66 gpurun_out/mma_probe.log
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r_suite.log
tail -n 5 gpurun_out/r_suite.log
timeout -s KILL 300 python scripts/bf_bench.py sel 0 6 7 12 15 18 > gpurun_out/r_bf_bench.log 2>&1
cat gpurun_out/r_bf_bench.log | tail -n 7
for cfg in 3 2 4 1; do
  timeout -s KILL 300 python bench.py --config $cfg --steps 40 --warmup 8 --no-corr-shapes > gpurun_out/r_bench_cfg${cfg}.log 2>&1
  echo "cfg$cfg: $(tail -n 1 gpurun_out/r_bench_cfg${cfg}.log | cut -c1-170)"
done
timeout -s KILL 300 python bench.py --config 5 --batch 8 --steps 10 --warmup 3 --no-corr-shapes > gpurun_out/r_bench_cfg5.log 2>&1
echo "cfg5: $(tail -n 1 gpurun_out/r_bench_cfg5.log | cut -c1-170)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 3 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/r_layers3.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/r_layers4.log 2>&1
