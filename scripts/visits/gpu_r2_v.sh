#!/bin/bash
# visit v: final validation of the round -- full GPU suite, smoke, default bench + reference arm, all five configurations
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/v_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/v_suite.log
tail -n 4 gpurun_out/v_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; tail -n 2 gpurun_out/v_smoke.log
timeout -s KILL 900 python bench.py --gpus 1 --steps 50 --warmup 10 > gpurun_out/v_bench_default.log 2>&1
echo "default: $(tail -n 1 gpurun_out/v_bench_default.log | cut -c1-200)"
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/v_bench_reference.log 2>&1
echo "reference: $(tail -n 1 gpurun_out/v_bench_reference.log | cut -c1-200)"
for cfg in 1 2 4; do
  timeout -s KILL 600 python bench.py --config $cfg --steps 40 --warmup 8 > gpurun_out/v_bench_cfg${cfg}.log 2>&1
  echo "cfg$cfg: $(tail -n 1 gpurun_out/v_bench_cfg${cfg}.log | cut -c1-170)"
done
timeout -s KILL 600 python bench.py --config 5 --batch 8 --steps 10 --warmup 3 > gpurun_out/v_bench_cfg5_b8.log 2>&1
echo "cfg5 b8: $(tail -n 1 gpurun_out/v_bench_cfg5_b8.log | cut -c1-170)"
timeout -s KILL 600 python bench.py --config 5 --steps 20 --warmup 5 --no-corr-shapes > gpurun_out/v_bench_cfg5_b1.log 2>&1
echo "cfg5 b1: $(tail -n 1 gpurun_out/v_bench_cfg5_b1.log | cut -c1-170)"
