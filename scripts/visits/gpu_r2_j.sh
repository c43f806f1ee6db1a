#!/bin/bash
# visit j: PDL on every kernel of the step, wgrad_bf parity patches -- validation + timing
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/j_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/j_conv_bf.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/j_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/j_suite.log
B="python bench.py --steps 50 --warmup 10 --no-corr-shapes --no-cpu-baseline"
timeout -s KILL 600 $B > gpurun_out/j_bench_default.log 2>&1
MS_PDL=0 timeout -s KILL 600 $B > gpurun_out/j_bench_nopdl.log 2>&1
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/j_bf_bench.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --config 4 --steps 20 --warmup 5 --no-corr-shapes --no-cpu-baseline > gpurun_out/j_bench_cfg4.log 2>&1
timeout -s KILL 900 python bench.py --config 2 --steps 20 --warmup 5 --no-corr-shapes --no-cpu-baseline > gpurun_out/j_bench_cfg2.log 2>&1
timeout -s KILL 900 python bench.py --config 1 --steps 50 --warmup 10 --no-corr-shapes --no-cpu-baseline > gpurun_out/j_bench_cfg1.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1
for f in gpurun_out/j_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-170; done
tail -n 3 gpurun_out/j_conv_bf.log gpurun_out/j_suite.log gpurun_out/j_smoke.log
