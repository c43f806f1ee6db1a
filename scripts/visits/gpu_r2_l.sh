#!/bin/bash
# visit l: validation after the continual-adaptation (proxy loss) merge
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/l_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/l_suite.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 10 > gpurun_out/l_bench_default.log 2>&1
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/l_bench_reference.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1
tail -n 3 gpurun_out/l_suite.log gpurun_out/l_smoke.log; tail -n 1 gpurun_out/l_bench_default.log | cut -c1-200; tail -n 1 gpurun_out/l_bench_reference.log | cut -c1-200
