#!/bin/bash
# visit n: banded mma.sync correlation (DispNet, forward + backward), strided-wgrad diagnosis
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py -q --timeout 300 -k "corr" > gpurun_out/n_corr.log 2>&1
echo "rc=$?" >> gpurun_out/n_corr.log
tail -n 6 gpurun_out/n_corr.log
python - <<'PY'
import json
try:
    j = json.load(open('gpurun_out/corr_vs_reference.json')); print(json.dumps(j.get('dispnet_1280x384')))
except Exception as e: print(e)
PY
timeout -s KILL 300 python scripts/bf_bench.py sel 12 15 16 17 18 19 > gpurun_out/n_bf_bench.log 2>&1
cat gpurun_out/n_bf_bench.log | tail -n 8
timeout -s KILL 300 python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/n_bench_cfg4.log 2>&1
echo "cfg4: $(tail -n 1 gpurun_out/n_bench_cfg4.log | cut -c1-200)"
MS_BENCH_LAYERS=1 timeout -s KILL 300 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes --no-parity-check > gpurun_out/n_layers4.log 2>&1
timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -k "dispnet or Dispnet or config4 or cfg4" > gpurun_out/n_dispnet_tests.log 2>&1
echo "rc=$?" >> gpurun_out/n_dispnet_tests.log
tail -n 4 gpurun_out/n_dispnet_tests.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf_kernel -c 1 -o gpurun_out/n_wgrad_s2 python scripts/bf_bench.py one 15 > gpurun_out/n_ncu.log 2>&1
tail -n 2 gpurun_out/n_ncu.log
