#!/bin/bash
# visit g: validation of the final conv_bf epilogue, all BASELINE configurations, per-layer table, ncu evidence, sanitizer
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/g_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/g_conv_bf.log
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/g_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/g_suite.log
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/g_bf_bench.log 2>&1
MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --steps 50 --warmup 10 > gpurun_out/g_bench_cfg3.log 2>&1
for C in 1 2 4; do MS_BENCH_LAYERS=1 timeout -s KILL 900 python bench.py --config $C --steps 20 --warmup 5 --no-corr-shapes > gpurun_out/g_bench_cfg$C.log 2>&1; done
timeout -s KILL 900 python bench.py --config 5 --steps 10 --warmup 3 --no-corr-shapes > gpurun_out/g_bench_cfg5_b1.log 2>&1
timeout -s KILL 900 python bench.py --config 5 --batch 8 --steps 5 --warmup 3 --no-corr-shapes --no-cpu-baseline > gpurun_out/g_bench_cfg5_b8.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_bf_kernel -c 1 -o gpurun_out/g_ncu_conv_bf_128 python scripts/bf_bench.py one 0 > gpurun_out/g_ncu1.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf_kernel -c 1 -o gpurun_out/g_ncu_wgrad_bf_128 python scripts/bf_bench.py one 0 > gpurun_out/g_ncu2.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:corr_fwd4_kernel -c 1 -o gpurun_out/g_ncu_corr_fused python scripts/corr_one.py fused > gpurun_out/g_ncu3.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check --no-corr-shapes > gpurun_out/g_launch_bench.log 2>&1
timeout -s KILL 1800 bash scripts/sanitize.sh > gpurun_out/g_sanitize.log 2>&1
tail -3 gpurun_out/g_conv_bf.log gpurun_out/g_suite.log
for f in gpurun_out/g_bench_cfg*.log; do tail -1 $f | cut -c1-200; done
