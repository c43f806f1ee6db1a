#!/bin/bash
# visit u: evidence for the final build -- ncu captures (conv_bf, wgrad_bf, corr_mma), launch list, sanitizer on the new kernels; carve-out A/B
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
for c in 0 1; do
  MS_CARVEOUT=$c timeout -s KILL 300 python bench.py --config 3 --steps 40 --warmup 8 --no-corr-shapes --no-parity-check > gpurun_out/u_carve$c.log 2>&1
  echo "MS_CARVEOUT=$c: $(tail -n 1 gpurun_out/u_carve$c.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"],1), "FPS", round(j["ms_per_step"],3), "ms")' 2>&1)"
done
NCU="ncu --set full --clock-control none --import-source on"
timeout -s KILL 300 $NCU -k regex:conv_bf_kernel -c 1 -s 2 -o gpurun_out/u_ncu_conv_bf python scripts/bf_bench.py one 0 > gpurun_out/u_ncu1.log 2>&1; tail -n 1 gpurun_out/u_ncu1.log
timeout -s KILL 300 $NCU -k regex:wgrad_bf_kernel -c 1 -s 2 -o gpurun_out/u_ncu_wgrad_bf python scripts/bf_bench.py one 0 > gpurun_out/u_ncu2.log 2>&1; tail -n 1 gpurun_out/u_ncu2.log
timeout -s KILL 300 $NCU -k regex:corr_mma -c 2 -s 4 -o gpurun_out/u_ncu_corr_mma python scripts/corr_wide_one.py > gpurun_out/u_ncu3.log 2>&1; tail -n 1 gpurun_out/u_ncu3.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/u_launches.csv python bench.py --steps 2 --warmup 1 --no-corr-shapes --no-parity-check > gpurun_out/u_launches_bench.log 2>&1
tail -n 1 gpurun_out/u_launches.csv | cut -c1-200
SEL_OPS='(wide_window or stem or case18] or case19] or case20] or case21]) and not shape4]'
SEL_BF='(transpose and (case4] or case5])) or ((forward or dgrad or test_wgrad) and (case0] or case9] or case17]))'
for TOOL in memcheck racecheck; do
  timeout -s KILL 900 compute-sanitizer --tool $TOOL --log-file gpurun_out/u_sanitizer_$TOOL.ops.log \
    python -m pytest -q -p no:cacheprovider --timeout 800 tests/test_ops_gpu.py -k "$SEL_OPS" > gpurun_out/u_sanitizer_$TOOL.ops.pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/u_sanitizer_$TOOL.ops.pytest.log
  timeout -s KILL 900 compute-sanitizer --tool $TOOL --log-file gpurun_out/u_sanitizer_$TOOL.bf.log \
    python -m pytest -q -p no:cacheprovider --timeout 800 tests/test_conv_bf_gpu.py -k "$SEL_BF" > gpurun_out/u_sanitizer_$TOOL.bf.pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/u_sanitizer_$TOOL.bf.pytest.log
done
tail -n 2 gpurun_out/u_sanitizer_*.log | cut -c1-160
tail -n 2 gpurun_out/u_sanitizer_*.pytest.log | cut -c1-160
