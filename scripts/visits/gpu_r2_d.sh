#!/bin/bash
# kill-switch timing of conv_bf on the dominant layer (what bounds the main loop?) + DispNet bisect
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
for D in 0 1 2 4 6 7 8 9 15; do
  echo "MS_BF_DEBUG=$D" >> gpurun_out/d_kill.log
  MS_BF_DEBUG=$D timeout -s KILL 200 python scripts/bf_bench.py sel 0 6 7 >> gpurun_out/d_kill.log 2>&1
done
for N in 128 64; do
  echo "MS_BF_N=$N" >> gpurun_out/d_kill.log
  MS_BF_N=$N timeout -s KILL 200 python scripts/bf_bench.py sel 0 6 7 >> gpurun_out/d_kill.log 2>&1
done
for V in "MS_CONV_IMPL=tf32" "MS_BF_WGRAD=0" "MS_HEADS=0" "X=1"; do
  echo "== $V" >> gpurun_out/d_dispnet.log
  env $V timeout -s KILL 600 python -m pytest tests/test_dispnet_gpu.py -q --timeout 300 2>&1 | grep -E "passed|failed|Error|assert " | head -8 >> gpurun_out/d_dispnet.log
done
timeout -s KILL 900 python bench.py --config 4 --steps 10 --warmup 3 --no-corr-shapes > gpurun_out/d_bench_cfg4.log 2>&1
timeout -s KILL 900 python bench.py --config 2 --steps 10 --warmup 3 --no-corr-shapes > gpurun_out/d_bench_cfg2.log 2>&1
tail -40 gpurun_out/d_kill.log
