#!/bin/bash
# visit f: scale-carrying fp16 planes (DispNet fix), per-CTA clock profile of conv_bf, suite, bench
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/f_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/f_conv_bf.log
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/f_bf_bench.log 2>&1
for S in 0 7 8; do MS_BF_PROF=1 timeout -s KILL 120 python scripts/bf_bench.py prof $S >> gpurun_out/f_prof.log 2>&1; done
echo "--- no loads (MS_BF_DEBUG=6)" >> gpurun_out/f_prof.log
MS_BF_DEBUG=6 MS_BF_PROF=1 timeout -s KILL 120 python scripts/bf_bench.py prof 0 >> gpurun_out/f_prof.log 2>&1
echo "--- no MMA no loads no stores (15)" >> gpurun_out/f_prof.log
MS_BF_DEBUG=15 MS_BF_PROF=1 timeout -s KILL 120 python scripts/bf_bench.py prof 0 >> gpurun_out/f_prof.log 2>&1
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/f_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/f_suite.log
timeout -s KILL 900 python bench.py --steps 30 --warmup 5 > gpurun_out/f_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/f_bench.log
cat gpurun_out/f_prof.log | tail -60
tail -3 gpurun_out/f_conv_bf.log gpurun_out/f_suite.log
