#!/bin/bash
# round-2 second visit: conv_bf + wgrad_bf op tests, full GPU suite on the split-bf16 engine, benches, ncu of the dominant layer
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 -x > gpurun_out/b_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/b_conv_bf.log
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/b_bf_bench.log 2>&1
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/b_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/b_suite.log
timeout -s KILL 600 python bench.py --steps 30 --warmup 5 > gpurun_out/b_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/b_bench.log
MS_BF_WGRAD=0 timeout -s KILL 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/b_bench_nowg.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_bf_kernel -c 1 -o gpurun_out/b_ncu_conv_bf_128 python scripts/bf_bench.py one 0 > gpurun_out/b_ncu1.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_bf_kernel -c 1 -o gpurun_out/b_ncu_wgrad_bf_128 python scripts/bf_bench.py one 0 > gpurun_out/b_ncu2.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_launch_bench.log 2>&1
tail -3 gpurun_out/b_conv_bf.log gpurun_out/b_suite.log gpurun_out/b_bench.log
