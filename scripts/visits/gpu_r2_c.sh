#!/bin/bash
# round-2 third visit: conv_bf v2 (fp16 forward planes, 64-channel K blocks, bulk weights) -- op tests, mixed-format probe,
# micro-benchmarks, full suite, bench, ncu of the dominant layer, launch list
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -1
rm -f gpurun_out/conv_bf_errors.jsonl gpurun_out/baseline_parity.jsonl
timeout -s KILL 900 python -m pytest tests/test_conv_bf_gpu.py -q --timeout 180 > gpurun_out/c_conv_bf.log 2>&1
echo "conv_bf rc=$?" >> gpurun_out/c_conv_bf.log
timeout -s KILL 300 python scripts/wgrad_mixed_probe.py > gpurun_out/c_mixed_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/c_mixed_probe.log
timeout -s KILL 300 python scripts/bf_bench.py > gpurun_out/c_bf_bench.log 2>&1
timeout -s KILL 2400 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_conv_bf_gpu.py > gpurun_out/c_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/c_suite.log
timeout -s KILL 900 python bench.py --steps 30 --warmup 5 > gpurun_out/c_bench.log 2>&1
echo "bench rc=$?" >> gpurun_out/c_bench.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:conv_bf_kernel -c 1 -o gpurun_out/c_ncu_conv_bf_128 python scripts/bf_bench.py one 0 > gpurun_out/c_ncu1.log 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity-check --no-corr-shapes > gpurun_out/c_launch_bench.log 2>&1
tail -3 gpurun_out/c_conv_bf.log gpurun_out/c_mixed_probe.log gpurun_out/c_suite.log gpurun_out/c_bench.log
