#!/bin/bash
# visit x: does leaving half of the shared memory free (so the next kernel's CTAs can become resident early) shorten the step?
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
for kb in 220 110 72; do
  echo "MS_BF_SMEM_KB=$kb" >> gpurun_out/x_chain.log
  MS_BF_SMEM_KB=$kb timeout -s KILL 120 python scripts/chain_bench.py >> gpurun_out/x_chain.log 2>&1
  MS_BF_SMEM_KB=$kb timeout -s KILL 200 python bench.py --config 3 --steps 40 --warmup 8 --no-corr-shapes --no-parity-check > gpurun_out/x_bench_$kb.log 2>&1
  echo "MS_BF_SMEM_KB=$kb cfg3: $(tail -n 1 gpurun_out/x_bench_$kb.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"],1), "FPS", round(j["ms_per_step"],3), "ms")' 2>&1)" | tee -a gpurun_out/x_chain.log
done
cat gpurun_out/x_chain.log
MS_BF_SMEM_KB=110 timeout -s KILL 600 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/x_suite_110.log 2>&1
echo "suite(110) rc=$?" >> gpurun_out/x_suite_110.log
tail -n 3 gpurun_out/x_suite_110.log
