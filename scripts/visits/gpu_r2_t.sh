#!/bin/bash
# visit t: steady-state per-launch cost of conv_bf inside a graph, with kill switches
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
for d in 0 1 8 9 15; do
  echo "MS_BF_DEBUG=$d" >> gpurun_out/t_chain.log
  MS_BF_DEBUG=$d timeout -s KILL 120 python scripts/chain_bench.py >> gpurun_out/t_chain.log 2>&1
done
echo "MS_PDL=0" >> gpurun_out/t_chain.log
MS_PDL=0 timeout -s KILL 120 python scripts/chain_bench.py >> gpurun_out/t_chain.log 2>&1
echo "MS_BF_KSPLIT_MAX=1" >> gpurun_out/t_chain.log
MS_BF_KSPLIT_MAX=1 timeout -s KILL 120 python scripts/chain_bench.py >> gpurun_out/t_chain.log 2>&1
cat gpurun_out/t_chain.log
