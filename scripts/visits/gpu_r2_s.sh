#!/bin/bash
# visit s: where the remaining time goes after the uniform issue loops (CTA clocks), heuristics A/B, corrected MMA probe
mkdir -p gpurun_out
make -C real-time-self-adaptive-deep-stereo_b200/csrc -j16 2>&1 | tail -n 1
timeout -s KILL 300 python scripts/mma_probe.py > gpurun_out/mma_probe2.log 2>&1
tail -n 66 gpurun_out/mma_probe2.log
for i in 0 6 7; do MS_BF_PROF=1 timeout -s KILL 120 python scripts/bf_bench.py prof $i >> gpurun_out/s_cta_clock.log 2>&1; done
cat gpurun_out/s_cta_clock.log
ab() { # name, env assignments..., config
  local name=$1; shift; local cfg=$1; shift
  env "$@" timeout -s KILL 300 python bench.py --config $cfg --steps 40 --warmup 8 --no-corr-shapes --no-parity-check > gpurun_out/s_ab_${name}.log 2>&1
  echo "$name: $(tail -n 1 gpurun_out/s_ab_${name}.log | python -c 'import sys,json; j=json.loads(sys.stdin.read()); print(round(j["value"],1), "FPS", round(j["ms_per_step"],3), "ms")' 2>&1)"
}
ab cfg3_base 3 MS_NOP=1
ab cfg3_ksplit4 3 MS_BF_KSPLIT_MAX=4
ab cfg3_ksplit16 3 MS_BF_KSPLIT_MAX=16
ab cfg3_splitcyc4096 3 MS_BF_SPLIT_CYCLES=4096
ab cfg3_nooverlap 3 MS_WGRAD_OVERLAP=0
ab cfg3_nopdl 3 MS_PDL=0
ab cfg3_notilesearch 3 MS_BF_TILE_SEARCH=0
ab cfg4_base 4 MS_NOP=1
ab cfg4_nogroups 4 MS_WB_GROUPS=0
ab cfg4_nostem 4 MS_STEM=0
ab cfg2_base 2 MS_NOP=1
ab cfg2_nooverlap 2 MS_WGRAD_OVERLAP=0
