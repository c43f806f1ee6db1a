#!/bin/bash
# GPU visit 3: optimised corr v4 + generalised small conv kernels; timing experiments on the tcgen05 conv kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/round3.log
echo "== targeted tests" > $L
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -4 >> $L
for v in "" "MS_CORR4_TW=128" "MS_CORR4_SLACK=32" "MS_CORR4_TW=32 MS_CORR4_SLACK=8" "MS_CORR4_ST=0"; do
  echo "== corr bench [$v]" >> $L
  env $v timeout 120 python scripts/corr_bench.py 2>&1 | grep -v DispNet >> $L
done
echo "== conv_tc_ts kill-switch timing, 128->128 3x3 @96x320 (16 = no A split, 32 = no B split, 64 = no MMA)" >> $L
for d in 0 16 32 48 64 80 96 112; do
  echo "-- MS_TC_DEBUG=$d" >> $L
  MS_TC_DEBUG=$d timeout 60 python scripts/tc_bench.py 0 2>&1 | tail -1 >> $L
done
echo "== per-role cycle counters" >> $L
timeout 60 python scripts/tc_prof.py >> $L 2>&1
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== bench (defaults)" >> $L
timeout 600 python bench.py > $O/bench_r3.json 2>> $L
echo "== ncu corr v4" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_fwd4 -c 1 -s 3 -o $O/prof_corr4b_r1 -f python scripts/corr_one.py 2>&1 | tail -2 >> $L
python - <<'PY' >> $L 2>&1
import json
j = json.load(open('gpurun_out/bench_r3.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['profile_ms_per_step'], j['corr_kernel']['large'])
PY
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -75
