#!/bin/bash
# GPU visit 2: corr v4 knobs + ncu capture, (kb,tap) split-K A/B, full suite, bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/round2.log
echo "== targeted tests" > $L
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_conv_tc_gpu.py -x -q -m gpu >> $L 2>&1
echo "rc=$?" >> $L
for v in "" "MS_CORR4_TW=64" "MS_CORR4_TW=32" "MS_CORR4_TW=64 MS_CORR4_SLACK=16" "MS_CORR4_TW=32 MS_CORR4_SLACK=8" "MS_CORR4_SO=0" "MS_CORR4_TW=64 MS_CORR4_LP=1"; do
  echo "== corr bench [$v]" >> $L
  env $v timeout 120 python scripts/corr_bench.py 2>&1 | grep -v DispNet >> $L
done
echo "== ncu corr v4" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:corr_fwd4 -c 1 -s 3 -o $O/prof_corr4_r1 -f python scripts/corr_one.py >> $L 2>&1
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== bench (defaults)" >> $L
timeout 600 python bench.py > $O/bench_r2.json 2>> $L
echo "== bench (MS_TC_TAPSPLIT=0)" >> $L
MS_TC_TAPSPLIT=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_r2_notapsplit.json 2>> $L
python - <<'PY' >> $L 2>&1
import json
for f in ('bench_r2.json', 'bench_r2_notapsplit.json'):
    j = json.load(open('gpurun_out/' + f)); print(f, j['value'], j['ms_per_step'], j['e2e']['value'], j['profile_ms_per_step'], j['corr_kernel']['large'])
PY
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -70
