#!/bin/bash
# GPU visit 4: leaner tcgen05 splitters (truncation split, no profiling code, incremental tap counters, fast epilogue).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/round4.log
echo "== targeted tests" > $L
timeout 300 python -m pytest tests/test_conv_tc_gpu.py tests/test_ops_gpu.py -x -q -m gpu 2>&1 | tail -4 >> $L
echo "== accuracy" >> $L
timeout 120 python scripts/tc_accuracy.py >> $L 2>&1
echo "== tc_bench" >> $L
timeout 200 python scripts/tc_bench.py >> $L 2>&1
echo "== per-role cycle counters (PROF build)" >> $L
timeout 60 python scripts/tc_prof.py >> $L 2>&1
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== bench (defaults)" >> $L
timeout 600 python bench.py > $O/bench_r4.json 2>> $L
python - <<'PY' >> $L 2>&1
import json
j = json.load(open('gpurun_out/bench_r4.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['profile_ms_per_step'], j['roofline']['achieved'])
PY
echo "== launch list" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_r4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -60
