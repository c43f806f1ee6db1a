#!/bin/bash
# GPU visit 5: operand-stage depth of the tcgen05 conv / wgrad kernels (MS_TC_NS, MS_WG_NS), tests, bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/round5.log
echo "== targeted tests (defaults NS=3)" > $L
timeout 300 python -m pytest tests/test_conv_tc_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
for ns in 2 3 4; do
  echo "== tc_bench MS_TC_NS=$ns" >> $L
  for sh in 0 2 4 7; do MS_TC_NS=$ns timeout 60 python scripts/tc_bench.py $sh 2>&1 | tail -1 | cut -c1-75 >> $L; done
done
for ns in 2 3; do
  echo "== wg_bench MS_WG_NS=$ns" >> $L
  MS_WG_NS=$ns timeout 90 python scripts/wg_bench.py >> $L 2>&1
done
echo "== accuracy (NS=3 => one main accumulator at cout=128)" >> $L
timeout 120 python scripts/tc_accuracy.py 2>&1 | grep " tc " >> $L
echo "== full gpu suite" >> $L
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== bench (defaults)" >> $L
timeout 600 python bench.py > $O/bench_r5.json 2>> $L
echo "== bench NS=2 both" >> $L
MS_TC_NS=2 MS_WG_NS=2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_r5_ns2.json 2>> $L
python - <<'PY' >> $L 2>&1
import json
for f in ('bench_r5.json', 'bench_r5_ns2.json'):
    j = json.load(open('gpurun_out/' + f)); print(f, j['value'], j['ms_per_step'], j['e2e']['value'], j['profile_ms_per_step'], j['roofline']['achieved'])
PY
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -60
