#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/last.log
echo "== full gpu suite (defaults)" > $L
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== bench (defaults)" >> $L
timeout 300 python bench.py > $O/bench_last.json 2>> $L
echo "== MS_CONV_SMALL16=2: tests" >> $L
MS_CONV_SMALL16=2 timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_madnet_gpu.py -x -q -m gpu 2>&1 | tail -3 >> $L
echo "== MS_CONV_SMALL16=2: bench" >> $L
MS_CONV_SMALL16=2 timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_last_small16.json 2>> $L
python - <<'PY' >> $L 2>&1
import json
for f in ('bench_last.json', 'bench_last_small16.json'):
    j = json.load(open('gpurun_out/' + f)); print(f, j['value'], j['ms_per_step'], j['e2e']['value'], j['profile_ms_per_step'])
PY
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -30
