#!/bin/bash
# Final evidence visit: full GPU suite, smoke, bench (both arms), ncu capture of the dominant kernel, launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
L=$O/final.log
echo "== full gpu suite" > $L
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $L
tail -3 $O/pytest_gpu.log >> $L
echo "== smoke" >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
echo "== bench" >> $L
timeout 600 python bench.py > $O/bench_final.json 2>> $L
echo "== bench --impl reference" >> $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $L
echo "== ncu conv_tc_ts 128->128 @96x320" >> $L
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ts -c 1 -s 2 -o $O/prof_tc_ts_final -f python scripts/tc_bench.py 0 2>&1 | tail -2 >> $L
echo "== launch list (first 330 launches)" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $O/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY' >> $L 2>&1
import json
j = json.load(open('gpurun_out/bench_final.json'))
print(j['value'], j['ms_per_step'], j['e2e'], j['profile_ms_per_step'], j['roofline']['achieved'], j['roofline']['dominant_layer'], j['corr_kernel']['large'], j.get('cpu_baseline'), j.get('clocks'))
print(open('gpurun_out/bench_reference.json').read()[:600])
PY
grep -v "^===\|Creation\|Validated\|Meta op\|Network ready" $L | tail -40
