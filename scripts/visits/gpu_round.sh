#!/bin/bash
# One GPU-box visit: targeted tests of the new kernels, A/B timings through env switches, full GPU suite, bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
echo "== targeted tests (defaults: corr v4, small conv kernels)" > $O/round.log
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "correlation or cost_volume or conv2d_fwd" >> $O/round.log 2>&1
echo "rc=$?" >> $O/round.log
echo "== corr bench: v4 LP=2 / v4 LP=1 / v4 no TMA store / v3" >> $O/round.log
timeout 120 python scripts/corr_bench.py >> $O/round.log 2>&1
MS_CORR4_LP=1 timeout 120 python scripts/corr_bench.py >> $O/round.log 2>&1
MS_CORR4_ST=0 timeout 120 python scripts/corr_bench.py >> $O/round.log 2>&1
MS_CORR4_TW=64 timeout 120 python scripts/corr_bench.py >> $O/round.log 2>&1
MS_CORR_V=3 timeout 120 python scripts/corr_bench.py >> $O/round.log 2>&1
echo "== full gpu suite" >> $O/round.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/round.log
tail -5 $O/pytest_gpu.log >> $O/round.log
echo "== bench (defaults)" >> $O/round.log
timeout 600 python bench.py > $O/bench_new.json 2>> $O/round.log
echo "== bench (old paths: MS_CORR_V=3 MS_CONV_SMALL=0)" >> $O/round.log
MS_CORR_V=3 MS_CONV_SMALL=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_old.json 2>> $O/round.log
echo "== modes" >> $O/round.log
timeout 300 python scripts/modes_bench.py > $O/modes.json 2>> $O/round.log
cat $O/bench_new.json | head -c 1500
tail -60 $O/round.log
