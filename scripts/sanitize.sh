#!/bin/bash
# compute-sanitizer over the operator-level GPU tests at small shapes (all kernel generations, split-K on and off).
# usage (on the GPU box): bash scripts/sanitize.sh   -> gpurun_out/sanitizer_{memcheck,racecheck}.log (+ .pytest.log)
mkdir -p gpurun_out
SEL_BF='case0] or case1] or case2] or case3] or case4] or case6] or case7] or case8] or case9] or case10] or case16] or case19] or case20]'
for TOOL in memcheck racecheck; do
  timeout -s KILL 1500 compute-sanitizer --tool $TOOL --log-file gpurun_out/sanitizer_$TOOL.log \
    python -m pytest -q -p no:cacheprovider --timeout 1200 \
      tests/test_conv_bf_gpu.py -k "$SEL_BF" > gpurun_out/sanitizer_$TOOL.bf.pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$TOOL.bf.pytest.log
  timeout -s KILL 1500 compute-sanitizer --tool $TOOL --log-file gpurun_out/sanitizer_$TOOL.ops.log \
    python -m pytest -q -p no:cacheprovider --timeout 1200 \
      tests/test_ops_gpu.py tests/test_conv_tc_gpu.py -k "not speed and not 1280 and not large" > gpurun_out/sanitizer_$TOOL.ops.pytest.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_$TOOL.ops.pytest.log
done
MS_SMOKE=1 timeout -s KILL 1200 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck.engine.log \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.engine.out 2>&1
tail -2 gpurun_out/sanitizer_*.log | cut -c1-200
