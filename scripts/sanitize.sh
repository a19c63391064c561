#!/bin/bash
# compute-sanitizer (memcheck / racecheck / synccheck / initcheck) over the tensor-core conv tests and one small step
# (eager, replayed CUDA graph, pipeline slot), plus ncu's graph-node profiling of the same step (ncu_rc=9).
# Run under gpurun from the repo root; logs land in gpurun_out/sanitize/.
OUT=gpurun_out/sanitize
mkdir -p $OUT
TOOLS=${SAN_TOOLS:-"memcheck synccheck racecheck"}
for tool in $TOOLS; do
  timeout 900 compute-sanitizer --tool $tool --log-file $OUT/${tool}_step.log --error-exitcode 3 \
      python scripts/sanitize_step.py > $OUT/${tool}_step.out 2>&1
  echo "$tool step rc=$?" | tee -a $OUT/summary.txt
  if [ "${SAN_TESTS:-1}" = "1" ]; then
    timeout 1200 compute-sanitizer --tool $tool --log-file $OUT/${tool}_convtests.log --error-exitcode 3 \
        python -m pytest tests/test_gpu_conv_tc.py -x -q -m gpu > $OUT/${tool}_convtests.out 2>&1
    echo "$tool conv tests rc=$?" | tee -a $OUT/summary.txt
  fi
done
for f in $OUT/*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Barrier error" $f | sort | uniq -c | head -12; done | tee -a $OUT/summary.txt
# ncu over the graphed step: every kernel NODE of the replayed graphs must profile
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/ncu_graph_step.csv \
    python scripts/sanitize_step.py > $OUT/ncu_graph_step.out 2>&1
echo "ncu graph step rc=$?" | tee -a $OUT/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/ncu_smoke.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > $OUT/ncu_smoke.out 2>&1
echo "ncu smoke rc=$?" | tee -a $OUT/summary.txt
