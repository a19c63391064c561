#!/bin/bash
# A/B of the stage-ending conv2 epilogue (MODE 4, SAR_TC_MODE4=0 restores the generic MODE -1) and a pipeline-depth sweep.
OUT=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $OUT/ab_gputest.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/ab_gputest.log
pick='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("value %.0f ms %.4f single %.4f conv_ms %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d.get("single_stream",{}).get("ms_per_step",0), r["conv_ms_per_step"], r["frac"], d["e2e"]["value"]))'
for rep in 1 2; do
for m in 0 1; do
  echo "== MODE4=$m B=64 rep $rep";  SAR_TC_MODE4=$m timeout 300 python bench.py --quick --no-cpu-baseline 2>/dev/null | python -c "$pick"
done; done
for m in 0 1; do
  echo "== MODE4=$m B=512"; SAR_TC_MODE4=$m timeout 300 python bench.py --quick --no-cpu-baseline --batch 512 2>/dev/null | python -c "$pick"
  echo "== MODE4=$m cfg5";  SAR_TC_MODE4=$m timeout 300 python bench.py --quick --no-cpu-baseline --config cfg5 2>/dev/null | python -c "$pick"
done
for d in 3 6 8; do
  echo "== pipeline $d B=64"; timeout 300 python bench.py --quick --no-cpu-baseline --pipeline $d 2>/dev/null | python -c "$pick"
done
