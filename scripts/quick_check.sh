#!/bin/bash
# GPU test suite + quick bench lines (B=64, B=512, cfg5) -- the check run after a kernel change.
OUT=gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > $OUT/qc_gputest.log 2>&1; echo "gpu tests rc=$?"; tail -4 $OUT/qc_gputest.log
pick='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("value %.0f ms %.4f single %.4f conv_ms %.4f frac %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d.get("single_stream",{}).get("ms_per_step",0), r["conv_ms_per_step"], r["frac"], d["e2e"]["value"]))'
for rep in 1 2; do echo "== B=64 rep $rep"; timeout 300 python bench.py --quick --no-cpu-baseline 2>/dev/null | python -c "$pick"; done
echo "== B=512"; timeout 300 python bench.py --quick --no-cpu-baseline --batch 512 2>/dev/null | python -c "$pick"
echo "== cfg5";  timeout 300 python bench.py --quick --no-cpu-baseline --config cfg5 2>/dev/null | python -c "$pick"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
