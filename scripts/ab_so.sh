#!/bin/bash
# same-box A/B of two builds of the library: aesrc2020_b200/csrc/ab_<name>.so are copied over libsarnet_sm100.so in turn
# usage: bash scripts/ab_so.sh nameA nameB [bench args...]
A=$1; B=$2; shift 2
L=aesrc2020_b200/csrc/libsarnet_sm100.so
pick='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; print("value %.0f ms %.4f single %.4f conv_ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d.get("single_stream",{}).get("ms_per_step",0), r["conv_ms_per_step"], r["frac"]))'
for rep in 1 2 3; do for v in $A $B; do
  cp aesrc2020_b200/csrc/ab_$v.so $L
  echo -n "$v rep $rep $@: "; timeout 300 python bench.py --quick --no-cpu-baseline "$@" 2>/dev/null | python -c "$pick"
done; done
