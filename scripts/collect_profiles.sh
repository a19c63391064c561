#!/bin/bash
# Round-2 evidence pass (run under gpurun): full bench line, ncu launch lists (slot path = replayed CUDA graphs, and
# eager), DRAM traffic of the conv launches for bench.py's roofline.traffic, sanitizers on the small step.
OUT=gpurun_out
TAG=${TAG:-r2}
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
# (1) launch list of the path `value` is measured on: 4 pipeline slots, replayed graphs (kernel NODES profiled one by one)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_slots_ncu.csv \
    python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_slots.out 2>&1
echo "ncu slots rc=$?"
# (2) eager launch list + DRAM bytes per launch (cold cache, serialised): one step after 3 warm-up steps
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 60 --csv \
    --log-file $OUT/${TAG}_launches_eager_traffic.csv python bench.py --eager --pipeline 1 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_eager.out 2>&1
echo "ncu eager rc=$?"
# (3) sanitizers on one small step (eager + replayed graph + pipeline slot)
SAN_TESTS=0 bash scripts/sanitize.sh > $OUT/${TAG}_sanitize.out 2>&1
tail -12 $OUT/${TAG}_sanitize.out
