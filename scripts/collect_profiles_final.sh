#!/bin/bash
# Final evidence pass of round 2 (run under gpurun from the repo root; ~12 minutes of GPU):
#   (0) GPU test suite, (1) full bench line, (2) ncu launch list of the slot path (the path `value` is measured on),
#   (3) eager launch list + DRAM bytes per launch (-> profiles/r2_conv_traffic.json via scripts/summarize_launches.py),
#   (4) per-conv-launch counters at B=512 (warm L2, no chains: slab / pair / generic kernels one by one),
#   (5) ncu --set full of the pair kernel and the stage-1 slab kernel at B=512, (6) sanitizers (default and forced CTA pairs).
OUT=gpurun_out
TAG=${TAG:-r2f}
mkdir -p $OUT/sanitize
(time timeout 900 python -m pytest tests -m gpu -x -q) > $OUT/${TAG}_gputest.log 2>&1
echo "gpu tests rc=$?"; tail -3 $OUT/${TAG}_gputest.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_slots_ncu.csv \
    python bench.py --quick --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_slots.out 2>&1
echo "ncu slots rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 90 --csv \
    --log-file $OUT/${TAG}_launches_eager_traffic.csv python bench.py --eager --pipeline 1 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches_eager.out 2>&1
echo "ncu eager rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.max,launch__registers_per_thread
SAR_CHAIN_STAGES="0" timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:"conv_tc_" -s 108 -c 36 --csv \
    --log-file $OUT/${TAG}_conv_counters_b512.csv python bench.py --eager --pipeline 1 --steps 1 --warmup 3 --no-cpu-baseline --batch 512 > $OUT/${TAG}_conv_counters_b512.out 2>&1
echo "ncu counters b512 rc=$?"
SAR_CHAIN_STAGES="0" timeout 600 ncu --set full --cache-control none --clock-control none --import-source on \
    -k regex:"conv_tc_pair_kernel" -s 24 -c 2 -f -o $OUT/${TAG}_pair_full_b512 python bench.py --eager --pipeline 1 --steps 1 --warmup 1 --no-cpu-baseline --batch 512 > $OUT/${TAG}_pair_full.out 2>&1
echo "ncu full pair rc=$?"
SAR_CHAIN_STAGES="0" timeout 600 ncu --set full --cache-control none --clock-control none --import-source on \
    --kernel-name-base demangled -k regex:"conv_tc_slab_kernel<32, 1, 0>|conv_tc_slab_kernel<.int.32, .bool.1, .int.0>" -c 1 -f -o $OUT/${TAG}_slab_s1_full_b512 python bench.py --eager --pipeline 1 --steps 1 --warmup 1 --no-cpu-baseline --batch 512 > $OUT/${TAG}_slab_s1_full.out 2>&1
echo "ncu full slab rc=$?"
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 1000 --log-file $OUT/sanitize/${TAG}_${tool}_step.log --error-exitcode 3 python scripts/sanitize_step.py > $OUT/sanitize/${TAG}_${tool}_step.out 2>&1
  echo "$tool step rc=$?"
done
for tool in memcheck synccheck racecheck; do
  SAR_TC_PAIR=2 SAN_B=8 timeout 600 compute-sanitizer --tool $tool --print-limit 1000 --log-file $OUT/sanitize/${TAG}_${tool}_pair.log --error-exitcode 3 python scripts/sanitize_step.py > $OUT/sanitize/${TAG}_${tool}_pair.out 2>&1
  echo "$tool pair-forced step rc=$?"
done
for f in $OUT/sanitize/${TAG}_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Barrier error|    at " $f | sed "s/by thread.*//" | sort | uniq -c | sort -rn | head -8; done
ls -la $OUT/*.ncu-rep
# DRAM bytes per residual-block conv launch of THIS build -> the file bench.py's roofline.traffic reads (copy it to profiles/r2_conv_traffic.json)
python scripts/summarize_launches.py $OUT/${TAG}_launches_eager_traffic.csv --align --traffic-json $OUT/${TAG}_conv_traffic.json \
    --build-id $(python -c "import bench; print(bench.conv_build_id())") --B 64 | tail -3
