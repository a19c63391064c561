#!/bin/bash
# ncu captures of the residual-block conv kernels launched layer by layer (eager, no stage chains), warm L2
# (--cache-control none): (1) a light section set over every slab/generic conv launch of one step, (2) --set full with
# source for the stage-1 / stage-2 / stage-3 / stage-4 conv2 kernels and the tensor-core VLAD kernel.
# Output: gpurun_out/r2_conv_sections.csv, gpurun_out/r2_conv_full.ncu-rep (+ csv pages extracted here).
export SAR_CHAIN_STAGES="0"     # no stage is chained (ncu drops empty-valued variables)
B=${PROF_B:-64}
CMD="python bench.py --eager --pipeline 1 --steps 1 --warmup 3 --no-cpu-baseline --batch $B"
ncu --cache-control none --clock-control none \
    --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
    -k regex:"conv_tc_|vlad_tc|stem_tc|bigru_tc" -s 120 -c 44 --csv --page raw --log-file gpurun_out/r2_conv_sections_b$B.csv $CMD > gpurun_out/r2_conv_sections_b$B.out 2>&1
echo "sections rc=$?"
ncu --set full --cache-control none --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"slab_kernel<.int.32, .bool.1, .int.3>|slab_kernel<.int.64, .bool.0, .int.3>|vlad_tc_kernel" -s 27 -c 9 -f -o gpurun_out/r2_conv_full_b$B $CMD > gpurun_out/r2_conv_full_b$B.out 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
