#!/bin/bash
# only the --set full pass of profile_conv.sh
export SAR_CHAIN_STAGES="0"
B=${PROF_B:-64}
CMD="python bench.py --eager --pipeline 1 --steps 1 --warmup 3 --no-cpu-baseline --batch $B"
ncu --set full --cache-control none --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"slab_kernel<.int.32, .bool.1, .int.3>|slab_kernel<.int.64, .bool.0, .int.3>|vlad_tc_kernel" -s 27 -c 9 -f -o gpurun_out/r2_conv_full_b$B $CMD > gpurun_out/r2_conv_full_b$B.out 2>&1
echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
