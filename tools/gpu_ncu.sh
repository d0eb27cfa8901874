#!/bin/bash
# ncu --set full of the two per-layer kernels mid-step: one clip (configs[1]) and the packed cfg3 batch
mkdir -p gpurun_out
export DSVC_STEP=0
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_pair_kernel --launch-skip 45 --launch-count 2 -f -o gpurun_out/prof_r2_b1 python tools/dev_prof.py 1 > gpurun_out/ncu_b1.log 2>&1; echo "ncu b1 rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_pair_kernel --launch-skip 45 --launch-count 2 -f -o gpurun_out/prof_r2_b8 python tools/dev_prof.py 8 > gpurun_out/ncu_b8.log 2>&1; echo "ncu b8 rc=$?"
tail -n 3 gpurun_out/ncu_b1.log gpurun_out/ncu_b8.log; ls -la gpurun_out/*.ncu-rep
