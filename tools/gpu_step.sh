#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dev_step.py > gpurun_out/r2d_step.txt 2>&1; echo "step rc=$?"
grep -v Warning gpurun_out/r2d_step.txt | tail -n 30
DSVC_LIB=diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_step_tl.py 689 8 > gpurun_out/r2d_step_tl_b8.txt 2>&1
grep -v Warning gpurun_out/r2d_step_tl_b8.txt | grep -E "cta   0|cta  66|mean"
