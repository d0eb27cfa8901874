#!/bin/bash
# One GPU call (run through gpurun from the repo root, after `python tools/build_variants.py`): cycle stamps of the
# per-layer kernels and of the fused layer kernel, A/B step timing (early vs late epilogue loads; two kernels vs fused),
# and the fused kernel's bit-identity tests.  Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out; rm -f gpurun_out/fused_*.log
L=diffsvc_b200/lib
( DSVC_LIB=$L/libdsvc_tl.so timeout 100 python tools/dev_timeline.py ) > gpurun_out/fused_timeline.log 2>&1
echo "timeline rc=$?" > gpurun_out/fused_rc.txt; cat gpurun_out/fused_timeline.log
( timeout 120 python tools/dev_fused.py; DSVC_LIB=$L/libdsvc_late.so timeout 120 python tools/dev_fused.py ) > gpurun_out/fused_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/fused_rc.txt; cat gpurun_out/fused_ab.log
( time timeout 240 python -m pytest tests/test_fused_layer.py -x -q ) > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$?" >> gpurun_out/fused_rc.txt; tail -n 6 gpurun_out/fused_tests.log
cat gpurun_out/fused_rc.txt
