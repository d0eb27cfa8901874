#!/bin/bash
# One GPU call: cycle stamps of the fused layer kernel (+ hoisted epilogue loads), A/B step timing, bit-identity tests.
mkdir -p gpurun_out; rm -f gpurun_out/fused_*.log
L=diffsvc_b200/lib
( DSVC_LIB=$L/libdsvc_tl.so timeout 100 python tools/dev_timeline.py; DSVC_LIB=$L/libdsvc_tlh.so timeout 100 python tools/dev_timeline.py ) > gpurun_out/fused_timeline.log 2>&1
echo "timeline rc=$?" > gpurun_out/fused_rc.txt; cat gpurun_out/fused_timeline.log
( timeout 120 python tools/dev_fused.py; DSVC_LIB=$L/libdsvc_hoist.so timeout 120 python tools/dev_fused.py ) > gpurun_out/fused_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/fused_rc.txt; cat gpurun_out/fused_ab.log
( time DSVC_TEST_EXPERIMENTS=1 timeout 240 python -m pytest tests/test_fused_layer.py -x -q ) > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$?" >> gpurun_out/fused_rc.txt; tail -n 6 gpurun_out/fused_tests.log
( DSVC_LIB=$L/libdsvc_hoist.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "chain_full or composition or nsf_golden" ) > gpurun_out/fused_hoist_subset.log 2>&1
echo "hoist subset rc=$?" >> gpurun_out/fused_rc.txt; tail -n 4 gpurun_out/fused_hoist_subset.log
cat gpurun_out/fused_rc.txt
