#!/bin/bash
# One GPU call: (1) bit-identity tests of the fused layer kernel, (2) A/B timing, (3) a default-path subset + smoke.
mkdir -p gpurun_out; rm -f gpurun_out/fused_*.log
( time DSVC_TEST_EXPERIMENTS=1 timeout 240 python -m pytest tests/test_fused_layer.py -x -q ) > gpurun_out/fused_tests.log 2>&1
echo "fused tests rc=$?" > gpurun_out/fused_rc.txt; tail -n 15 gpurun_out/fused_tests.log
( time timeout 150 python tools/dev_fused.py ) > gpurun_out/fused_ab.log 2>&1
echo "ab rc=$?" >> gpurun_out/fused_rc.txt; cat gpurun_out/fused_ab.log
( time timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or chain_full or composition or ragged" ) > gpurun_out/fused_default_subset.log 2>&1
echo "default subset rc=$?" >> gpurun_out/fused_rc.txt; tail -n 5 gpurun_out/fused_default_subset.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/fused_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/fused_rc.txt; tail -n 2 gpurun_out/fused_smoke.log
cat gpurun_out/fused_rc.txt
