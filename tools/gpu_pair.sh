#!/bin/bash
# First runs of the CTA-pair kernels: parity under the watchdog build (a lost arrive traps instead of hanging), then
# the A/B timing with the product build.  Outputs in gpurun_out/.
mkdir -p gpurun_out; rm -f gpurun_out/pair_*.log
L=diffsvc_b200/lib
( DSVC_LIB=$L/libdsvc_wd.so timeout 900 python -m pytest tests/test_tc_pair.py -x -q -s -m gpu ) > gpurun_out/pair_tests_wd.log 2>&1; rc=$?
echo "pair tests (watchdog build) rc=$rc" > gpurun_out/pair_rc.txt
grep -v Warning gpurun_out/pair_tests_wd.log | tail -n 25
if [ $rc -eq 0 ]; then
  ( timeout 600 python tools/dev_pair.py ) > gpurun_out/pair_ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/pair_rc.txt; cat gpurun_out/pair_ab.log
  ( timeout 900 python -m pytest tests/test_tc_pair.py tests/test_fused_layer.py tests/test_gpu_parity.py -q -m gpu ) > gpurun_out/pair_tests.log 2>&1; echo "product tests rc=$?" >> gpurun_out/pair_rc.txt
  tail -n 8 gpurun_out/pair_tests.log
fi
( timeout 900 python -m pytest tests/test_svc_infer_gpu.py -q -s -m gpu ) > gpurun_out/pair_svc.log 2>&1; echo "svc tests rc=$?" >> gpurun_out/pair_rc.txt
grep -v Warning gpurun_out/pair_svc.log | tail -n 30
cat gpurun_out/pair_rc.txt
