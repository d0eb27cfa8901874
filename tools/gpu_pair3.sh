#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/p3_*.log gpurun_out/p3_*.json
L=diffsvc_b200/lib
( DSVC_LIB=$L/libdsvc_wd.so timeout 900 python -m pytest tests/test_tc_pair.py -x -q -m gpu -k vocoder ) > gpurun_out/p3_tests_wd.log 2>&1; rc=$?
echo "vocoder pair tests (watchdog build) rc=$rc" > gpurun_out/p3_rc.txt
grep -v Warning gpurun_out/p3_tests_wd.log | tail -n 15
if [ $rc -eq 0 ]; then
  ( time timeout 1800 python -m pytest tests -q -m gpu ) > gpurun_out/p3_tests.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/p3_rc.txt
  grep -v Warning gpurun_out/p3_tests.log | tail -n 30
  timeout 900 python bench.py > gpurun_out/p3_bench_n1.json 2> gpurun_out/p3_bench_n1.err; echo "bench rc=$?" >> gpurun_out/p3_rc.txt
  cat gpurun_out/p3_bench_n1.json; tail -n 3 gpurun_out/p3_bench_n1.err
  ( DSVC_TL_PARTS=0,1 DSVC_LIB=$L/libdsvc_tl.so timeout 200 python tools/dev_timeline.py ) > gpurun_out/p3_timeline.log 2>&1; echo "timeline rc=$?" >> gpurun_out/p3_rc.txt
  cat gpurun_out/p3_timeline.log
  ( timeout 300 python tools/dev_voc.py ) > gpurun_out/p3_voc.log 2>&1; cat gpurun_out/p3_voc.log
fi
cat gpurun_out/p3_rc.txt
