#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/p2_*.log gpurun_out/p2_*.json
L=diffsvc_b200/lib
( DSVC_LIB=$L/libdsvc_wd.so timeout 900 python -m pytest tests/test_tc_pair.py -x -q -s -m gpu ) > gpurun_out/p2_tests_wd.log 2>&1; rc=$?
echo "pair tests (watchdog build) rc=$rc" > gpurun_out/p2_rc.txt
grep -v Warning gpurun_out/p2_tests_wd.log | tail -n 25
if [ $rc -eq 0 ]; then
  ( timeout 600 python tools/dev_pair.py ) > gpurun_out/p2_ab.log 2>&1; echo "ab rc=$?" >> gpurun_out/p2_rc.txt; cat gpurun_out/p2_ab.log
  ( time timeout 1800 python -m pytest tests -q -m gpu ) > gpurun_out/p2_tests.log 2>&1; echo "all gpu tests rc=$?" >> gpurun_out/p2_rc.txt
  grep -v Warning gpurun_out/p2_tests.log | tail -n 30
  timeout 900 python bench.py > gpurun_out/p2_bench_n1.json 2> gpurun_out/p2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/p2_rc.txt
  cat gpurun_out/p2_bench_n1.json
fi
cat gpurun_out/p2_rc.txt
