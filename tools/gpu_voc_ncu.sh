#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_voc_r2f.csv python tools/voc_once.py > gpurun_out/launches_voc.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_narrow_kernel --launch-skip 60 --launch-count 2 -f -o gpurun_out/prof_r2f_narrow python tools/voc_once.py > gpurun_out/ncu_narrow.log 2>&1; echo "ncu narrow rc=$?"
tail -n 2 gpurun_out/ncu_narrow.log
