#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/voc_small.py 20 > gpurun_out/r2e_voc_small.txt 2>&1; echo "small rc=$?"; grep -v Warning gpurun_out/r2e_voc_small.txt | tail -n 3
timeout 600 python tools/dev_voc.py > gpurun_out/r2e_voc.txt 2>&1; echo "voc rc=$?"
grep -v Warning gpurun_out/r2e_voc.txt | tail -n 20
