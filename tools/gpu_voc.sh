#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/dev_voc.py > gpurun_out/r2e_voc.txt 2>&1; echo "voc rc=$?"
grep -v Warning gpurun_out/r2e_voc.txt | tail -n 20
