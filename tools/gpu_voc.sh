#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/voc_small.py 20 > gpurun_out/r2e_voc_small.txt 2>&1; echo "small rc=$?"; grep -v Warning gpurun_out/r2e_voc_small.txt | tail -n 3
timeout 600 python tools/dev_voc.py > gpurun_out/r2f_voc.txt 2>&1; echo "voc rc=$?"
grep -v Warning gpurun_out/r2f_voc.txt | tail -n 20
( timeout 600 python -m pytest -q -m gpu tests/test_tc_pair.py tests/test_gpu_parity.py -k "vocoder or nsf or hifigan" ) > gpurun_out/r2f_voc_tests.log 2>&1; echo "voc tests rc=$?"; tail -n 5 gpurun_out/r2f_voc_tests.log
