#!/bin/bash
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "nsf_tensor_core or nsf_full or tiny_and_odd or hifigan24k" ) > gpurun_out/last_voc.log 2>&1; echo "voc rc=$?" > gpurun_out/last_rc.txt; tail -n 3 gpurun_out/last_voc.log
( DSVC_TEST_EXPERIMENTS=1 timeout 40 python -m pytest tests/test_fused_layer.py -x -q -k "deterministic or plms" ) > gpurun_out/last_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/last_rc.txt; tail -n 3 gpurun_out/last_fused.log
cat gpurun_out/last_rc.txt
