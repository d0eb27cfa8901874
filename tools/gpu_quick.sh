#!/bin/bash
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2g_unroll.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, "tools"); sys.argv = ["x", "tc3f16"]
import dev_time as d
d.time_ddpm("tc3f16", 1, 862, 200); d.time_ddpm("tc3f16", 1, 862, 200); d.time_ddpm("tc3f16", 1, 43, 200); d.time_ddpm("tc3f16", 8, 689, 50)
PY
grep -v Warning gpurun_out/r2g_unroll.txt | tail -n 14
( timeout 600 python -m pytest -q -m gpu tests/test_gpu_parity.py -k "ddpm or philox or zero_steps" ) > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/r2g_tests.log
