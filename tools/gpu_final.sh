#!/bin/bash
# last check of a change that touches every wide-tile epilogue: timings, the whole GPU suite, smoke, the bench line
mkdir -p gpurun_out
timeout 300 python - > gpurun_out/r2h_time.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, "tools"); sys.argv = ["x", "tc3f16"]
import dev_time as d
d.time_ddpm("tc3f16", 1, 862, 200); d.time_ddpm("tc3f16", 8, 689, 50); d.time_ddpm("tc3f16", 8, 689, 50); d.time_voc(1, 862); d.time_voc(8, 689)
PY
grep -v Warning gpurun_out/r2h_time.txt | tail -n 12
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -n 4 gpurun_out/test_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
