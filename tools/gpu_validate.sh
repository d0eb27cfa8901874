#!/bin/bash
# Round-end validation on a GPU box (run through gpurun from the repo root): GPU parity tests, smoke, the bench
# line, the dominant kernel's ncu capture.  Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/bench_*.json
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/rc.txt
# the dominant kernel's `ncu --set full` capture and the launch list of this state: tools/gpu_ncu.sh / the commands in profiles/README.md
cat gpurun_out/rc.txt; tail -n 6 gpurun_out/test_gpu.log; tail -n 3 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_ref.json
