#!/bin/bash
# Round-end validation on a GPU box (run through gpurun from the repo root): GPU parity tests, smoke, the bench
# line, the dominant kernel's ncu capture.  Outputs land in gpurun_out/ (scratch); copy what should be judged into profiles/.
mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/bench_*.json
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/rc.txt
# ncu (default cache control: caches flushed before each replay pass, so dram bytes are the kernel's own traffic)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair_kernel --launch-skip 45 --launch-count 2 -f -o gpurun_out/prof_r2_b1_cold python tools/dev_prof.py 1 > gpurun_out/ncu_b1_cold.log 2>&1; echo "ncu rc=$?" >> gpurun_out/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 20 --no-extras > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -n 6 gpurun_out/test_gpu.log; tail -n 3 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_ref.json
