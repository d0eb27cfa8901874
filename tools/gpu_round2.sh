#!/bin/bash
# Round-2 validation on a GPU box (through gpurun): the new parity / Svc tests, then both bench arms.
mkdir -p gpurun_out; rm -f gpurun_out/r2_*.log gpurun_out/r2_*.json
( time timeout 1500 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_svc_infer_gpu.py -q -s -m gpu ) > gpurun_out/r2_newtests.log 2>&1; echo "new tests rc=$?" > gpurun_out/r2_rc.txt
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?" >> gpurun_out/r2_rc.txt
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/r2_rc.txt
cat gpurun_out/r2_rc.txt; grep -v Warning gpurun_out/r2_newtests.log | tail -n 40; cat gpurun_out/r2_bench_n1.json; tail -n 5 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_ref.json
