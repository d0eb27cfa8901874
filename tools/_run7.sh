mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.json
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --batch 8 --frames 689 --steps 2 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench b8 rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -n 6 gpurun_out/test_gpu.log; cat gpurun_out/smoke.log | tail -n 2; cat gpurun_out/bench_n1.json gpurun_out/bench_b8.json gpurun_out/bench_ref.json; tail -n 3 gpurun_out/bench_n1.err
