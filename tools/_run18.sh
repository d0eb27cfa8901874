mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.json
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --batch 8 --frames 689 --steps 2 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
timeout 600 python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 20 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiGate -s 45 -c 2 -o gpurun_out/prof_conv_r1d python bench.py --steps 1 --warmup 1 --ddpm-steps 3 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiOutProj -s 45 -c 2 -o gpurun_out/prof_outproj_r1d python bench.py --steps 1 --warmup 1 --ddpm-steps 3 > gpurun_out/ncu3.log 2>&1
cat gpurun_out/rc.txt; tail -n 8 gpurun_out/test_gpu.log; tail -n 1 gpurun_out/smoke.log; cat gpurun_out/latency.json; cut -c1-330 gpurun_out/bench_n1.json; cut -c1-330 gpurun_out/bench_b8.json
