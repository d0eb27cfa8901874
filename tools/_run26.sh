mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.json gpurun_out/*.csv gpurun_out/*.md
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/test_gpu.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/rc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py --batch 8 --frames 689 --steps 2 --warmup 3 > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
timeout 600 python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err; echo "latency rc=$?" >> gpurun_out/rc.txt
DSVC_NSF_TILE8=1 timeout 300 python tools/dev_voc.py 2>&1 | grep "tc B" > gpurun_out/voc_tile8.log
timeout 300 python tools/dev_voc.py > gpurun_out/voc_time.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 20 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiVoc -s 40 -c 3 -o gpurun_out/prof_voc_r1e python tools/voc_once.py > gpurun_out/ncu2.log 2>&1
cat gpurun_out/rc.txt; tail -n 6 gpurun_out/test_gpu.log; tail -n 3 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_b8.json; cat gpurun_out/latency.json; tail -n 2 gpurun_out/latency.err; echo TILE8; cat gpurun_out/voc_tile8.log; cat gpurun_out/voc_time.log; ls -la gpurun_out
