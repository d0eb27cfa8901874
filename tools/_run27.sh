mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 300 python tools/dev_e2e.py > gpurun_out/e2e.log 2>&1
timeout 300 python tools/latency.py > gpurun_out/latency2.json 2> gpurun_out/latency2.err
cat gpurun_out/e2e.log gpurun_out/latency2.json
