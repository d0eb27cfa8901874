mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.json
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" > gpurun_out/rc.txt
timeout 600 python tools/latency.py > gpurun_out/latency.json 2> gpurun_out/latency.err; echo "latency rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests -m gpu -q -k "cond_encoder or mel_ or compact or pitch or kernels_match" 2>&1 | tail -n 3 > gpurun_out/test_q.log
cat gpurun_out/rc.txt gpurun_out/test_q.log; cut -c1-250 gpurun_out/bench_n1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e'], d['roofline']['achieved'], d['clocks'])"; cat gpurun_out/latency.json
