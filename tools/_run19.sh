mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.json
timeout 600 python tools/eager_baseline.py > gpurun_out/eager.json 2> gpurun_out/eager.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "tiny_and_odd or zero_steps or nsf_golden or hifigan24k or cond_encoder or test_diffnet_golden" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" > gpurun_out/rc.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "tiny_and_odd and tc3f16 and 129 or nsf_golden" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/rc.txt
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/rc.txt gpurun_out/eager.json; tail -n 3 gpurun_out/eager.err; tail -n 6 gpurun_out/memcheck.log; tail -n 6 gpurun_out/racecheck.log; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'])"
