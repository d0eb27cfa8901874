mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -x -k "eval_full and tc3f16" 2>&1 | tail -n 8 > gpurun_out/test1.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -n 8 > gpurun_out/test.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time.log
DSVC_PERSISTENT=0 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm | head -3 > gpurun_out/time_np.log
timeout 300 python tools/latency.py > gpurun_out/latency.json 2>gpurun_out/latency.err
cat gpurun_out/test1.log gpurun_out/test.log gpurun_out/time.log gpurun_out/time_np.log gpurun_out/latency.json; tail -n 3 gpurun_out/latency.err
