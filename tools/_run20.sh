mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or tiny or zero_steps or plms" 2>&1 | tail -n 6 > gpurun_out/test.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time.log
DSVC_DATAFLOW=0 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_nodf.log
timeout 900 python tools/dev_chain.py 1000 64 tc3f16 > gpurun_out/chain.log 2>&1
cat gpurun_out/test.log; echo DF; cat gpurun_out/time.log; echo NODF; cat gpurun_out/time_nodf.log; cat gpurun_out/chain.log
