mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or tiny or zero_steps" 2>&1 | tail -n 6 > gpurun_out/test_mc.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_mc.log
DSVC_TC_CLUSTER=2 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_mc2.log
DSVC_TC_CLUSTER=1 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_nomc.log
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part" | head -10 > gpurun_out/tl.log
cat gpurun_out/test_mc.log; echo AUTO; cat gpurun_out/time_mc.log; echo CS2; cat gpurun_out/time_mc2.log; echo CS1; cat gpurun_out/time_nomc.log; cat gpurun_out/tl.log
