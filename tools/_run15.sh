mkdir -p gpurun_out; rm -f gpurun_out/*.log
DSVC_TC_BN=256 timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or tiny or zero_steps" 2>&1 | tail -n 6 > gpurun_out/test_256.log
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or tiny or zero_steps" 2>&1 | tail -n 3 > gpurun_out/test_auto.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time.log
DSVC_TC_BN=256 DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part" | tail -n 10 > gpurun_out/tl.log
cat gpurun_out/test_256.log gpurun_out/test_auto.log gpurun_out/time.log gpurun_out/tl.log
