mkdir -p gpurun_out; rm -f gpurun_out/*.log
DSVC_CONV_HALO=1 timeout 300 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition" 2>&1 | tail -n 3 > gpurun_out/test_halo.log
DSVC_CONV_HALO=1 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_halo.log
DSVC_CONV_HALO=1 DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part 0" | head -5 > gpurun_out/tl.log
cat gpurun_out/test_halo.log gpurun_out/time_halo.log gpurun_out/tl.log
