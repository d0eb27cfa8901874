mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or tiny or zero_steps" 2>&1 | tail -n 6 > gpurun_out/test.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time.log
timeout 900 python tools/dev_chain.py 1000 64 tc3f16 > gpurun_out/chain.log 2>&1
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part" | head -n 10 > gpurun_out/tl.log
cat gpurun_out/test.log gpurun_out/time.log gpurun_out/chain.log gpurun_out/tl.log
