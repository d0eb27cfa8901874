mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition" 2>&1 | tail -n 5 > gpurun_out/test_q.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time.log
DSVC_CONV_HALO=1 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_halo.log
DSVC_CONV_HALO=1 timeout 300 python -m pytest tests -m gpu -q -k "eval_full and tc3f16" 2>&1 | tail -n 2 >> gpurun_out/time_halo.log
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part" | head -10 > gpurun_out/tl.log
timeout 900 python tools/dev_chain.py 1000 64 tc3f16 > gpurun_out/chain.log 2>&1
cat gpurun_out/test_q.log; cat gpurun_out/time.log; echo HALO; cat gpurun_out/time_halo.log; cat gpurun_out/tl.log gpurun_out/chain.log
