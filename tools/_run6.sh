mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "eval_full or ddpm_chain_full or ragged or batch_composition or plms_chain" > gpurun_out/test_q.log 2>&1; echo "q rc=$?" > gpurun_out/rc.txt
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A5 "timeline" | head -30 > gpurun_out/tl.log
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_tc.log 2>&1
DSVC_NO_PDL=1 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep ddpm > gpurun_out/time_nopdl.log
DSVC_TC_BN=128 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_bn128.log
DSVC_TC_BN=64 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_bn64.log
cat gpurun_out/rc.txt; tail -n 4 gpurun_out/test_q.log; cat gpurun_out/tl.log; echo AUTO; cat gpurun_out/time_tc.log; echo NOPDL; cat gpurun_out/time_nopdl.log; echo BN128;  cat gpurun_out/time_bn128.log; echo BN64; cat gpurun_out/time_bn64.log
