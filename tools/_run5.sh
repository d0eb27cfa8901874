mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "eval_full or ddpm_chain_full or ragged or batch_composition" > gpurun_out/test_q.log 2>&1; echo "q rc=$?" > gpurun_out/rc.txt
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A6 "timeline" | head -40 > gpurun_out/tl.log
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_tc.log 2>&1
cat gpurun_out/rc.txt; tail -n 4 gpurun_out/test_q.log; cat gpurun_out/tl.log gpurun_out/time_tc.log
