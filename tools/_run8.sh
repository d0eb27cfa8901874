mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "eval_full or ddpm_chain_full or ragged or batch_composition" > gpurun_out/test_q.log 2>&1; echo "halo tests rc=$?" > gpurun_out/rc.txt
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_halo.log 2>&1
DSVC_CONV_HALO=0 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_nohalo.log
DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A4 "timeline part 0" | head -12 > gpurun_out/tl.log
cat gpurun_out/rc.txt; tail -n 12 gpurun_out/test_q.log; echo HALO; cat gpurun_out/time_halo.log; echo NOHALO; cat gpurun_out/time_nohalo.log; cat gpurun_out/tl.log
