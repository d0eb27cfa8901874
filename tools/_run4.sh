mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 600 python -m pytest tests -m gpu -q -k "eval_full or ddpm_chain_full or test_ddpm_golden or test_diffnet_golden or nsf_golden or plms_golden" > gpurun_out/test_q.log 2>&1; echo "q rc=$?" > gpurun_out/rc.txt
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_tc.log 2>&1
cat gpurun_out/rc.txt; tail -n 4 gpurun_out/test_q.log; cat gpurun_out/time_tc.log
