mkdir -p gpurun_out; rm -f gpurun_out/*.log
for cfg in "DSVC_CONV_HALO=0" "DSVC_BO_MODE=1" "DSVC_BO_MODE=0"; do
  echo "== $cfg" >> gpurun_out/halo.log
  env $cfg timeout 300 python -m pytest tests -m gpu -q -k "eval_full and tc3f16" 2>&1 | grep -E "passed|failed|AssertionError: \(" >> gpurun_out/halo.log
done
DSVC_CONV_HALO=0 timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_nohalo.log
DSVC_CONV_HALO=0 timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition" 2>&1 | tail -n 3 > gpurun_out/test_nohalo.log
DSVC_CONV_HALO=0 DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A3 "timeline part" | head -20 > gpurun_out/tl.log
timeout 300 python tools/dev_time.py tc3f16 2>&1 | grep -A2 ddpm > gpurun_out/time_halo.log
cat gpurun_out/halo.log; echo NOHALO; cat gpurun_out/time_nohalo.log gpurun_out/test_nohalo.log; echo HALO; cat gpurun_out/time_halo.log; cat gpurun_out/tl.log
