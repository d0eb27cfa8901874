mkdir -p gpurun_out; rm -f gpurun_out/*.log
for mode in tc1f16 tc3f16; do for bn in 64 128; do
  echo "== mode=$mode BN=$bn" >> gpurun_out/micro.log
  DSVC_TC_BN=$bn DSVC_LIB=$PWD/diffsvc_b200/lib/libdsvc_tl.so timeout 300 python tools/dev_time.py $mode 2>&1 | grep -A2 "timeline part" | head -6 >> gpurun_out/micro.log
done; done
cat gpurun_out/micro.log
