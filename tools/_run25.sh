mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q -s -k "nsf or hifigan or vocoder" 2>&1 | grep -v "^$" | tail -n 12 > gpurun_out/test_voc.log
timeout 600 python tools/dev_voc.py > gpurun_out/voc_time.log 2>&1
DSVC_NSF_SMALLCONV=0 timeout 600 python tools/dev_voc.py 2>&1 | grep "tc B" > gpurun_out/voc_time_nosmall.log
cat gpurun_out/test_voc.log gpurun_out/voc_time.log; echo NOSMALL; cat gpurun_out/voc_time_nosmall.log
