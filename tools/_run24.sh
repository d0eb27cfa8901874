mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests -m gpu -q -s -k "nsf or hifigan or vocoder or after_infer" 2>&1 | grep -v "^$" | tail -n 20 > gpurun_out/test_voc.log
timeout 600 python tools/dev_voc.py > gpurun_out/voc_time.log 2>&1
cat gpurun_out/test_voc.log gpurun_out/voc_time.log
