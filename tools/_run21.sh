mkdir -p gpurun_out; rm -f gpurun_out/*.log
timeout 900 python -m pytest tests/test_mel_analysis.py tests/test_pitch_extractor.py -m gpu -q -s 2>&1 | tail -n 40 > gpurun_out/test_new.log
cat gpurun_out/test_new.log
