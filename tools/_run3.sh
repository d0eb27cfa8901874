mkdir -p gpurun_out; rm -f gpurun_out/*.log
nproc > gpurun_out/cpu.txt; lscpu | grep -E "Model name|Socket|Thread|Core" >> gpurun_out/cpu.txt
timeout 600 python -m pytest tests -m gpu -q -k "tc3f16 or batch_composition or philox or test_ddpm_golden" > gpurun_out/test_tc.log 2>&1; echo "tc rc=$?" > gpurun_out/rc.txt
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_tc.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --ddpm-steps 3 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiGate -s 25 -c 2 -o gpurun_out/prof_conv_r1 python bench.py --steps 1 --warmup 1 --ddpm-steps 3 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiOutProj -s 25 -c 2 -o gpurun_out/prof_outproj_r1 python bench.py --steps 1 --warmup 1 --ddpm-steps 3 > gpurun_out/ncu3.log 2>&1
timeout 900 python tools/dev_chain.py 1000 64 fp32,tc3f16,tc1f16 > gpurun_out/chain.log 2>&1
cat gpurun_out/rc.txt gpurun_out/cpu.txt; tail -n 4 gpurun_out/test_tc.log; cat gpurun_out/time_tc.log gpurun_out/bench_n1.json gpurun_out/chain.log; tail -n 3 gpurun_out/bench_n1.err gpurun_out/ncu2.log
