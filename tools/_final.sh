bash tools/gpu_validate.sh
DSVC_SPLITK=0 timeout 300 python tools/latency.py > gpurun_out/latency_nosplitk.json 2> gpurun_out/latency_nosplitk.err
echo NOSPLITK; cat gpurun_out/latency_nosplitk.json
