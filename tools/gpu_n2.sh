#!/bin/bash
# bench.py at N = 2 as the driver launches it (one rank per GPU over NCCL)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
wc -l gpurun_out/bench_n2.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "per-rank", d["per_rank_ms_per_step"])
print("cfg3", {k: d["cfg3_sliced_batch"][k] for k in ("n_slices", "audio_sec_per_s", "job_ms", "per_rank_busy_ms", "per_rank_us_per_ddpm_step", "gather_ms_max", "imbalance")})
PY
