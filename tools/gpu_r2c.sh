#!/bin/bash
# Round-2c dev run: packed ragged batches (parity + timing), out-projection tile-width experiment, targeted tests.
mkdir -p gpurun_out
timeout 600 python tools/dev_pack.py > gpurun_out/r2c_pack.txt 2>&1; echo "pack rc=$?"
grep -v Warning gpurun_out/r2c_pack.txt | tail -n 30
for v in default 128; do
  if [ $v = default ]; then unset DSVC_OUT_BN; else export DSVC_OUT_BN=$v; fi
  echo "== DSVC_OUT_BN=$v" >> gpurun_out/r2c_outbn.txt
  timeout 300 python - >> gpurun_out/r2c_outbn.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, "tools"); sys.argv = ["x", "tc3f16"]
import dev_time as d
d.time_ddpm("tc3f16", 1, 862, 60); d.time_ddpm("tc3f16", 1, 862, 60); d.time_ddpm("tc3f16", 1, 43, 60)
PY
done
unset DSVC_OUT_BN
grep -v Warning gpurun_out/r2c_outbn.txt
( timeout 900 python -m pytest -q -m gpu tests/test_tc_pair.py tests/test_gpu_parity.py -k "ragged or batch or empty or golden" ) > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
tail -n 8 gpurun_out/r2c_tests.log
