mkdir -p gpurun_out; rm -f gpurun_out/*.log gpurun_out/*.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_voc.csv python - <<'PY' > gpurun_out/ncu_voc.log 2>&1
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
torch.set_num_threads(16)
import diffsvc_oracle as O, diffsvc_b200 as D
D.hparams.update(use_nsf=True)
sd = O.synth_nsf_weights(O.NSF_H_44K)
voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device="cuda")
mel = (torch.randn(1, 862, 128) * 0.8 - 2.0).cuda(); f0 = O.synth_f0(1, 862).cuda()
for _ in range(2):
    w = voc.spec2wav_torch(mel, f0=f0, seed=1)
torch.cuda.synchronize()
PY
python tools/ncu_summary.py launches gpurun_out/launches_voc.csv gpurun_out/voc_summary.md
cat gpurun_out/voc_summary.md; tail -n 3 gpurun_out/ncu_voc.log
