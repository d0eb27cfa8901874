mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "plms_chain_full" > gpurun_out/test_tc.log 2>&1; echo "plms rc=$?" > gpurun_out/rc.txt
timeout 300 python tools/dev_time.py tc3f16 > gpurun_out/time_tc.log 2>&1
timeout 300 python tools/dev_time.py fp32 > gpurun_out/time_fp32.log 2>&1
timeout 600 python tools/dev_chain.py 1000 96 fp32,tc3f16,tc1f16 --fp64 > gpurun_out/chain.log 2>&1
cat > /tmp/prof.py <<'PY'
import sys; sys.path.insert(0,'.')
import torch, diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O
hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"]=1
sd=O.synth_diffnet_weights(); dn=D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(sd)
gd=D.GaussianDiffusion(None,128,dn,timesteps=1000,K_step=1000,spec_min=[-5.0],spec_max=[0.0]).cuda().eval()
cond=(torch.randn(1,256,862)*0.5).cuda(); x0=torch.randn(1,1,128,862).cuda()
gd.sample(x0,cond,4,None,None,seed=1); torch.cuda.synchronize()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_tc3_b1.csv python /tmp/prof.py > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel.*EpiGate -s 30 -c 2 -o gpurun_out/prof_conv_r1 python /tmp/prof.py > gpurun_out/ncu2.log 2>&1
cat gpurun_out/rc.txt; tail -3 gpurun_out/test_tc.log; cat gpurun_out/time_tc.log gpurun_out/time_fp32.log gpurun_out/chain.log; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
