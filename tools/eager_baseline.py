"""Second baseline: the reference's modules (oracle port: the same torch.nn.functional calls)
run through STOCK PyTorch eager on the B200 (cuDNN/cuBLAS), same synthetic workload as bench.py."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import diffsvc_oracle as O

T, steps = 862, 30
dev = "cuda"
sd = {k: v.to(dev) for k, v in O.synth_diffnet_weights().items()}
nsd = {k: v.to(dev) for k, v in O.synth_nsf_weights(O.NSF_H_44K).items()}
sched = {k: v.to(dev) for k, v in O.make_schedule(O.linear_beta_schedule(1000, 0.02)).items()}
g = torch.Generator().manual_seed(7)
cond = (torch.randn(1, 256, T, generator=g) * 0.5).to(dev)
x = torch.randn(1, 1, 128, T, generator=g).to(dev)
noise = torch.randn(steps, 1, 1, 128, T, generator=g).to(dev)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def sample(n):
    xx = x
    for i, t in enumerate(reversed(range(0, n))):
        xx = O.p_sample(sd, sched, xx, torch.full((1,), t, device=dev, dtype=torch.long), cond, noise[i])
    return xx


with torch.no_grad():
    sample(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); xf = sample(steps); torch.cuda.synchronize(); td = (time.perf_counter() - t0) / steps
    mel = O.mel_from_x(xf, torch.tensor([[[-5.0]]], device=dev), torch.tensor([[[0.0]]], device=dev)).clamp(-6, 1.5)
    f0 = O.synth_f0(1, T).to(dev)
    ri = torch.rand(1, 9, device=dev); sn = torch.randn(1, T * 512, 9, device=dev)
    O.spec2wav(nsd, O.NSF_H_44K, mel, f0, ri, sn); torch.cuda.synchronize()
    t0 = time.perf_counter(); O.spec2wav(nsd, O.NSF_H_44K, mel, f0, ri, sn); torch.cuda.synchronize(); tv = time.perf_counter() - t0
audio = T * 512 / 44100
print(json.dumps({"torch_eager_b200": {"ms_per_ddpm_step": td * 1e3, "vocoder_ms": tv * 1e3,
                                       "audio_sec_per_s_1000_steps": audio / (td * 1000 + tv), "fp32": True, "tf32": False}}))
