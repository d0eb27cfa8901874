"""Developer probe: full-length DDPM chain parity (1000 steps) of each math mode against the CPU oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 96
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "tc3f16", "tc1f16"]
sd = O.synth_diffnet_weights()
g = torch.Generator().manual_seed(7)
cond = torch.randn(1, 256, T, generator=g) * 0.5
x0 = torch.randn(1, 1, 128, T, generator=g)
noise = torch.randn(steps, 1, 1, 128, T, generator=g)
sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
t0 = time.time()
torch.set_num_threads(min(16, os.cpu_count()))
ref = O.sample(sd, sched, cond, x0, steps, noise)
print(f"oracle fp32: {time.time()-t0:.1f}s on {os.cpu_count()} threads", flush=True)
if "--fp64" in sys.argv:
    sd64 = {k: v.double() for k, v in sd.items()}
    r64 = O.sample(sd64, sched, cond.double(), x0.double(), steps, noise.double(), dtype=torch.float64)
    print("oracle fp32 vs fp64: max-abs mel %.3e" % ((ref.double() - r64).abs().max().item() * 2.5), flush=True)
for m in modes:
    dn = D.DiffNet(128, math_mode=m); dn.load_state_dict(sd)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
    x = gd.sample(x0.cuda(), cond.cuda(), steps, None, noise.cuda()).cpu()
    d = (x - ref).abs() * 2.5
    print(f"mode {m}: mel max-abs err {d.max().item():.3e}  mean {d.mean().item():.3e}  frac>1e-3 {(d>1e-3).float().mean().item():.4f}", flush=True)
    del gd, dn
