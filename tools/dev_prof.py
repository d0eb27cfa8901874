"""A short sampler run for ncu captures: python tools/dev_prof.py B [ddpm_steps]  (B = 1: one 862-frame clip; B = 8: the
cfg3 ragged batch, packed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
import synthetic as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
dn = D.DiffNet(128); dn.load_state_dict(S.synth_diffnet_weights(), strict=True)
gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, loss_type="l2", spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
g = torch.Generator().manual_seed(4242)
lens = None
T = 862
if B > 1:
    lens = (689 * (0.75 + 0.5 * torch.rand(B, generator=g))).round().long().tolist(); T = max(lens)
cond = (torch.randn(B, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T, generator=g).cuda()
gd.sample(x0, cond, steps, None, None, lengths=lens, seed=1); torch.cuda.synchronize()
print("done", B, T, steps)
