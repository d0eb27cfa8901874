"""Two NSF-HiFiGAN forwards of one 10 s clip (profiling target for ncu: launch list / EpiVoc kernel capture)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
torch.set_num_threads(16)
import diffsvc_oracle as O  # noqa: E402
import diffsvc_b200 as D  # noqa: E402

D.hparams.update(use_nsf=True)
sd = O.synth_nsf_weights(O.NSF_H_44K)
voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device="cuda")
mel = (torch.randn(1, 862, 128) * 0.8 - 2.0).cuda()
f0 = O.synth_f0(1, 862).cuda()
for _ in range(2):
    w = voc.spec2wav_torch(mel, f0=f0, seed=1)
torch.cuda.synchronize()
