"""One NSF-HiFiGAN forward of a short clip (compute-sanitizer target): python tools/voc_small.py [T]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffsvc_b200 as D
import synthetic as S
T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
D.hparams.update(use_nsf=True)
sd = S.synth_nsf_weights(S.NSF_H_44K)
voc = D.NsfHifiGAN.from_state_dict(dict(S.NSF_H_44K), sd, device="cuda")
mel = (torch.randn(1, T, 128) * 0.8 - 2.0).cuda()
f0 = S.synth_f0(1, T).cuda()
w = voc.spec2wav_torch(mel, f0=f0, seed=1)
torch.cuda.synchronize()
print("ok", w.shape, float(w.abs().max()))
