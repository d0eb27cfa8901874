"""Vocoder timing: NSF-HiFiGAN (44.1 kHz topology) on one 10 s clip / 8 clips, tcgen05 ResBlocks vs all-FFMA."""
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
for B, T in ((1, 862), (8, 689), (1, 43)):
    mel = (torch.randn(B, T, 128) * 0.8 - 2.0).cuda()
    f0 = O.synth_f0(B, T).cuda()
    res = {}
    for mode in ("tc", "tc-wide-only", "fp32"):     # tc: every ResBlock stage on tcgen05 (narrow stages through tap packing)
        os.environ.pop("DSVC_NSF_MATH", None); os.environ.pop("DSVC_NSF_NARROW", None)
        if mode == "fp32":
            os.environ["DSVC_NSF_MATH"] = "fp32"
        elif mode == "tc-wide-only":
            os.environ["DSVC_NSF_NARROW"] = "0"
        voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), sd, device="cuda")
        for _ in range(3):
            w = voc.spec2wav_torch(mel, f0=f0, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            w = voc.spec2wav_torch(mel, f0=f0, seed=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        res[mode] = w
        flop = 648527872.0 * B * T
        print("[voc] %s B=%d T=%d: %.3f ms  (%.1f TFLOP/s algorithmic, %.0fx real time)" %
              (mode, B, T, ms, flop / ms / 1e9, B * T * 512 / 44100 / (ms / 1e3)))
    print("      max |tc - fp32| = %.2e   max |tc-wide-only - fp32| = %.2e   rms |tc - fp32| = %.2e" % (
        (res["tc"] - res["fp32"]).abs().max().item(), (res["tc-wide-only"] - res["fp32"]).abs().max().item(),
        (res["tc"] - res["fp32"]).pow(2).mean().sqrt().item()))
