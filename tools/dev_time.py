"""Developer timing probe (not the bench): per-mode DDPM step time, layer kernels, vocoder."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200 import _lib
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
lib = _lib.load()


def ev():
    return torch.cuda.Event(enable_timing=True)


def time_ddpm(mode, B, T, steps):
    sd = O.synth_diffnet_weights()
    dn = D.DiffNet(128, math_mode=mode); dn.load_state_dict(sd)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
    cond = (torch.randn(B, 256, T) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T).cuda()
    gd.sample(x0, cond, 3, None, None, seed=1); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record(); gd.sample(x0, cond, steps, None, None, seed=1); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    audio = B * T * 512 / 44100.0
    print(f"[ddpm] mode={mode} B={B} T={T}: {ms*1000:.1f} us/step -> 1000 steps {ms:.3f} s -> {audio/ms:.1f} audio-s/s (sampler only)", flush=True)
    h = dn.handle()
    for part in (0, 1):
        it = 50
        _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, 5, _lib.current_stream()))
        a, b = ev(), ev(); a.record(); _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, it, _lib.current_stream())); b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / it * 1000
        fl = (2 * 3 * 384 * 768 if part == 0 else 2 * 384 * 768) * B * T
        print(f"   layer part{part}: {us:.2f} us  {fl/us/1e6:.1f} TFLOP/s algorithmic", flush=True)
    del gd, dn


def time_voc(B, T):
    nsd = O.synth_nsf_weights(O.NSF_H_44K)
    voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), nsd, device="cuda")
    mel = (torch.randn(B, T, 128) * 0.8 - 2).cuda(); f0 = O.synth_f0(B, T).cuda()
    voc.spec2wav_torch(mel, f0=f0, seed=1); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record(); voc.spec2wav_torch(mel, f0=f0, seed=1); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"[voc] B={B} T={T}: {ms:.2f} ms -> {B*T*512/44100/ms*1000:.1f} audio-s/s; {0.6485*B*T/ms:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["fp32", "tc3f16"]
    for m in modes:
        for (B, T, st) in ((1, 862, 30), (8, 689, 10)):
            try:
                time_ddpm(m, B, T, st)
            except Exception as e:
                print("FAILED", m, B, T, repr(e), flush=True)
    time_voc(1, 862); time_voc(8, 689)
