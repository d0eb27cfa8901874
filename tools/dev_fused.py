"""Developer A/B probe (not the bench): DDPM step time and per-layer kernel time, default path (two kernels per
WaveNet layer) vs the fused layer kernel (DSVC_FUSED_LAYER, csrc/tc_layer.cuh), with and without the next-layer
weight prefetch.  One process: the switches are read per handle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200 import _lib
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
lib = _lib.load()
sd = O.synth_diffnet_weights()
ev = lambda: torch.cuda.Event(enable_timing=True)


def model(env):
    for k in ("DSVC_FUSED_LAYER", "DSVC_FUSED_PREFETCH", "DSVC_FUSED_FENCE", "DSVC_SPLITK"):
        os.environ.pop(k, None)
    os.environ.update(env)
    dn = D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(sd)
    return D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()


def probe(tag, env, B, T, steps):
    gd = model(env)
    g = torch.Generator().manual_seed(1)
    cond = (torch.randn(B, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T, generator=g).cuda()
    out0 = gd.sample(x0, cond, 3, None, None, seed=1); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = ev(), ev(); a.record(); gd.sample(x0, cond, steps, None, None, seed=1); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / steps * 1000)
    h = gd.denoise_fn.handle()
    parts = []
    for part in (0, 1, 2):
        if part == 2 and "DSVC_FUSED_LAYER" not in env:
            continue
        it = 100
        _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, 10, _lib.current_stream()))
        a, b = ev(), ev(); a.record(); _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, it, _lib.current_stream())); b.record(); torch.cuda.synchronize()
        parts.append("part%d %.2f us" % (part, a.elapsed_time(b) / it * 1000))
    print("[%s] B=%d T=%d: %.1f us/DDPM step (best of 3 x %d steps) | layer 3 back-to-back: %s"
          % (tag, B, T, best, steps, ", ".join(parts)), flush=True)
    return out0.cpu()


if __name__ == "__main__":
    print("== lib %s" % os.environ.get("DSVC_LIB", "product"), flush=True)
    for (B, T, steps) in ((1, 862, 60), (1, 43, 100)):
        ref = probe("default (no split-K)", {"DSVC_SPLITK": "0"}, B, T, steps)
        if T == 43:
            probe("default (split-K)   ", {}, B, T, steps)
        for tag, env in (("fused               ", {"DSVC_FUSED_LAYER": "2"}),
                         ("fused, device fence ", {"DSVC_FUSED_LAYER": "2", "DSVC_FUSED_FENCE": "1"})):
            try:
                out = probe(tag, env, B, T, steps)
                print("    bit-identical to default: %s" % bool(torch.equal(out, ref)), flush=True)
            except Exception as e:
                print("[%s] B=%d T=%d FAILED: %r" % (tag, B, T, e), flush=True)
                raise SystemExit(1)    # a trapped kernel poisons the context: stop here
