"""Developer A/B probe (not the bench): DDPM step time and per-layer kernel times, CTA-pair main loop (tc_pair.cuh,
default) vs the single-CTA kernels (DSVC_TC_PAIR=0).  One process: the switch is read per launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200 import _lib
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
import synthetic as S

hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
lib = _lib.load()
sd = S.synth_diffnet_weights()
ev = lambda: torch.cuda.Event(enable_timing=True)


def probe(tag, env, B, T, steps):
    for k in ("DSVC_TC_PAIR", "DSVC_SPLITK", "DSVC_TC_BN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    dn = D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(sd)
    gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=1000, spec_min=[-5.0], spec_max=[0.0]).cuda().eval()
    g = torch.Generator().manual_seed(1)
    cond = (torch.randn(B, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T, generator=g).cuda()
    out0 = gd.sample(x0, cond, 3, None, None, seed=1); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = ev(), ev(); a.record(); gd.sample(x0, cond, steps, None, None, seed=1); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / steps * 1000)
    h = gd.denoise_fn.handle()
    parts = []
    for part in (0, 1):
        it = 100
        _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, 10, _lib.current_stream()))
        a, b = ev(), ev(); a.record(); _lib.check(lib.dsvc_diffnet_run_layer(h, 3, part, it, _lib.current_stream())); b.record(); torch.cuda.synchronize()
        parts.append("part%d %.2f us" % (part, a.elapsed_time(b) / it * 1000))
    print("[%s] B=%d T=%d: %.1f us/DDPM step (best of 3 x %d steps) | layer 3 back-to-back: %s"
          % (tag, B, T, best, steps, ", ".join(parts)), flush=True)
    return out0.cpu()


if __name__ == "__main__":
    print("== lib %s" % os.environ.get("DSVC_LIB", "product"), flush=True)
    for (B, T, steps) in ((1, 862, 60), (8, 689, 30), (1, 43, 100)):
        ref = probe("single-CTA          ", {"DSVC_TC_PAIR": "0", "DSVC_SPLITK": "0"}, B, T, steps)
        out = probe("CTA pairs           ", {"DSVC_SPLITK": "0"}, B, T, steps)
        print("    max |pair - single| / max|single| = %.2e" % ((out - ref).abs().max().item() / ref.abs().max().item()), flush=True)
        if B == 8:
            probe("CTA pairs, BN = 128 ", {"DSVC_SPLITK": "0", "DSVC_TC_BN": "128"}, B, T, steps)
        if T == 43:
            probe("single-CTA + split-K", {"DSVC_TC_PAIR": "0"}, B, T, steps)
            probe("CTA pairs + split-K ", {}, B, T, steps)
