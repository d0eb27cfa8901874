"""Developer check of the step kernel (csrc/tc_step.cuh): DSVC_STEP=1 vs the per-layer kernels (DSVC_STEP=0) -- results
(bit-identical at the same tile width), oracle parity on a short chain, us per DDPM step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
import synthetic as S
from oracle import diffsvc_oracle as O
hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
DEV = "cuda"
sd = S.synth_diffnet_weights()


def model(steps):
    dn = D.DiffNet(128); dn.load_state_dict(sd, strict=True)
    return D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=steps, loss_type="l2", spec_min=[-5.0], spec_max=[0.0]).cuda().eval()


def inputs(B, T, steps, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 256, T, generator=g) * 0.5, torch.randn(B, 1, 128, T, generator=g), torch.randn(steps, B, 1, 128, T, generator=g))


sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
for (B, T, steps, lens) in ((1, 862, 6, None), (1, 43, 6, None), (3, 150, 6, [150, 97, 33]), (8, 700, 4, [700, 650, 512, 1, 0, 300, 699, 257])):
    cond, x0, noise = inputs(B, T, steps, 11)
    res = {}
    for step in ("1", "0"):
        os.environ["DSVC_STEP"] = step
        gd = model(steps)
        res[step] = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
        if step == "1":
            again = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
            print("B=%d T=%d step kernel deterministic: %s" % (B, T, torch.equal(again, res[step])), flush=True)
            pl = gd.sample(x0.to(DEV), cond.to(DEV), 100, 20, None, lengths=lens).cpu()
        else:
            pl0 = gd.sample(x0.to(DEV), cond.to(DEV), 100, 20, None, lengths=lens).cpu()
            print("   plms step vs per-layer: %.3e" % (pl - pl0).abs().max().item(), flush=True)
    print("B=%d T=%d: step vs per-layer kernels max-abs %.3e (equal: %s)" % (B, T, (res["1"] - res["0"]).abs().max().item(), torch.equal(res["1"], res["0"])), flush=True)
    if T <= 300 or B == 8:
        for b in range(B):
            n = lens[b] if lens else T
            if n == 0 or (B == 8 and b not in (0, 3, 7)):
                continue
            with torch.no_grad():
                ref = O.sample(sd, sched, cond[b:b + 1, :, :n], x0[b:b + 1, :, :, :n], steps, noise[:, b:b + 1, :, :, :n])
            print("   item %d vs oracle: %.3e" % (b, (res["1"][b:b + 1, :, :, :n] - ref).abs().max().item()), flush=True)

g = torch.Generator().manual_seed(4242)
lens8 = (689 * (0.75 + 0.5 * torch.rand(8, generator=g))).round().long().tolist()
for (B, T, lens) in ((1, 862, None), (4, 689, None), (8, max(lens8), lens8)):
    cond = (torch.randn(B, 256, T) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T).cuda()
    for step in ("1", "0"):
        os.environ["DSVC_STEP"] = "0" if step == "0" else "1"
        os.environ.pop("DSVC_STEP_BN", None)
        if step == "64":
            if B == 1:
                continue
            os.environ["DSVC_STEP_BN"] = "64"
        gd = model(1000)
        gd.sample(x0, cond, 20, None, None, lengths=lens, seed=1)
        best = 1e9
        for i in range(3):
            torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gd.sample(x0, cond, 200, None, None, lengths=lens, seed=1); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 200 * 1000)
        print("[time] B=%d T=%d DSVC_STEP=%s: %.1f us per DDPM step" % (B, T, step, best), flush=True)
