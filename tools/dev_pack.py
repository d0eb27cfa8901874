# dev check of the packed ragged batch: parity per item vs oracle, pack vs no-pack, timing of the cfg3 sampler
import os, sys, time, torch
sys.path.insert(0, "/root/repo")
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
g = torch.Generator().manual_seed(5)
steps, lens = 6, [150, 97, 33, 0, 260]
B, T = len(lens), max(lens)
cond = torch.randn(B, 256, T, generator=g) * 0.5; x0 = torch.randn(B, 1, 128, T, generator=g); noise = torch.randn(steps, B, 1, 128, T, generator=g)
sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
res = {}
for pack in ("1", "0"):
    os.environ["DSVC_PACK"] = pack
    gd = model(steps)
    xf = gd.sample(x0.to(DEV), cond.to(DEV), steps, None, noise.to(DEV), lengths=lens).cpu()
    res[pack] = xf
    for b, n in enumerate(lens):
        if n == 0: continue
        with torch.no_grad():
            ref = O.sample(sd, sched, cond[b:b+1, :, :n], x0[b:b+1, :, :, :n], steps, noise[:, b:b+1, :, :, :n])
        print("pack", pack, "item", b, "len", n, "err %.3e" % (xf[b:b+1, :, :, :n] - ref).abs().max().item())
    # PLMS + eval on the same handle
    t = torch.full((B,), 37, dtype=torch.long)
    ev = gd.denoise_fn(x0.to(DEV), t.to(DEV), cond.to(DEV)).cpu() if pack == "0" else None
for b, n in enumerate(lens):
    if n: print("pack vs nopack item", b, "%.3e" % (res["1"][b, :, :, :n] - res["0"][b, :, :, :n]).abs().max().item())
# timing: cfg3 sampler (8 slices 689 +- 25 %)
g = torch.Generator().manual_seed(4242)
lens = (689 * (0.75 + 0.5 * torch.rand(8, generator=g))).round().long().tolist()
T = max(lens); B = 8
cond = (torch.randn(B, 256, T, generator=g) * 0.5).cuda(); x0 = torch.randn(B, 1, 128, T, generator=g).cuda()
for pack in ("1", "0"):
    os.environ["DSVC_PACK"] = pack
    gd = model(1000)
    for i in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gd.sample(x0, cond, 1000, None, None, lengths=lens, seed=3)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("cfg3 sampler pack=%s lens=%s: %.1f ms per 1000 steps (%.1f us/step)" % (pack, lens, dt * 1e3, dt * 1e3))
