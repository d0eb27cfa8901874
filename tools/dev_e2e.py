"""Developer probe: where does the end-to-end (public API, host inputs) step spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
gd, voc, sd, nsd = B.build_models("tc3f16", 1000)
hub, m2p, f0, f0hz = B.synth_inputs(1, 862, seed=1000)
hub, m2p, f0, f0hz = (t.pin_memory() for t in (hub, m2p, f0, f0hz))
out = torch.empty(862 * 512).pin_memory()

def seg(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"  {name:28s} {1e3*(time.perf_counter()-t0):8.2f} ms", flush=True); return r

with torch.no_grad():
    for it in range(3):
        print("iter", it)
        d = seg("h2d", lambda: [t.cuda(non_blocking=True) for t in (hub, m2p, f0, f0hz)])
        ret = seg("fs2 (cond encoder)", lambda: gd.fs2(d[0], d[1], None, None, d[2], None, None, skip_decoder=True, infer=True))
        cond = ret["decoder_inp"].transpose(1, 2)
        x = seg("randn", lambda: torch.randn(1, 1, 128, 862, device="cuda"))
        seg("prepare", lambda: gd.denoise_fn.prepare(cond, None))
        xs = seg("sample (1000 DDPM steps)", lambda: gd.sample(x, cond, 1000, None, None, None, seed=17 + it))
        mel = gd.denorm_spec(xs[:, 0].transpose(1, 2)).clamp(-6, 1.5)
        wav = seg("vocoder", lambda: voc.spec2wav_torch(mel, f0=d[3], seed=it))
        seg("d2h", lambda: out.copy_(wav))
        seg("full forward()", lambda: gd(d[0], d[1], None, None, d[2].clone(), None, None, infer=True, seed=99 + it))
