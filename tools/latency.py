"""BASELINE.json configs[2] and configs[4] on the native path (not the headline bench):
  cfg3: PNDM 25-step (pndm_speedup=40) @44.1 kHz, 10 s clip, NSF-HiFiGAN  -> audio-sec/s
  cfg5: flask_api real-time path: 0.5 s chunks (43 frames), 50-step PNDM  -> p50 end-to-end latency
Both through the public classes with host inputs (H2D + D2H inside the timed region)."""
import json, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from diffsvc_b200.hparams import hparams

gd, voc, sd, nsd = B.build_models("tc3f16", 1000)


def run(T, speedup, n_calls, warm):
    hparams["pndm_speedup"] = speedup
    hub, m2p, f0, f0hz = B.synth_inputs(1, T, seed=3)
    hub, m2p, f0, f0hz = (t.pin_memory() for t in (hub, m2p, f0, f0hz))
    out = torch.empty(T * B.HOP).pin_memory()
    ts = []
    for i in range(warm + n_calls):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ret = gd(hub.cuda(non_blocking=True), m2p.cuda(non_blocking=True), None, None, f0.cuda(non_blocking=True), None, None, infer=True)
        wav = voc.spec2wav_torch(ret["mel_out"].clamp(-6.0, 1.5), f0=f0hz.cuda(non_blocking=True), seed=i)
        out.copy_(wav); torch.cuda.synchronize()
        if i >= warm:
            ts.append(time.perf_counter() - t0)
    return ts


with torch.no_grad():
    a = run(862, 40, 10, 3)
    b = run(43, 20, 100, 5)
res = {"cfg3_pndm25_10s": {"ms_p50": statistics.median(a) * 1e3, "audio_sec_per_s": 862 * 512 / 44100 / statistics.median(a)},
       "cfg5_flask_0.5s_pndm50": {"ms_p50": statistics.median(b) * 1e3, "ms_p95": sorted(b)[int(0.95 * len(b))] * 1e3,
                                  "rtf": statistics.median(b) / (43 * 512 / 44100)}}
print(json.dumps(res))
