#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on synthetic inputs.

  metric  : audio-sec/s (= 1/RTF) for 1000-step DDPM @ 44.1 kHz, NSF-HiFiGAN vocoder
  workload: BASELINE.json configs[1]: one 10 s clip (862 mel frames, 128 bins), per GPU
            (`--batch B` / `--frames T` widen it; at N > 1 every rank runs its own clip(s) -- weak
            scaling -- and the final waveforms are all-gathered with NCCL inside the timed step)
  a "step": one full pass of the hot path over the batch: 1000 DDPM denoising steps through the
            20-layer WaveNet, denormalise + clip, NSF-HiFiGAN mel -> waveform.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path: the oracle port,
                                                            rank 0 only, bounded sample per step)
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR, HOP, MEL, HID = 44100, 512, 128, 256
FLOP_EVAL_PER_FRAME = 47_677_440         # DiffNet eval, conditioner projections hoisted
FLOP_COND_PER_FRAME = 7_864_320          # one-off conditioner projections of all 20 layers
FLOP_CONV_PER_FRAME = 2 * 3 * 384 * 768  # dilated conv of one layer (the dominant kernel)
FLOP_VOC_PER_FRAME = 648_527_872


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="clips per GPU")
    ap.add_argument("--frames", type=int, default=862, help="mel frames per clip (862 = 10 s @ 44.1 kHz)")
    ap.add_argument("--ddpm-steps", type=int, default=1000)
    ap.add_argument("--math", default="tc3f16", choices=["tc3f16", "fp32", "tc1f16"])
    ap.add_argument("--ref-ddpm-sample", type=int, default=6, help="reference arm: DDPM steps timed per bench step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


def synth_inputs(B, T, seed):
    """Synthetic utterances: hubert-like units at 50 fps gathered to T mel frames, log2-f0, all frames valid."""
    from oracle import diffsvc_oracle as O
    g = torch.Generator().manual_seed(seed)
    Th = max(2, int(T * 50 * HOP / SR))
    hubert = torch.randn(B, Th, HID, generator=g) * 0.5
    mel2ph = (torch.arange(T, dtype=torch.float32) * (Th / T)).long().clamp(max=Th - 1)[None].repeat(B, 1) + 1
    f0_hz = O.synth_f0(B, T, seed=seed + 1)
    f0 = torch.where(f0_hz > 0, torch.log2(f0_hz.clamp(min=1.0)), torch.zeros_like(f0_hz))   # norm_interp_f0 'log'
    return hubert, mel2ph, f0, f0_hz


def build_models(math_mode, ddpm_steps):
    import diffsvc_b200 as D
    from diffsvc_b200.hparams import hparams, DEFAULTS_44K
    from oracle import diffsvc_oracle as O
    hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
    sd = O.synth_diffnet_weights()
    dn = D.DiffNet(MEL, math_mode=math_mode)
    dn.load_state_dict(sd, strict=True)
    gd = D.GaussianDiffusion(None, MEL, dn, timesteps=1000, K_step=ddpm_steps, loss_type="l2", spec_min=[-5.0], spec_max=[0.0])
    gd = gd.cuda().eval()
    nsd = O.synth_nsf_weights(O.NSF_H_44K)
    voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), nsd, device="cuda")
    return gd, voc, sd, nsd


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_threads_autotune(sd, T):
    """Pick the torch thread count that runs one DiffNet eval fastest on this host (many-core boxes
    oversubscribe badly on these small convs); the reference gets its best configuration."""
    from oracle import diffsvc_oracle as O
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 1, MEL, T, generator=g); c = torch.randn(1, HID, T, generator=g)
    best = None
    ncpu = os.cpu_count() or 1
    for n in sorted({min(ncpu, k) for k in (8, 16, 32, 64, ncpu)}):
        torch.set_num_threads(n)
        O.diffnet_forward(sd, x, torch.tensor([5]), c)
        t0 = time.perf_counter()
        O.diffnet_forward(sd, x, torch.tensor([5]), c)
        dt = time.perf_counter() - t0
        if best is None or dt < best[1]:
            best = (n, dt)
    torch.set_num_threads(best[0])
    return best[0]


def cpu_sample(sd, nsd, T, n_ddpm, seed):
    """A bounded sample of the workload on the host: n_ddpm DDPM steps + one vocoder pass of ONE clip.
    Returns (seconds for the DDPM sample, seconds for the vocoder pass)."""
    from oracle import diffsvc_oracle as O
    g = torch.Generator().manual_seed(seed)
    cond = torch.randn(1, HID, T, generator=g) * 0.5
    x = torch.randn(1, 1, MEL, T, generator=g)
    noise = torch.randn(n_ddpm, 1, 1, MEL, T, generator=g)
    sched = O.make_schedule(O.linear_beta_schedule(1000, 0.02))
    with torch.no_grad():
        t0 = time.perf_counter()
        x = O.sample(sd, sched, cond, x, n_ddpm, noise)
        t1 = time.perf_counter()
        mel = O.mel_from_x(x, torch.tensor([[[-5.0]]]), torch.tensor([[[0.0]]])).clamp(-6.0, 1.5)
        f0 = O.synth_f0(1, T)
        O.spec2wav(nsd, O.NSF_H_44K, mel, f0, torch.rand(1, 9, generator=g), torch.randn(1, T * HOP, 9, generator=g))
        t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def run_reference(args, rank):
    """`--impl reference`: the reference's CPU path (oracle port of its torch modules, all host threads it
    can use), rank 0 only.  Each bench step times a bounded sample and extrapolates linearly to the
    full 1000-step clip; the JSON says so in cpu_baseline.sample."""
    if rank != 0:
        return
    from oracle import diffsvc_oracle as O
    sd, nsd = O.synth_diffnet_weights(), O.synth_nsf_weights(O.NSF_H_44K)
    T, n = args.frames, args.ref_ddpm_sample
    threads = cpu_threads_autotune(sd, T)
    times = []
    for i in range(args.warmup + args.steps):
        td, tv = cpu_sample(sd, nsd, T, n, seed=100 + i)
        if i >= args.warmup:
            times.append(td / n * args.ddpm_steps + tv)
    audio = T * HOP / SR
    per_clip = sum(times) / len(times)
    val = audio / per_clip
    line = base_line(args, val, per_clip * 1000.0 * args.batch)
    line.update({"impl": "reference", "dtype": "f32", "gpu_launches": 0,
                 "cpu_baseline": {"value": val, "unit": "audio-sec/s", "cores": threads, "kind": "port",
                                  "sample": "%d of %d DDPM steps + 1 NSF-HiFiGAN pass of one %d-frame clip per bench step, "
                                            "DDPM part extrapolated linearly; torch CPU fp32, %d threads (autotuned of %d cpus)"
                                            % (n, args.ddpm_steps, T, threads, os.cpu_count() or 1)},
                 "e2e": {"value": val, "unit": "audio-sec/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


def base_line(args, value, ms_per_step):
    return {"metric": "audio-sec/s (1/RTF), 1000-step DDPM @44.1kHz + NSF-HiFiGAN", "value": value, "unit": "audio-sec/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: %d clip(s)/GPU x %d frames (%.2f s @44.1kHz, 128 mel bins), "
                                   "%d-step DDPM + NSF-HiFiGAN" % (args.batch, args.frames, args.frames * HOP / SR, args.ddpm_steps),
                       "batch_per_gpu": args.batch, "frames": args.frames, "ddpm_steps": args.ddpm_steps,
                       "parallelism": "independent clips per GPU, NCCL all-gather of waveforms" if args.gpus > 1 else "single GPU",
                       "l2": "256 MiB L2 flush between timed steps"}}


# ------------------------------------------------------------------------------------------ our arm
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    assert torch.cuda.is_available(), "bench.py needs a B200 (there is no CPU fallback on the product path)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)
    from diffsvc_b200 import _lib
    lib = _lib.load()
    B, T, NS = args.batch, args.frames, args.ddpm_steps
    gd, voc, sd, nsd = build_models(args.math, NS)
    hubert, mel2ph, f0, f0_hz = synth_inputs(B, T, seed=1000 + rank)
    pin = lambda t: t.pin_memory()
    h_hubert, h_mel2ph, h_f0, h_f0hz = pin(hubert), pin(mel2ph), pin(f0), pin(f0_hz)
    h_wav = torch.empty(B * T * HOP, dtype=torch.float32).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    gathered = torch.empty(world * B * T * HOP, dtype=torch.float32, device="cuda") if world > 1 else None

    def vocode(mel, f0d, seed):
        mel = mel.clamp(-6.0, 1.5)                                   # after_infer clip (infer_tool.py:183)
        wav = voc.spec2wav_torch(mel, f0=f0d, seed=seed)
        if world > 1:
            dist.all_gather_into_tensor(gathered, wav)               # C0: the only cross-GPU exchange
        return wav

    # device-resident arm: conditioning + initial noise already in HBM
    with torch.no_grad():
        ret0 = gd.fs2(hubert.cuda(), mel2ph.cuda(), None, None, f0.cuda().clone(), None, None, skip_decoder=True, infer=True)
    cond_d = ret0["decoder_inp"].transpose(1, 2).contiguous()
    x0_d = torch.randn(B, 1, MEL, T, device="cuda")
    f0hz_d = f0_hz.cuda()

    def step_device(i):
        x = gd.sample(x0_d, cond_d, NS, None, None, seed=17 + i)
        mel = gd.denorm_spec(x[:, 0].transpose(1, 2))
        return vocode(mel, f0hz_d, seed=i)

    def step_e2e(i):
        hub = h_hubert.cuda(non_blocking=True); m2p = h_mel2ph.cuda(non_blocking=True)
        f0d = h_f0.cuda(non_blocking=True); f0hz = h_f0hz.cuda(non_blocking=True)
        ret = gd(hub, m2p, None, None, f0d, None, None, infer=True, seed=17 + i)     # the public call (diffusion.py:227)
        wav = vocode(ret["mel_out"], f0hz, seed=i)
        h_wav.copy_(wav, non_blocking=True)
        return wav

    def timed(fn, n_warm, n_steps, sampler=None):
        tot = 0.0
        for i in range(n_warm + n_steps):
            if dist is not None:
                dist.barrier()
            flush.fill_(float(i)); torch.cuda.synchronize()
            if i == n_warm and sampler is not None:
                sampler.start()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(i); b.record()
            torch.cuda.synchronize()
            if i >= n_warm:
                tot += a.elapsed_time(b)
        if dist is not None:
            t = torch.tensor([tot], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); tot = float(t.item())
        return tot / n_steps     # ms per step, max over ranks

    with torch.no_grad():
        clk = ClockSampler(local)
        l0 = lib.dsvc_launch_count()
        ms_dev = timed(step_device, args.warmup, args.steps, clk)
        clocks = clk.stop()
        launches = (lib.dsvc_launch_count() - l0) // (args.warmup + args.steps) * args.steps
        ms_e2e = timed(step_e2e, args.warmup, args.steps)
        # dominant kernel alone: dilated conv + gate of one WaveNet layer (CUDA events on the launch stream)
        h = gd.denoise_fn.handle()
        it = 200
        _lib.check(lib.dsvc_diffnet_run_layer(h, 5, 0, 20, _lib.current_stream()))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(lib.dsvc_diffnet_run_layer(h, 5, 0, it, _lib.current_stream())); b.record(); torch.cuda.synchronize()
        conv_us = a.elapsed_time(b) / it * 1000.0

    audio = world * B * T * HOP / SR
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    peak_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
    conv_tf = FLOP_CONV_PER_FRAME * B * T / conv_us / 1e6
    sampler_flop = (FLOP_EVAL_PER_FRAME * NS + FLOP_COND_PER_FRAME) * B * T
    line = base_line(args, audio / (ms_dev / 1000.0), ms_dev)
    line.update({
        "dtype": "f32 (fp16 hi/lo split x3 on tcgen05, fp32 accumulate)" if args.math == "tc3f16" else args.math,
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": audio / (ms_e2e / 1000.0), "unit": "audio-sec/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in (h_hubert, h_mel2ph, h_f0, h_f0hz))),
                "d2h_bytes_per_step": int(h_wav.numel() * 4), "api": "GaussianDiffusion.forward + NsfHifiGAN.spec2wav_torch"},
        "roofline": {"bound": "tensor", "kernel": "tc_gemm_kernel<EpiGate> (dilated conv + conditioner + gate, one layer)",
                     "achieved": conv_tf, "peak": peak_burst, "unit": "TFLOP/s", "frac": conv_tf / peak_burst,
                     "traffic": 7.54e6 if (B == 1 and T == 862) else None, "us_per_launch": conv_us,
                     "note": "algorithmic fp32-equivalent FLOPs (2*3*C*2C per frame) / CUDA-event time of the kernel run back to back; "
                             "the error-compensated fp16 hi/lo split executes 3x those FLOPs on the tensor pipe (ceiling of frac = 0.33). "
                             "traffic = dram bytes per launch from profiles/r1c_conv_kernel_full.md (ncu --set full; ~= the "
                             "algorithmic 7.5 MB: weights 3.5 + conditioner 2.6 + activations 1.3). "
                             "peak = MEASURED_PEAKS.json bf16_tflops (burst)" + ("" if peaks else " [fallback]")},
        "sampler_flops": {"achieved_tflops": sampler_flop / (ms_dev / 1000.0) / 1e12, "peak_sustained": peak_sus,
                          "note": "whole step incl. vocoder time; algorithmic DiffNet FLOPs only"},
    })
    if rank == 0 and world == 1:      # the CPU baseline is reported at N = 1 only (the other ranks would idle at the barrier)
        threads = cpu_threads_autotune(sd, T)
        td, tv = cpu_sample(sd, nsd, T, args.ref_ddpm_sample, seed=5)
        per_clip = td / args.ref_ddpm_sample * NS + tv
        line["cpu_baseline"] = {"value": (T * HOP / SR) / per_clip, "unit": "audio-sec/s", "cores": threads, "kind": "port",
                                "sample": "%d of %d DDPM steps (%.2f s) + 1 NSF-HiFiGAN pass (%.2f s) of one %d-frame clip, DDPM part "
                                          "extrapolated linearly; oracle port of the reference modules, torch CPU fp32, %d threads"
                                          % (args.ref_ddpm_sample, NS, td, tv, T, threads)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
