"""SURVEY.md section 8 row a17 / boundary row b, on hardware: the reference's OWN `Svc.infer` + `Svc.after_infer`
(infer_tools/infer_tool.py:104-201, the call of infer.py:59, batch.py:11 and flask_api.py:31) run UNCHANGED over the
native classes via `diffsvc_b200.dropin.install()`, from synthetic checkpoint FILES in the reference's layouts, and
compared with the unmodified reference run alone on the CPU (tests/svc_e2e.py; baseline/_ref on the GPU box).

Gates (BASELINE.json north_star): denoised mel <= 1e-3 max-abs, waveform <= 1e-4 RMS, f0 arrays identical.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import ref_harness as rh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rh.reference_available(), reason="no reference tree (baseline/_ref is made by __graft_entry__.build())")]


def _py(args, timeout=900):
    r = subprocess.run([sys.executable, os.path.join(HERE, "svc_e2e.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert "SVC_E2E_OK" in r.stdout or args[0] == "make", (r.stdout[-3000:], r.stderr[-5000:])
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-5000:])
    return r


def _arms(ws, acc, extra_native=()):
    nat, ref = os.path.join(ws, "native.npz"), os.path.join(ws, "reference.npz")
    _py(["run", ws, "--arm", "native", "--acc", str(acc), "--out", nat] + list(extra_native))
    _py(["run", ws, "--arm", "reference", "--acc", str(acc), "--out", ref, "--draws", nat])
    return np.load(nat), np.load(ref)


def _compare(a, b, tag):
    assert int(b["replayed"]) == int(b["recorded"]) > 0, (int(b["replayed"]), int(b["recorded"]))   # every draw was consumed in order
    assert int(a["launches"]) > 0                                    # the native library did the work
    assert a["mel_pred"].shape == b["mel_pred"].shape and a["wav"].shape == b["wav"].shape
    mel_err = float(np.abs(a["mel_pred"] - b["mel_pred"]).max())
    rms = float(np.sqrt(np.mean((a["wav"].astype(np.float64) - b["wav"]) ** 2)))
    sig = float(np.sqrt(np.mean(b["wav"].astype(np.float64) ** 2)))
    print("%s: mel max-abs %.3e, wav rms %.3e (signal rms %.3e), %d frames" % (tag, mel_err, rms, sig, a["mel_pred"].shape[0]))
    assert np.array_equal(a["f0_gt"], b["f0_gt"])
    assert np.allclose(a["f0_pred"], b["f0_pred"], rtol=1e-6, atol=1e-4)
    assert np.allclose(a["f0_voc"], b["f0_voc"], rtol=1e-6, atol=1e-4)
    assert sig > 1e-3
    # The gates are stated for mels in the nominal range (after_infer clips to [-6, 1.5]).  With synthetic weights the
    # un-clamped PLMS solve leaves it by orders of magnitude (the DDPM chain clamps x0 every step and does not), and
    # what is compared here is the clipped mel: scale the gate by the unclipped range, as tests/test_gpu_parity.py does.
    scale = max(1.0, float(b["mel_absmax"]) / 6.0)
    print("    unclipped |mel| max %.3e -> gate scale %.2f" % (float(b["mel_absmax"]), scale))
    assert abs(float(a["mel_absmax"]) - float(b["mel_absmax"])) <= 1e-3 * scale
    assert mel_err <= 1e-3 * scale, (mel_err, scale)
    assert rms <= 1e-4 * scale, (rms, scale)
    return mel_err, rms


def test_svc_infer_plms_through_reference_glue(tmp_path):
    """50-iteration PLMS (acc = 20, K_step = 1000), 3 s clip: `Svc.infer` of the unmodified infer_tool over the native
    GaussianDiffusion / DiffNet / NsfHifiGAN, then the same with the device-side `after_infer` bound over it."""
    ws = str(tmp_path / "proj")
    _py(["make", ws, "--seconds", "3", "--k-step", "1000"])
    a, b = _arms(ws, 20)
    _compare(a, b, "Svc.infer PLMS-50 (reference after_infer)")
    nat2 = os.path.join(ws, "native_glue.npz")
    _py(["run", ws, "--arm", "native", "--acc", "20", "--patch-after-infer", "--out", nat2])
    c = np.load(nat2)
    # same draws (same generator seed), device-side mask / clip instead of the numpy round trip: same waveform
    assert np.array_equal(c["mel_pred"].reshape(a["mel_pred"].shape), a["mel_pred"])
    assert float(np.abs(c["wav"] - a["wav"]).max()) <= 1e-6
    _compare(c, b, "Svc.infer PLMS-50 (device-side after_infer)")


def test_svc_infer_ddpm_through_reference_glue(tmp_path):
    """Plain DDPM (acc = 1) with K_step = 100: per-step noise injected / replayed in call order."""
    ws = str(tmp_path / "proj")
    _py(["make", ws, "--seconds", "2", "--k-step", "100"])
    a, b = _arms(ws, 1)
    _compare(a, b, "Svc.infer DDPM-100")
