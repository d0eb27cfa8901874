"""One process, two GPUs: handles live on their tensors' device, whatever the current device is (function attributes such
as the dynamic shared-memory limit are per device, every launch runs under the tensors' device).  Skipped on a 1-GPU box."""
import pytest
import torch

import diffsvc_b200 as D
from diffsvc_b200.hparams import hparams, DEFAULTS_44K
from oracle import diffsvc_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_sampler_and_vocoder_on_the_second_device_while_the_first_is_current():
    hparams.clear(); hparams.update(DEFAULTS_44K); hparams["pndm_speedup"] = 1
    sd, nsd = O.synth_diffnet_weights(), O.synth_nsf_weights(O.NSF_H_44K)
    steps, T = 4, 200
    g = torch.Generator().manual_seed(9)
    cond = torch.randn(2, 256, T, generator=g) * 0.5
    x0 = torch.randn(2, 1, 128, T, generator=g)
    noise = torch.randn(steps, 2, 1, 128, T, generator=g)
    f0 = O.synth_f0(1, T)
    rand_ini = torch.rand(1, 9, generator=g)
    sn = torch.randn(1, T * 512, 9, generator=g)
    out = {}
    torch.cuda.set_device(0)                                  # stays current for both runs
    for dev in ("cuda:0", "cuda:1"):
        dn = D.DiffNet(128, math_mode="tc3f16"); dn.load_state_dict(sd, strict=True)
        gd = D.GaussianDiffusion(None, 128, dn, timesteps=1000, K_step=steps, loss_type="l2", spec_min=[-5.0], spec_max=[0.0]).to(dev).eval()
        x = gd.sample(x0.to(dev), cond.to(dev), steps, None, noise.to(dev), lengths=[T, 150])
        voc = D.NsfHifiGAN.from_state_dict(dict(O.NSF_H_44K), nsd, device=dev)
        mel = gd.denorm_spec(x[:1, 0].transpose(1, 2)).clamp(-6.0, 1.5)
        wav = voc.spec2wav_torch(mel, f0=f0.to(dev), rand_ini=rand_ini, sine_noise=sn)
        assert x.device == torch.device(dev) and wav.device == torch.device(dev)
        out[dev] = (x.cpu(), wav.cpu())
    assert torch.cuda.current_device() == 0
    assert torch.equal(out["cuda:0"][0], out["cuda:1"][0])    # same kernels, same inputs: same bits on either GPU
    assert torch.equal(out["cuda:0"][1], out["cuda:1"][1])
