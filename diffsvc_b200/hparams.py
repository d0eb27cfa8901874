"""The global `hparams` dict (reference: utils/hparams.py:6).

When the reference tree is importable (drop-in use from inside prophesier/diff-svc) this IS the
reference's dict object, so `infer_tools.infer_tool` writing `hparams['pndm_speedup']`
(infer_tools/infer_tool.py:276) is seen here.  Stand-alone (tests, bench.py on the GPU box) it is a
private dict pre-filled with the 44.1 kHz defaults of training/config_nsf.yaml.
"""

DEFAULTS_44K = dict(
    audio_num_mel_bins=128, audio_sample_rate=44100, hop_size=512, fft_size=2048, win_size=2048, fmin=40, fmax=16000,
    hidden_size=256, residual_layers=20, residual_channels=384, dilation_cycle_length=4, keep_bins=128,
    timesteps=1000, K_step=1000, max_beta=0.02, schedule_type="linear", diff_loss_type="l2",
    spec_min=[-5.0], spec_max=[0.0], pndm_speedup=10, use_nsf=True, no_fs2=True, use_pitch_embed=True,
    use_energy_embed=False, use_spk_embed=False, use_spk_id=False, pitch_norm="log", use_uv=False,
    f0_bin=256, f0_max=1100.0, f0_min=40.0, mel_vmin=-6.0, mel_vmax=1.5,
    vocoder="diffsvc_b200.vocoders.nsf_hifigan.NsfHifiGAN", vocoder_ckpt="checkpoints/nsf_hifigan/model",
)

try:  # drop-in: share the reference's dict
    from utils.hparams import hparams, set_hparams  # type: ignore  # noqa: F401
    SHARED_WITH_REFERENCE = True
except Exception:  # stand-alone
    hparams = dict(DEFAULTS_44K)
    SHARED_WITH_REFERENCE = False

    def set_hparams(config="", exp_name="", hparams_str="", print_hparams=True, global_hparams=True, reset=True, infer=True):
        """Minimal stand-alone loader: one yaml file (no base_config chain), same override syntax."""
        import yaml
        hp = dict(DEFAULTS_44K)
        if config:
            with open(config, encoding="utf-8") as f:
                hp.update(yaml.safe_load(f) or {})
        for kv in filter(None, hparams_str.split(",")):
            k, v = kv.split("=")
            hp[k] = type(hp[k])(v) if k in hp and not isinstance(hp[k], bool) else eval(v)
        hp["infer"] = infer
        if global_hparams:
            hparams.clear()
            hparams.update(hp)
        return hp
