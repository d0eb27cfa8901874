"""Mel analysis (`wav2spec`) -- drop-in for modules/nsf_hifigan/nvSTFT.py (`STFT`, `load_wav_to_torch`).

The tensor work of `STFT.get_mel` (nvSTFT.py:72-104) is ONE kernel launch through `dsvc_mel_analysis`
(include/dsvc.h): reflect pad, framing, Hann window, 2048-point DFT, magnitude, mel filterbank, log-clamp and
the log10 rescale of `NsfHifiGAN.wav2spec` (network/vocoders/nsf_hifigan.py:76-92) fused.  There is no CPU path
for it.  What stays host code is what the reference also does on the host: reading the file
(`load_wav_to_torch`, nvSTFT.py:14-43) and building the constant mel basis.

The mel basis is `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` (nvSTFT.py:87) of librosa 0.9.1 (pinned in
the reference's requirements.txt:39; not vendored, absent here): Slaney mel scale (`htk=False`) with Slaney area
normalisation, restated from its published algorithm in `slaney_mel_basis`; `tests/test_mel_analysis.py` pins it
against torchaudio's independent implementation of the same filterbank.
"""
import numpy as np
import torch

from .. import _lib

# ---- librosa.filters.mel (0.9.1) restated ---------------------------------------------------------------------
_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    mel = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    return np.where(log_t, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-300) / _MIN_LOG_HZ) / _LOGSTEP, mel)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def slaney_mel_basis(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """[n_mels, 1 + n_fft//2] float32 triangular filters, Slaney scale, area-normalised (librosa default)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


def band_ranges(basis):
    """half-open [lo, hi) range of the non-zero columns of every row of a [n_mels, n_bins] basis."""
    nz = basis != 0
    any_nz = nz.any(axis=1)
    lo = np.where(any_nz, nz.argmax(axis=1), 0)
    hi = np.where(any_nz, basis.shape[1] - nz[:, ::-1].argmax(axis=1), 0)
    return lo.astype(np.int32), hi.astype(np.int32)


# ---- host I/O (behaviour of nvSTFT.py:14-43) ----------------------------------------------------------------------
def _read_audio(path):
    """(samples [n, channels], rate) through soundfile when the host has it, else scipy's PCM/float wav reader."""
    try:
        import soundfile
    except ImportError:
        from scipy.io import wavfile
        rate, samples = wavfile.read(path)
        return samples.reshape(len(samples), -1), rate
    return soundfile.read(path, always_2d=True)


def _full_scale(samples):
    """Divisor that maps the file's sample format to [-1, 1]: int16 / int32 full scale, or 1 for float audio that is
    already normalised (floats beyond 1.01 are taken as integer-valued PCM stored as float)."""
    if np.issubdtype(samples.dtype, np.integer):
        return -np.iinfo(samples.dtype).min
    peak = max(np.amax(samples), -np.amin(samples))
    if peak > 2 ** 15:
        return 2 ** 31 + 1
    return 2 ** 15 + 1 if peak > 1.01 else 1.0


def load_wav_to_torch(full_path, target_sr=None, return_empty_on_exception=False):
    """First channel of an audio file as a float tensor in [-1, 1] (+ its rate), resampled to `target_sr`."""
    rate = None
    fallback_rate = lambda: rate or target_sr or 48000   # noqa: E731
    try:
        samples, rate = _read_audio(full_path)
    except Exception as ex:
        print(f"'{full_path}' failed to load.\nException:")
        print(ex)
        if return_empty_on_exception:
            return [], fallback_rate()
        raise Exception(ex)
    if samples.ndim > 1:
        samples = samples[:, 0]
        assert len(samples) > 2, "audio shorter than 3 samples (wrong axis?)"
    audio = torch.FloatTensor(samples.astype(np.float32)) / _full_scale(samples)
    if return_empty_on_exception and not torch.isfinite(audio).all():
        return [], fallback_rate()
    if target_sr is not None and rate != target_sr:
        import librosa                           # resampling stays the reference's host dependency
        audio = torch.from_numpy(librosa.core.resample(audio.numpy(), orig_sr=rate, target_sr=target_sr))
        rate = target_sr
    return audio, rate


# ---- STFT: nvSTFT.py:58-109 --------------------------------------------------------------------------------------
class STFT:
    def __init__(self, sr=22050, n_mels=80, n_fft=1024, win_size=1024, hop_length=256, fmin=20, fmax=11025,
                 clip_val=1e-5):
        self.target_sr = sr
        self.n_mels = n_mels
        self.n_fft = n_fft
        self.win_size = win_size
        self.hop_length = hop_length
        self.fmin = fmin
        self.fmax = fmax
        self.clip_val = clip_val
        self.mel_basis = {}
        self.hann_window = {}
        self._bands = {}

    def _constants(self, device):
        key = str(self.fmax) + "_" + str(device)
        if key not in self.mel_basis:
            basis = slaney_mel_basis(self.target_sr, self.n_fft, self.n_mels, self.fmin, self.fmax)
            lo, hi = band_ranges(basis)
            self.mel_basis[key] = torch.from_numpy(basis).to(device)
            self._bands[key] = (torch.from_numpy(lo).to(device), torch.from_numpy(hi).to(device))
            win = torch.hann_window(self.win_size)
            if self.win_size < self.n_fft:       # torch.stft centre-pads a short window to n_fft
                left = (self.n_fft - self.win_size) // 2
                win = torch.nn.functional.pad(win, (left, self.n_fft - self.win_size - left))
            assert win.numel() == self.n_fft, "win_size must not exceed n_fft"
            self.hann_window[str(device)] = win.to(device)
        return self.mel_basis[key], self._bands[key], self.hann_window[str(device)]

    def frames(self, n_samples):
        cfg = _lib.MelConfig(self.n_fft, self.hop_length, self.n_mels, self.clip_val, 1.0)
        return int(_lib.load().dsvc_mel_frames(cfg, int(n_samples)))

    def get_mel(self, y, center=False, out_scale=1.0, transpose=True):
        """y: CUDA fp32 [B, n_samples] in [-1, 1] -> natural-log mel [B, n_mels, frames] (nvSTFT.py:72-104).

        `out_scale` / `transpose=False` expose the kernel's native output ([B, frames, n_mels] times out_scale),
        which is what `wav2spec` wants."""
        assert not center, "the reference only uses center=False"
        if not y.is_cuda:
            raise _lib.DsvcError("STFT.get_mel: the mel analysis kernel has no CPU path; pass a CUDA tensor")
        if torch.min(y) < -1.:
            print('min value is ', torch.min(y))
        if torch.max(y) > 1.:
            print('max value is ', torch.max(y))
        lib = _lib.load()
        basis, (lo, hi), win = self._constants(y.device)
        y = y.contiguous().float()
        B, n = y.shape
        cfg = _lib.MelConfig(self.n_fft, self.hop_length, self.n_mels, float(self.clip_val), float(out_scale))
        T = int(lib.dsvc_mel_frames(cfg, n))
        if T < 0:
            raise _lib.DsvcError("STFT.get_mel: %d samples cannot be reflect-padded by %d" %
                                 (n, (self.n_fft - self.hop_length) // 2))
        out = torch.empty(B, T, self.n_mels, device=y.device, dtype=torch.float32)
        if T == 0:
            return out.transpose(1, 2) if transpose else out
        with torch.cuda.device(y.device):
            for b in range(B):
                _lib.check(lib.dsvc_mel_analysis(cfg, _lib.dptr(y[b]), n, _lib.dptr(win), _lib.dptr(basis),
                                                 _lib.dptr(lo), _lib.dptr(hi), _lib.dptr(out[b]),
                                                 _lib.current_stream()))
        return out.transpose(1, 2) if transpose else out

    def __call__(self, audiopath):
        audio, sr = load_wav_to_torch(audiopath, target_sr=self.target_sr)
        return self.get_mel(audio.unsqueeze(0).cuda()).squeeze(0)
