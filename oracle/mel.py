"""CPU restatement of the NSF-HiFiGAN mel front end: STFT.get_mel, reference nsf_hifigan/nvSTFT.py:73-117, on the same
ATen operators, plus librosa's mel filterbank.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the product.

librosa (``librosa.filters.mel``) is a third-party dependency of the reference, unpinned in its requirements.txt and not
installed here (nor is soundfile, so nvSTFT.py cannot be imported as is): ``librosa_mel`` restates the published algorithm
(Slaney mel scale, triangular filters, 'slaney' area normalisation; librosa 0.8-0.10 are identical for these arguments)
and tests/test_oracle_mel.py cross-checks it against torchaudio's independent implementation of the same definition.
``load_reference_stft`` imports the reference's nvSTFT.py with librosa / soundfile stubbed by that restatement, so the
get_mel restated here is pinned on the reference's own code (tests + tests/golden/mel_*.npz).
"""
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


def _hz_to_mel(freq):
    freq = np.asanyarray(freq, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    mels = (freq - f_min) / f_sp
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if freq.ndim:
        log_t = freq >= min_log_hz
        mels[log_t] = min_log_mel + np.log(freq[log_t] / min_log_hz) / logstep
    elif freq >= min_log_hz:
        mels = min_log_mel + np.log(freq / min_log_hz) / logstep
    return mels


def _mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=np.float64)
    f_min, f_sp = 0.0, 200.0 / 3
    freqs = f_min + f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = (min_log_hz - f_min) / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * np.exp(logstep * (mels[log_t] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **unused):
    """librosa.filters.mel(htk=False, norm='slaney', dtype=float32): loop form of the published code."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    fftfreqs = np.linspace(0, float(sr) / 2, int(1 + n_fft // 2), endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def get_mel(y, sr=44100, n_mels=128, n_fft=2048, win_size=2048, hop_length=512, fmin=40, fmax=16000, clip_val=1e-5,
            keyshift=0, speed=1, center=False):
    """STFT.get_mel, nvSTFT.py:73-117.  y [B, T] torch CPU fp32 -> [B, n_mels, n_frames]."""
    factor = 2 ** (keyshift / 12)
    n_fft_new = int(np.round(n_fft * factor))
    win_size_new = int(np.round(win_size * factor))
    hop_length_new = int(np.round(hop_length * speed))
    mel_basis = torch.from_numpy(librosa_mel(sr=sr, n_fft=n_fft, n_mels=n_mels, fmin=fmin, fmax=fmax)).float()
    window = torch.hann_window(win_size_new)
    pad_left = (win_size_new - hop_length_new) // 2
    pad_right = max((win_size_new - hop_length_new + 1) // 2, win_size_new - y.size(-1) - pad_left)
    mode = "reflect" if pad_right < y.size(-1) else "constant"
    y = F.pad(y.unsqueeze(1), (pad_left, pad_right), mode=mode).squeeze(1)
    spec = torch.stft(y, n_fft_new, hop_length=hop_length_new, win_length=win_size_new, window=window, center=center,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + (1e-9))
    if keyshift != 0:
        size = n_fft // 2 + 1
        resize = spec.size(1)
        if resize < size:
            spec = F.pad(spec, (0, 0, 0, size - resize))
        spec = spec[:, :size, :] * win_size / win_size_new
    spec = torch.matmul(mel_basis, spec)
    return torch.log(torch.clamp(spec, min=clip_val))


def load_reference_stft():
    """The reference's own nsf_hifigan/nvSTFT.py module with librosa / soundfile (absent here) stubbed: ``librosa_mel_fn``
    resolves to the restatement above, everything else in STFT.get_mel is the reference's code on torch."""
    from oracle import ref_loader
    ref_loader.load()
    if "librosa" not in sys.modules or not hasattr(sys.modules["librosa"], "__b2d_stub__"):
        try:
            import librosa  # noqa: F401  (a real librosa, if ever present, wins)
        except Exception:
            lib = types.ModuleType("librosa"); lib.__b2d_stub__ = True
            util = types.ModuleType("librosa.util"); util.normalize = lambda x, *a, **k: x
            filt = types.ModuleType("librosa.filters"); filt.mel = librosa_mel
            core = types.ModuleType("librosa.core"); core.resample = None
            lib.util, lib.filters, lib.core = util, filt, core
            sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt, "librosa.core": core})
    if "soundfile" not in sys.modules:
        try:
            import soundfile  # noqa: F401
        except Exception:
            sys.modules["soundfile"] = types.ModuleType("soundfile")
    import nsf_hifigan.nvSTFT as ref_stft
    return ref_stft
