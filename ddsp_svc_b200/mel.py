"""Log-mel front end of the NSF-HiFiGAN vocoder on the GPU: drop-in for nsf_hifigan.nvSTFT.STFT (reference
nsf_hifigan/nvSTFT.py:59-122), the consumer of the synthesizer's waveform in enhancer.py:113, diffusion/vocoder.py:147
and reflow/vocoder.py:125.

``STFT.get_mel(y)`` runs ONE kernel (csrc/mel.cu: padding + Hann frames + 2048-point FFT + magnitude + sparse mel
projection + log) for the shape every shipped configuration uses -- keyshift 0, speed 1, n_fft = win_size = 2048.
Other shapes raise NotImplementedError (there is no PyTorch fallback in this package).

The mel filterbank is librosa's (``librosa.filters.mel``, Slaney scale and normalisation: a third-party dependency of
the reference, unpinned in requirements.txt and absent here); ``mel_filterbank`` restates its published algorithm.
"""
import numpy as np
import torch

from . import _lib
from .ops import _count, _need_cuda_f32, _stream


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, min_log_hz) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) with its defaults (htk=False, norm='slaney'):
    triangular filters with corners equally spaced on the Slaney mel scale, each normalised to unit area -> float32
    [n_mels, 1 + n_fft // 2]."""
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    fftfreqs = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


def _support(basis):
    """[n_mels, 2] int32: first and one-past-last non-zero bin of each filter (empty filters -> 0, 0)."""
    nz = basis != 0
    lo = np.where(nz.any(1), nz.argmax(1), 0)
    hi = np.where(nz.any(1), basis.shape[1] - nz[:, ::-1].argmax(1), 0)
    return np.stack([lo, hi], 1).astype(np.int32)


class STFT:
    def __init__(self, sr=22050, n_mels=80, n_fft=1024, win_size=1024, hop_length=256, fmin=20, fmax=11025, clip_val=1e-5):
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

    def _tables(self, device):
        key = str(self.fmax) + "_" + str(device)
        if key not in self.mel_basis:
            mel = mel_filterbank(self.target_sr, self.n_fft, self.n_mels, self.fmin, self.fmax)
            self.mel_basis[key] = (torch.from_numpy(mel).to(device), torch.from_numpy(_support(mel)).to(device))
            self.hann_window[key] = torch.hann_window(self.win_size).to(device)
        return self.mel_basis[key] + (self.hann_window[key],)

    def get_mel(self, y, keyshift=0, speed=1, center=False):
        """y [B, T] CUDA fp32 -> log-mel [B, n_mels, n_frames] (nvSTFT.py:73-117)."""
        if keyshift != 0 or speed != 1 or center:
            raise NotImplementedError("the B200 mel kernel covers keyshift=0, speed=1, center=False (the inference call of "
                                      "enhancer.py:113); got keyshift=%r speed=%r center=%r" % (keyshift, speed, center))
        if self.n_fft != 2048 or self.win_size != 2048 or self.n_mels > 128:
            raise NotImplementedError("the B200 mel kernel is built for n_fft = win_size = 2048 and n_mels <= 128 "
                                      "(the 44.1 kHz NSF-HiFiGAN configuration); got %d / %d / %d" % (self.n_fft, self.win_size, self.n_mels))
        _need_cuda_f32("y", y)
        if y.dim() != 2:
            raise ValueError("y must be [B, n_samples]")
        y = y.contiguous()
        B, T = y.shape
        L = _lib.lib()
        n_frames = L.b2d_mel_frames(T, self.n_fft, self.win_size, int(self.hop_length))
        if n_frames <= 0:
            raise ValueError("signal of %d samples is too short for one frame" % T)
        basis, lohi, window = self._tables(y.device)
        out = torch.empty(B, self.n_mels, n_frames, dtype=torch.float32, device=y.device)
        _lib.check(L.b2d_mel_spectrogram(y.data_ptr(), window.data_ptr(), basis.data_ptr(), lohi.data_ptr(), B, T, self.n_fft,
                                         self.win_size, int(self.hop_length), self.n_mels, float(self.clip_val), out.data_ptr(),
                                         _stream()), "b2d_mel_spectrogram")
        _count(1)
        return out
