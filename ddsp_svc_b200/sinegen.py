"""Drop-in for the NSF-HiFiGAN sine generator (reference nsf_hifigan/models.py:101-165)."""
import torch

from . import ops
from .vocoder import _host_seed


class SineGen(torch.nn.Module):
    """SineGen(samp_rate, harmonic_num=0, sine_amp=0.1, noise_std=0.003, voiced_threshold=0)

    forward(f0 [B, n_frames], upp) -> [B, n_frames*upp, harmonic_num+1].  No parameters or
    buffers, like the reference.  The random initial phases are drawn with torch.rand on the
    input's device exactly like the reference (models.py:144); the additive Gaussian noise comes
    from the in-kernel Philox generator unless ``noise`` is given.
    """

    def __init__(self, samp_rate, harmonic_num=0, sine_amp=0.1, noise_std=0.003, voiced_threshold=0):
        super().__init__()
        self.sine_amp = sine_amp
        self.noise_std = noise_std
        self.harmonic_num = harmonic_num
        self.dim = self.harmonic_num + 1
        self.sampling_rate = samp_rate
        self.voiced_threshold = voiced_threshold

    @torch.no_grad()
    def forward(self, f0, upp, rand_ini=None, noise=None, utterance_offset=0):
        if rand_ini is None:
            rand_ini = torch.rand(1, 1, self.dim, device=f0.device)
            rand_ini[..., 0] = 0
        return ops.sinegen(f0, int(upp), self.sampling_rate, self.dim, rand_ini, self.sine_amp, self.noise_std,
                           self.voiced_threshold, noise_in=noise, seed=0 if noise is not None else _host_seed(),
                           utterance_offset=utterance_offset)
