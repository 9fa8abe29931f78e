"""Drop-in for the NSF-HiFiGAN sine generator (reference nsf_hifigan/models.py:101-165)."""
import torch

from . import ops
from .vocoder import _host_seed


class SineGen(torch.nn.Module):
    """SineGen(samp_rate, harmonic_num=0, sine_amp=0.1, noise_std=0.003, voiced_threshold=0)

    forward(f0 [B, n_frames], upp) -> [B, n_frames*upp, harmonic_num+1].  No parameters or
    buffers, like the reference.  Seeding contract: the in-kernel Philox stream is keyed by one 62-bit seed drawn per
    forward from torch's CPU generator (``torch.manual_seed`` makes runs reproducible; ``torch.cuda.manual_seed`` has no
    effect) -- drawing it on the host keeps the call free of device syncs.  The random initial phases are drawn with torch.rand on the
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


class SourceModuleHnNSF(torch.nn.Module):
    """Drop-in for nsf_hifigan.models.SourceModuleHnNSF (models.py:168-204): SineGen followed by
    tanh(Linear(harmonic_num+1 -> 1)), executed as ONE kernel that never writes the [B, T, dim] sines.
    Same constructor (including the reference's ``voiced_threshod`` spelling) and state-dict keys
    (``l_linear.weight``, ``l_linear.bias``)."""

    def __init__(self, sampling_rate, harmonic_num=0, sine_amp=0.1, add_noise_std=0.003, voiced_threshod=0):
        super().__init__()
        self.sine_amp = sine_amp
        self.noise_std = add_noise_std
        self.l_sin_gen = SineGen(sampling_rate, harmonic_num, sine_amp, add_noise_std, voiced_threshod)
        self.l_linear = torch.nn.Linear(harmonic_num + 1, 1)
        self.l_tanh = torch.nn.Tanh()
        self.__dict__["_bias_cache"] = None

    def _bias(self):
        """The Linear bias as a host float for the kernel's scalar argument.  Read back from the device only when the
        parameter changed (in-place version counter / storage), not on every forward: no per-call device sync, and the
        forward stays capturable in a CUDA graph once warmed up."""
        b = self.l_linear.bias
        key = (b._version, b.data_ptr())
        c = self.__dict__.get("_bias_cache")
        if c is None or c[0] != key:
            c = (key, float(b.detach().reshape(-1)[0]))
            self.__dict__["_bias_cache"] = c
        return c[1]

    def forward(self, x, upp, rand_ini=None, noise=None, utterance_offset=0):
        g = self.l_sin_gen
        if torch.is_grad_enabled() and self.l_linear.weight.requires_grad and self.training:
            raise NotImplementedError("the fused source module is forward-only; use it under torch.no_grad()/eval()")
        if rand_ini is None:
            rand_ini = torch.rand(1, 1, g.dim, device=x.device)
            rand_ini[..., 0] = 0
        return ops.source_module(x, int(upp), g.sampling_rate, g.dim, rand_ini, self.l_linear.weight,
                                 self._bias(), g.sine_amp, g.noise_std,
                                 g.voiced_threshold, noise_in=noise, seed=0 if noise is not None else _host_seed(),
                                 utterance_offset=utterance_offset)
