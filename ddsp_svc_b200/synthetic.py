"""Seeded synthetic control tensors for parity tests and the benchmark.

The hot path consumes frame-rate controls (f0 + the raw, pre-activation outputs
of the reference's Unit2Control) and produces audio.  There is no network in the
build, so every workload uses synthetic controls of the reference's shapes:

* f0: per-utterance base pitch ``110 * 2**(2u)`` Hz, u ~ U(0,1), with a 5.5 Hz,
  +-3 % vibrato sampled at the frame rate;
* raw controls ~ N(mu, sigma) per split, laid out as ONE dense
  ``[B, n_frames, n_out]`` tensor whose splits are strided views -- exactly how
  ``split_to_dict`` (reference ddsp/unit2control.py:12-23) hands them to the
  synthesizer.

Everything is generated on the CPU with explicit generators so the CPU oracle and
the CUDA path see identical bits.
"""
import math
from collections import OrderedDict

import torch

# (mean, std) of the raw controls, chosen so output RMS is ~0.006-0.01.
CTRL_STATS = {
    "amplitudes": (-2.0, 0.5),
    "group_delay": (0.0, 0.3),
    "noise_magnitude": (-3.0, 0.5),
    "harmonic_magnitude": (-2.0, 0.5),
    "harmonic_phase": (0.0, 0.3),
    "noise_phase": (0.0, 0.3),
}


def n_frames_for(seconds, sampling_rate=44100, block_size=512):
    """Frame convention of the reference data loader (data_loaders.py:198)."""
    return int(seconds * sampling_rate / block_size)


def make_f0(batch, n_frames, sampling_rate=44100, block_size=512, seed=1234,
            unvoiced_fraction=0.0, sweep_row=None):
    """f0_frames [B, n_frames, 1] fp32 (Hz)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(batch, generator=g, dtype=torch.float64)
    base = 110.0 * torch.pow(torch.tensor(2.0, dtype=torch.float64), 2.0 * u)
    k = torch.arange(n_frames, dtype=torch.float64)
    vib = 1.0 + 0.03 * torch.sin(2.0 * math.pi * 5.5 * k * block_size / sampling_rate)
    f0 = base[:, None] * vib[None, :]
    if sweep_row is not None and batch > sweep_row:
        # one utterance sweeping 65 -> 1100 Hz: exercises the Nyquist mask
        f0[sweep_row] = torch.exp(torch.linspace(math.log(65.0), math.log(1100.0), n_frames,
                                                 dtype=torch.float64))
    if unvoiced_fraction > 0:
        m = torch.rand(batch, n_frames, generator=g) < unvoiced_fraction
        f0 = f0.masked_fill(m, 0.0)
    return f0.to(torch.float32).unsqueeze(-1).contiguous()


def make_ctrl(batch, n_frames, split_map, seed=7):
    """Dense raw control tensor [B, n_frames, n_out] + dict of strided split views."""
    g = torch.Generator().manual_seed(seed)
    n_out = sum(split_map.values())
    dense = torch.empty(batch, n_frames, n_out, dtype=torch.float32)
    off = 0
    for name, width in split_map.items():
        mu, sd = CTRL_STATS[name]
        dense[:, :, off:off + width] = torch.randn(batch, n_frames, width, generator=g) * sd + mu
        off += width
    return dense, split_views(dense, split_map)


def split_views(dense, split_map):
    views = torch.split(dense, list(split_map.values()), dim=-1)
    return OrderedDict(zip(split_map.keys(), views))


def sins_split_map(n_harmonics=128, n_mag_allpass=256, n_mag_noise=256):
    return OrderedDict([("amplitudes", n_harmonics), ("group_delay", n_mag_allpass),
                        ("noise_magnitude", n_mag_noise)])


def combsub_split_map(n_mag_allpass=256, n_mag_harmonic=512, n_mag_noise=256):
    return OrderedDict([("group_delay", n_mag_allpass), ("harmonic_magnitude", n_mag_harmonic),
                        ("noise_magnitude", n_mag_noise)])


def superfast_split_map(win_length=2048):
    n = win_length // 2 + 1
    return OrderedDict([("harmonic_magnitude", n), ("harmonic_phase", n),
                        ("noise_magnitude", n), ("noise_phase", n)])


def combsubfast_split_map(block_size=512):
    """CombSubFast predicts block_size+1 bins per control (reference ddsp/vocoder.py:728-732)."""
    n = block_size + 1
    return OrderedDict([("harmonic_magnitude", n), ("harmonic_phase", n), ("noise_magnitude", n)])


def uniform_noise(batch, n_samples, seed):
    """What ``torch.rand_like(x) * 2 - 1`` yields in the reference (ddsp/vocoder.py:603)
    when ``torch.manual_seed(seed)`` was called just before forward()."""
    torch.manual_seed(seed)
    return torch.rand(batch, n_samples) * 2 - 1


def normal_noise(shape, seed):
    """``torch.randn_like`` under ``torch.manual_seed(seed)`` (ddsp/vocoder.py:687)."""
    torch.manual_seed(seed)
    return torch.randn(*shape)
