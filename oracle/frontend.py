"""CPU restatement (numpy) of the caller-side prologue / epilogue around the synthesis path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the product.

Each function follows the reference line by line; tests/test_oracle_frontend.py pins them against the live
reference (Volume_Extractor imported from ddsp/vocoder.py, cross_fade compiled from main.py's own source) whenever the
reference is present, and tests/golden/frontend_*.npz holds outputs of the live reference for the GPU box.
"""
import numpy as np
import torch


def volume_extract(audio, hop_size=512):
    """Volume_Extractor.extract, reference ddsp/vocoder.py:150-157.  audio: 1-D numpy array."""
    n_frames = int(len(audio) // hop_size) + 1
    audio2 = audio ** 2
    audio2 = np.pad(audio2, (int(hop_size // 2), int((hop_size + 1) // 2)), mode="reflect")
    volume = np.array([np.mean(audio2[int(n * hop_size): int((n + 1) * hop_size)]) for n in range(n_frames)])
    return np.sqrt(volume)


def volume_mask(volume, threshold_db=-60.0):
    """Silence mask at frame rate, reference main.py:211-213."""
    mask = (volume > 10 ** (float(threshold_db) / 20)).astype("float")
    mask = np.pad(mask, (4, 4), constant_values=(mask[0], mask[-1]))
    return np.array([np.max(mask[n: n + 9]) for n in range(len(mask) - 8)])


def upsample(signal, factor):
    """reference ddsp/core.py:66-70 on a [B, nF, C] torch tensor."""
    signal = signal.permute(0, 2, 1)
    signal = torch.nn.functional.interpolate(torch.cat((signal, signal[:, :, -1:]), 2), size=signal.shape[-1] * factor + 1,
                                             mode="linear", align_corners=True)
    signal = signal[:, :, :-1]
    return signal.permute(0, 2, 1)


def mask_apply(seg_output, mask_frames, block_size, start_frame=0):
    """main.py:213-215 + 260: mask [nF] -> torch [1, nF, 1] -> upsample -> seg_output *= mask[:, start*block : (start+n)*block]."""
    mask = torch.from_numpy(np.asarray(mask_frames)).float().unsqueeze(-1).unsqueeze(0)
    mask = upsample(mask, block_size).squeeze(-1)
    n = seg_output.shape[1] // block_size
    out = seg_output.clone()
    out *= mask[:, start_frame * block_size: (start_frame + n) * block_size]
    return out


def cross_fade(a, b, idx):
    """reference main.py:142-149."""
    result = np.zeros(idx + b.shape[0])
    fade_len = a.shape[0] - idx
    np.copyto(dst=result[:idx], src=a[:idx])
    k = np.linspace(0, 1.0, num=fade_len, endpoint=True)
    result[idx: a.shape[0]] = (1 - k) * a[idx:] + k * b[: fade_len]
    np.copyto(dst=result[a.shape[0]:], src=b[fade_len:])
    return result
