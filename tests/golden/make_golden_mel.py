"""Generate tests/golden/mel_*.npz with the reference's OWN nsf_hifigan/nvSTFT.py (build container only):

    python tests/golden/make_golden_mel.py

nvSTFT.py imports librosa and soundfile, which are not installed: oracle.mel.load_reference_stft() stubs them (the mel
filterbank function resolves to the restatement of librosa's published algorithm, cross-checked against torchaudio in
tests/test_oracle_mel.py); padding, torch.stft, magnitude, projection and log are the reference's code."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mel as om  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {   # name: (seed, B, T, hop)
    "mel_b2_f12": (1, 2, 512 * 12, 512),
    "mel_b1_ragged": (2, 1, 512 * 7 + 301, 512),
    "mel_b1_short_constpad": (3, 1, 700, 512),        # pad_right >= T: the 'constant' padding branch (nvSTFT.py:99-102)
    "mel_b1_hop256": (4, 1, 256 * 21, 256),
}


def signal(seed, B, T):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float64) / 44100
    f = 110.0 * 2 ** (3 * torch.rand(B, 1, generator=g, dtype=torch.float64))
    y = sum((0.3 / h) * torch.sin(2 * np.pi * h * f * t) for h in (1, 2, 3, 5, 8)) + 0.01 * torch.randn(B, T, generator=g, dtype=torch.float64)
    return y.float()


def main():
    ref = om.load_reference_stft()
    for name, (seed, B, T, hop) in CASES.items():
        st = ref.STFT(44100, 128, 2048, 2048, hop, 40, 16000)
        y = signal(seed, B, T)
        with torch.no_grad():
            mel = st.get_mel(y)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), y=y.numpy(), hop=np.int64(hop), mel=mel.numpy())
        print(name, tuple(mel.shape), float(mel.min()), float(mel.max()))


if __name__ == "__main__":
    main()
