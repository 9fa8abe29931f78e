"""GPU parity of the mel front-end kernel (csrc/mel.cu, ddsp_svc_b200.mel.STFT.get_mel) against fixtures produced by the
reference's nvSTFT.py and against the oracle restatement, through the C ABI."""
import glob
import os

import numpy as np
import pytest
import torch

from ddsp_svc_b200 import mel as pm
from oracle import mel as om
from tests import report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "mel_*.npz")))
# log-mel values span about [-11.5, 3]; fp32 FFT round-off of the 2048-point transforms shows up as ~1e-5 absolute in the
# log domain for bins far below the frame's peak (relative error of a tiny magnitude)
TOL_MAX, TOL_RMS = 2e-3, 5e-5


def _stft(hop):
    return pm.STFT(44100, 128, 2048, 2048, hop, 40, 16000)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_mel_matches_reference_fixture(path):
    z = np.load(path)
    got = _stft(int(z["hop"])).get_mel(torch.from_numpy(z["y"]).to(DEV)).cpu().numpy()
    assert got.shape == z["mel"].shape
    d = got.astype(np.float64) - z["mel"]
    report.record("mel/" + os.path.basename(path)[:-4], max=float(np.abs(d).max()), rms=float(np.sqrt((d ** 2).mean())))
    assert np.abs(d).max() < TOL_MAX and np.sqrt((d ** 2).mean()) < TOL_RMS


def test_mel_full_size_rows_and_contract():
    """B = 32 x 10 s (the synthesizer's BASELINE batch): two sampled utterances against the oracle, determinism, and the
    unsupported shapes raise instead of silently falling back."""
    B, T = 32, 861 * 512
    g = torch.Generator().manual_seed(12)
    y = 0.1 * torch.randn(B, T, generator=g)
    st = _stft(512)
    a = st.get_mel(y.to(DEV))
    assert a.shape == (B, 128, 861) and torch.isfinite(a).all()
    assert torch.equal(a, st.get_mel(y.to(DEV)))
    for r in (0, 19):
        with torch.no_grad():
            want = om.get_mel(y[r:r + 1])
        d = (a[r:r + 1].cpu() - want).double()
        report.record("mel/full_row%d" % r, max=d.abs().max().item(), rms=d.pow(2).mean().sqrt().item())
        assert d.abs().max().item() < TOL_MAX and d.pow(2).mean().sqrt().item() < TOL_RMS
    with pytest.raises(NotImplementedError):
        st.get_mel(y[:1].to(DEV), keyshift=2)
    with pytest.raises(NotImplementedError):
        pm.STFT(22050, 80, 1024, 1024, 256, 20, 11025).get_mel(y[:1].to(DEV))
    with pytest.raises(ValueError):
        st.get_mel(y[:1])                                        # CPU tensor: no fallback
