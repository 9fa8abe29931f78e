"""GPU parity of the caller-side prologue / epilogue kernels (csrc/frontend.cu) against the live-reference fixtures and
the numpy restatement, through the C ABI."""
import glob
import os

import numpy as np
import pytest
import torch

from ddsp_svc_b200 import Volume_Extractor, frontend as fr
from oracle import frontend as fe
from tests import report

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden", "frontend_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_kernels_match_the_reference_fixtures(path):
    z = np.load(path)
    hop = int(z["hop"])
    audio = torch.from_numpy(z["audio"]).to(DEV)
    vol = fr.volume_extract(audio[None], hop)
    e_vol = float(np.abs(vol[0].cpu().numpy() - z["volume"]).max() / max(z["volume"].max(), 1e-12))
    mask = fr.volume_mask(torch.from_numpy(z["volume"]).to(DEV)[None], -40)           # the reference's volume: exact mask
    assert np.array_equal(mask[0].cpu().numpy(), z["mask"].astype(np.float32))
    sig = torch.ones(1, z["mask_up"].shape[1], device=DEV)
    fr.mask_apply_(sig, mask, hop)
    e_up = float(np.abs(sig.cpu().numpy() - z["mask_up"]).max())
    out = fr.cross_fade(torch.from_numpy(z["fade_a"]).to(DEV), torch.from_numpy(z["fade_b"]).to(DEV), int(z["fade_idx"]))
    e_cf = float(np.abs(out.cpu().numpy().astype(np.float64) - z["fade_out"]).max())
    report.record("frontend/" + os.path.basename(path)[:-4], volume_rel=e_vol, upsample_max=e_up, cross_fade_max=e_cf)
    assert e_vol < 2e-6           # numpy's float32 pairwise mean vs an fp64-accumulated mean
    assert e_up == 0.0            # the kernel reproduces torch's interpolation weights operation by operation
    assert e_cf < 1.5e-7          # the reference keeps float64; this is the fp32 rounding of its values (|x| <= ~4)


def test_volume_extractor_drop_in_contract_and_batches():
    g = np.random.default_rng(3)
    audio = (0.1 * g.standard_normal(12345)).astype(np.float32)
    ve = Volume_Extractor(512)
    v = ve.extract(audio)                                   # numpy in -> numpy out, like the reference
    assert isinstance(v, np.ndarray) and v.shape == (12345 // 512 + 1,) and v.dtype == np.float32
    want = fe.volume_extract(audio, 512)
    assert np.abs(v - want).max() < 2e-6 * want.max()
    batch = torch.from_numpy(np.stack([audio, audio[::-1].copy(), 2 * audio])).to(DEV)
    vb = ve.extract(batch)
    assert vb.is_cuda and vb.shape == (3, v.shape[0])
    assert torch.equal(vb[0].cpu(), torch.from_numpy(v))
    assert np.abs(vb[2].cpu().numpy() - 2 * want).max() < 4e-6 * want.max()
    with pytest.raises(ValueError):
        fr.volume_extract(torch.zeros(1, 100, device=DEV), 512)       # reflect padding needs hop/2 < T, like numpy


def test_mask_apply_segments_and_cross_fade_edges():
    """main.py:248-277 flow: a global frame-rate mask, segments multiplied by their slice of its upsampling in place,
    then joined with cross-fades; compared with the numpy / torch restatement."""
    g = torch.Generator().manual_seed(5)
    P, nF = 512, 40
    mask = (torch.rand(1, nF, generator=g) > 0.4).float()
    segs = [(3, 10), (13, 20), (30, 10)]
    for start, n in segs:
        seg = torch.randn(1, n * P, generator=g)
        want = fe.mask_apply(seg, mask[0].numpy(), P, start)
        got = fr.mask_apply_(seg.to(DEV).clone(), mask.to(DEV), P, frame_offset=start)
        assert torch.equal(got.cpu(), want)
    with pytest.raises(ValueError):
        fr.mask_apply_(torch.zeros(1, 5 * P, device=DEV), mask.to(DEV), P, frame_offset=38)   # runs past the mask
    a, b = torch.randn(300, generator=g), torch.randn(200, generator=g)
    for idx in (299, 150, 100):                                   # fade lengths 1, 150, 200 (= all of b)
        want = fe.cross_fade(a.numpy(), b.numpy(), idx)
        got = fr.cross_fade(a.to(DEV), b.to(DEV), idx).cpu().numpy()
        assert got.shape == want.shape and np.abs(got - want).max() < 3e-7
    with pytest.raises(ValueError):
        fr.cross_fade(a.to(DEV), b.to(DEV), 50)                   # fade of 250 samples > len(b): the reference would fail too
