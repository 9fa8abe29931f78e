"""GPU parity of the CombSubFast kernel against the live-reference goldens and the fp64 closed form.

The same kernel source is also executed on the CPU by tests/emu/ (tests/test_emu_combsubfast.py)."""
import numpy as np
import pytest
import torch

from ddsp_svc_b200 import CombSubFast, FixedControls, ops, synthetic as syn
from tests import report, util
from tests.golden import cases as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SR, P = G.SR, G.P
OFFICIAL_RMS, GATE_RMS = 1e-4, 2e-6


@pytest.mark.parametrize("name", [n for n, c in G.CASES.items() if c["kind"] == "combsubfast"])
def test_combsubfast_matches_reference_golden(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    ctrl = {k: v.to(DEV) for k, v in syn.split_views(inp["dense"].to(DEV), G.split_map(inp["case"])).items()}
    fixed = FixedControls(ctrl, torch.zeros(1, device=DEV))
    model = CombSubFast(SR, P, unit2ctrl=fixed).to(DEV)
    kw = {"initial_phase": inp["initial_phase"].to(DEV)} if "initial_phase" in inp else {}
    with torch.no_grad():
        signal, _, (s2, s3) = model(None, inp["f0"].to(DEV), None, noise=inp["noise"].to(DEV), **kw)
    assert s2 is signal and s3 is signal
    got = signal.cpu().numpy()
    assert got.shape == gold["signal"].shape
    e, m = util.rms(got - gold["signal"]), np.abs(got - gold["signal"]).max()
    pf = fixed.last_phase_frames.cpu().numpy()
    report.record("combsubfast/" + name, rms=e, max=m, ref_rms=util.rms(gold["signal"]),
                  phase_max=np.abs(pf - gold["phase_frames"]).max())
    assert e < OFFICIAL_RMS and e < GATE_RMS
    assert np.abs(pf - gold["phase_frames"]).max() < 2e-6


def test_combsubfast_chunk_boundaries_and_truth():
    """70 frames = two full 32-hop chunks + a ragged one; against the fp64 closed form"""
    from oracle import closed_form as cf
    B, nF = 2, 70
    f0 = syn.make_f0(B, nF, SR, P, seed=5, unvoiced_fraction=0.2)
    dense, views = syn.make_ctrl(B, nF, syn.combsubfast_split_map(P), seed=6)
    noise = syn.uniform_noise(B, nF * P, 9)
    truth = cf.combsubfast(f0.numpy(), {k: v.numpy() for k, v in views.items()}, SR, P, noise.numpy())
    ctrl = syn.split_views(dense.to(DEV), syn.combsubfast_split_map(P))
    model = CombSubFast(SR, P, unit2ctrl=FixedControls(ctrl, None)).to(DEV)
    with torch.no_grad():
        got = model(None, f0.to(DEV), None, noise=noise.to(DEV))[0].cpu().numpy()
    e = util.rms(got - truth["signal"])
    report.record("combsubfast_truth", rms=e, max=np.abs(got - truth["signal"]).max(), ref_rms=util.rms(truth["signal"]))
    assert e < GATE_RMS


def test_combsubfast_in_kernel_noise_is_shard_invariant():
    B, nF = 3, 12
    f0 = syn.make_f0(B, nF, SR, P, seed=2).to(DEV)
    dense, _ = syn.make_ctrl(B, nF, syn.combsubfast_split_map(P), seed=3)
    dense = dense.to(DEV)
    fp, _ = ops.phase_scan(f0, P, SR)
    comb = ops.comb_source(f0, fp, P, SR)
    c = syn.split_views(dense, syn.combsubfast_split_map(P))
    full = ops.combsubfast_filter(comb, c["harmonic_magnitude"], c["harmonic_phase"], c["noise_magnitude"], P, seed=4)
    c1 = syn.split_views(dense[1:], syn.combsubfast_split_map(P))
    part = ops.combsubfast_filter(comb[1:], c1["harmonic_magnitude"], c1["harmonic_phase"], c1["noise_magnitude"], P,
                                  seed=4, utterance_offset=1)
    assert torch.equal(full[1:], part)
    assert full.abs().max().item() < 10 and torch.isfinite(full).all()
