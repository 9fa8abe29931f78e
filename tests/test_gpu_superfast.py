"""GPU parity of CombSubSuperFast (fused comb source + STFT filtering + iSTFT kernel) against the
live-reference goldens and the oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from ddsp_svc_b200 import CombSubSuperFast, FixedControls, ops, synthetic as syn
from tests import report, util
from tests.golden import cases as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SR, P = G.SR, G.P
OFFICIAL_RMS = 1e-4
GATE_RMS = 2e-6


def _run(inp, noise=True):
    case = inp["case"]
    B, nF = case["B"], case["nF"]
    ctrls = syn.split_views(inp["dense"].to(DEV), G.split_map(case))
    hidden = torch.zeros(B, nF, 256, device=DEV)
    model = CombSubSuperFast(SR, P, case["win"], unit2ctrl=FixedControls(ctrls, hidden)).to(DEV)
    with torch.no_grad():
        out = model(None, inp["f0"].to(DEV), None, noise=inp["noise"].to(DEV) if noise else None)
    return model, out


@pytest.mark.parametrize("name", [n for n, c in G.CASES.items() if c["kind"] == "superfast"])
def test_superfast_forward_matches_reference_golden(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    model, (signal, hidden, (s1, s2)) = _run(inp)
    assert s1 is signal and s2 is signal                     # the reference returns the same tensor 3x
    got = signal.cpu().numpy()
    assert got.shape == gold["signal"].shape
    e, m = util.rms(got - gold["signal"]), np.abs(got - gold["signal"]).max()
    d = model.unit2ctrl.last_phase_frames.cpu().numpy() - gold["phase_frames"]
    d = (d + np.pi) % (2 * np.pi) - np.pi
    report.record("superfast_forward/" + name, signal_err=e, signal_max=m, signal_rms=util.rms(gold["signal"]),
                  phase_frames_max=np.abs(d).max())
    assert np.abs(d).max() < 2e-6
    assert e < OFFICIAL_RMS and e < GATE_RMS
    assert m < 5e-5


def test_superfast_state_dict_matches_reference_layout():
    m = CombSubSuperFast(SR, P, 2048, unit2ctrl=FixedControls())
    sd = m.state_dict()
    assert set(sd) == {"sampling_rate", "block_size", "win_length", "window"}
    assert sd["window"].shape == (2048,) and sd["sampling_rate"].dim() == 0


def test_superfast_long_utterance_and_chunking():
    """Longer than one CTA chunk (G hops) so chunk boundaries, ring wrap-around and the recomputed
    overlap frames are exercised; compared with the oracle's torch port."""
    from oracle import torch_port as tp
    B, nF = 2, 100
    sm = syn.superfast_split_map(2048)
    f0 = syn.make_f0(B, nF, SR, P, seed=5, unvoiced_fraction=0.05)
    dense, ctrls = syn.make_ctrl(B, nF, sm, seed=6)
    noise = syn.normal_noise((B, nF * P), 8)
    with torch.no_grad():
        ref = tp.superfast_forward(f0, ctrls, SR, P, 2048, noise=noise)
    ws, pf = ops.superfast_scan(f0.to(DEV), P, SR)
    dc = syn.split_views(dense.to(DEV), sm)
    sig = ops.superfast_synth(ws, dc["harmonic_magnitude"], dc["harmonic_phase"], dc["noise_magnitude"],
                              dc["noise_phase"], P, 2048, noise_in=noise.to(DEV)).cpu()
    e = util.rms(sig - ref["signal"])
    report.record("superfast_long", err=e, max=(sig - ref["signal"]).abs().max().item(), rms=util.rms(ref["signal"]))
    assert e < GATE_RMS


def test_superfast_in_kernel_noise_and_full_size():
    """BASELINE config 3 shape (B=32 x 10 s): finite output, deterministic per seed, shard invariant,
    and linear in the noise filter gain (size-independent property)."""
    B, nF = 32, 861
    sm = syn.superfast_split_map(2048)
    f0 = syn.make_f0(B, nF, SR, P).to(DEV)
    dense, _ = syn.make_ctrl(B, nF, sm)
    dc = syn.split_views(dense.to(DEV), sm)
    ws, _ = ops.superfast_scan(f0, P, SR)
    run = lambda seed, **kw: ops.superfast_synth(ws, dc["harmonic_magnitude"], dc["harmonic_phase"],
                                                 dc["noise_magnitude"], dc["noise_phase"], P, 2048, seed=seed, **kw)
    a, b_ = run(3), run(3)
    assert torch.isfinite(a).all() and torch.equal(a, b_)
    assert not torch.equal(a, run(4))
    # shard invariance: utterances 8.. computed alone with utterance_offset=8
    ws8, _ = ops.superfast_scan(f0[8:], P, SR)
    d8 = syn.split_views(dense[8:].to(DEV), sm)
    c = ops.superfast_synth(ws8, d8["harmonic_magnitude"], d8["harmonic_phase"], d8["noise_magnitude"],
                            d8["noise_phase"], P, 2048, seed=3, utterance_offset=8)
    assert torch.equal(a[8:], c)
    report.record("superfast_full", rms=a.pow(2).mean().sqrt().item())


def test_superfast_full_size_sampled_rows_match_oracle():
    """BASELINE config 3 shape with explicit noise: two sampled utterances of the B=32 x 10 s batch against the
    oracle port (the whole batch is too slow for the CPU oracle in seconds), as tests/test_gpu_sins.py does for Sins."""
    from oracle import torch_port as tp
    B, nF = 32, 861
    sm = syn.superfast_split_map(2048)
    f0 = syn.make_f0(B, nF, SR, P, unvoiced_fraction=0.03)
    dense, ctrls = syn.make_ctrl(B, nF, sm)
    rows = (5, 29)
    noise = torch.zeros(B, nF * P)
    for r in rows:
        noise[r] = syn.normal_noise((1, nF * P), 100 + r)[0]
    ws, _ = ops.superfast_scan(f0.to(DEV), P, SR)
    dc = syn.split_views(dense.to(DEV), sm)
    sig = ops.superfast_synth(ws, dc["harmonic_magnitude"], dc["harmonic_phase"], dc["noise_magnitude"],
                              dc["noise_phase"], P, 2048, noise_in=noise.to(DEV))
    assert torch.isfinite(sig).all()
    for r in rows:
        with torch.no_grad():
            ref = tp.superfast_forward(f0[r:r + 1], {k: v[r:r + 1] for k, v in ctrls.items()}, SR, P, 2048,
                                       noise=noise[r:r + 1])
        e = util.rms(sig[r:r + 1].cpu() - ref["signal"])
        report.record("superfast_full_row%d" % r, err=e, rms=util.rms(ref["signal"]))
        assert e < GATE_RMS
