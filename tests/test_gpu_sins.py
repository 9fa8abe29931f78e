"""GPU parity of the Sins path, stage by stage and end to end, through the C ABI
(ddsp_svc_b200.ops -> libb200ddsp.so), against the oracle and the live-reference goldens.

Tolerances: north star = 1e-4 RMS absolute on the waveforms; the internal gate used here is
2e-6 RMS (signal RMS is ~8e-3, so ~2.5e-4 relative) so that an indexing/windowing slip cannot
hide under the official bound.
"""
import numpy as np
import pytest
import torch

from ddsp_svc_b200 import FixedControls, Sins, ops, synthetic as syn
from tests import report, util
from tests.golden import cases as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SR, P = G.SR, G.P
OFFICIAL_RMS = 1e-4
GATE_RMS = 2e-6

SINS_CASES = [n for n, c in G.CASES.items() if c["kind"] == "sins"]


def _dev_ctrls(inp):
    dense = inp["dense"].to(DEV)
    return syn.split_views(dense, G.split_map(inp["case"]))


@pytest.mark.parametrize("name", SINS_CASES)
def test_phase_scan(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    ip = inp.get("initial_phase")
    fp, pf = ops.phase_scan(inp["f0"].to(DEV), P, SR, None if ip is None else ip.to(DEV))
    from oracle import closed_form as cf
    # frame_phase (unwrapped, fp64) vs the exact per-sample cumulative sum at frame starts
    f = inp["f0"].double().numpy()[..., 0]
    fe = np.concatenate([f, f[:, -1:]], 1)
    adv = (P * f + (fe[:, 1:] - fe[:, :-1]) * (P - 1) / 2) / SR
    S = np.concatenate([np.zeros((f.shape[0], 1)), np.cumsum(adv, 1)[:, :-1]], 1)
    if ip is not None:
        S = S + ip.double().numpy().reshape(-1, 1) / (2 * np.pi)
    err_s = np.abs(fp.cpu().numpy() - S).max()
    assert err_s < 1e-9
    if "phase_frames" in gold:
        d = pf.cpu().numpy() - gold["phase_frames"]
        d = (d + np.pi) % (2 * np.pi) - np.pi   # +-pi wrap ambiguity at exactly half a cycle
        report.record("phase_scan/" + name, frame_phase_max=err_s, phase_frames_max=np.abs(d).max())
        assert np.abs(d).max() < 2e-6


@pytest.mark.parametrize("name", ["sins_b2_f24_h128", "sins_b1_f7_h33", "sins_b1_f7_h1", "sins_b1_f7_h64",
                                  "sins_b1_f12_h40_m65_initphase"])
def test_bank_and_ir_stages(name):
    inp = G.build_inputs(name)
    ref = util.port_outputs(name, inp)
    c = _dev_ctrls(inp)
    f0 = inp["f0"].to(DEV)
    ip = inp.get("initial_phase")
    fp, _ = ops.phase_scan(f0, P, SR, None if ip is None else ip.to(DEV))
    sin_gpu = ops.sins_bank(f0, fp, c["amplitudes"], P, SR).cpu()
    e_bank = util.rms(sin_gpu - ref["sinusoids"])
    ir_ap = ops.ir_build(c["group_delay"], ops.IR_ALLPASS, SR).cpu()
    ir_n = ops.ir_build(c["noise_magnitude"], ops.IR_MAG_HANN, SR).cpu()
    e_ap = (ir_ap - ref["ir_allpass"]).abs().max().item()
    e_n = (ir_n - ref["ir_noise"]).abs().max().item()
    report.record("stages/" + name, bank_rms=e_bank, bank_ref_rms=util.rms(ref["sinusoids"]),
                  ir_allpass_max=e_ap, ir_allpass_rms=util.rms(ir_ap - ref["ir_allpass"]),
                  ir_noise_max=e_n, ir_noise_peak=ref["ir_noise"].abs().max().item())
    assert e_bank < 1e-6
    assert e_ap < 5e-5 and util.rms(ir_ap - ref["ir_allpass"]) < 5e-6
    assert e_n < 1e-7


@pytest.mark.parametrize("name", ["sins_b2_f24_h128", "sins_b1_f2_h128", "sins_b3_f1_h128",
                                  "sins_b1_f12_h40_m65_initphase"])
def test_fir_stage_tiled_vs_oracle_and_generic(name):
    """Feed the ORACLE's sinusoids and IRs to the FIR kernels: isolates the convolution."""
    inp = G.build_inputs(name)
    ref = util.port_outputs(name, inp)
    x = ref["sinusoids"].to(DEV)
    ir = ref["ir_allpass"].to(DEV).contiguous()
    y_t = ops.ltv_fir(x, ir, P).cpu()
    y_g = ops.ltv_fir(x, ir, P, generic=True).cpu()
    e_t, e_g = util.rms(y_t - ref["harmonic"]), util.rms(y_g - ref["harmonic"])
    report.record("fir/" + name, tiled_rms=e_t, generic_rms=e_g, tiled_vs_generic_max=(y_t - y_g).abs().max().item())
    assert e_t < 5e-7 and e_g < 5e-7


@pytest.mark.parametrize("name", SINS_CASES)
def test_sins_forward_matches_reference_golden(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    case = inp["case"]
    B, nF = case["B"], case["nF"]
    hidden = torch.zeros(B, nF, 256, device=DEV)
    model = Sins(SR, P, case["H"], case["Ma"], case["Mn"], unit2ctrl=FixedControls(_dev_ctrls(inp), hidden)).to(DEV)
    ip = inp.get("initial_phase")
    with torch.no_grad():
        signal, hid, (harm, nz) = model(None, inp["f0"].to(DEV), None, noise=inp["noise"].to(DEV),
                                        initial_phase=None if ip is None else ip.to(DEV))
    assert hid is hidden and signal.shape == (B, nF * P)
    got = {"signal": signal.cpu().numpy(), "harmonic": harm.cpu().numpy(), "noise": nz.cpu().numpy()}
    errs = {}
    for key in ("signal", "harmonic", "noise"):
        if key in gold:
            errs[key] = util.rms(got[key] - gold[key])
    report.record("sins_forward/" + name, signal_rms=util.rms(gold["signal"]), **{k + "_err": v for k, v in errs.items()},
                  signal_max=np.abs(got["signal"] - gold["signal"]).max())
    for key, e in errs.items():
        assert e < OFFICIAL_RMS, (key, e)
        assert e < GATE_RMS, (key, e)
    # the returned tensors must not alias each other (callers mutate in place, main.py:260)
    assert signal.data_ptr() != harm.data_ptr() != nz.data_ptr()


def test_sins_forward_vs_float64_truth():
    """Both the CUDA result and the reference must sit at the fp32 noise floor of the exact math."""
    name = "sins_b2_f24_h128"
    inp = G.build_inputs(name)
    truth = util.closed_form_outputs(name, inp)
    gold = util.load_golden(name)
    c = _dev_ctrls(inp)
    f0 = inp["f0"].to(DEV)
    fp, _ = ops.phase_scan(f0, P, SR)
    sig, harm, nz = ops.sins_synth(f0, fp, c["amplitudes"], c["group_delay"], c["noise_magnitude"], P, SR,
                                   noise_in=inp["noise"].to(DEV))
    e_gpu = util.rms(sig.cpu().numpy() - truth["signal"])
    e_ref = util.rms(gold["signal"] - truth["signal"])
    report.record("sins_truth", gpu_vs_truth=e_gpu, reference_vs_truth=e_ref)
    assert e_gpu < 1e-6 and e_gpu < 20 * max(e_ref, 2e-8)


def test_infer_false_rounds_phase_like_reference():
    """infer=False: the reference's fp32 cumsum (ddsp/vocoder.py:568)."""
    from oracle import torch_port as tp
    B, nF, H = 1, 30, 16
    f0 = syn.make_f0(B, nF, SR, P, seed=77)
    dense, ctrls = syn.make_ctrl(B, nF, syn.sins_split_map(H, 256, 256), seed=78)
    noise = syn.uniform_noise(B, nF * P, 79)
    with torch.no_grad():
        ref = tp.sins_forward(f0, ctrls, SR, P, noise=noise, infer=False)
    dc = syn.split_views(dense.to(DEV), syn.sins_split_map(H, 256, 256))
    model = Sins(SR, P, H, 256, 256, unit2ctrl=FixedControls(dc, None)).to(DEV)
    with torch.no_grad():
        sig, _, _ = model(None, f0.to(DEV), None, noise=noise.to(DEV), infer=False)
    e = util.rms(sig.cpu() - ref["signal"])
    report.record("sins_infer_false", err=e)
    assert e < 2e-5   # phase is quantised to fp32 ulps of up to ~1e-5 cycles here


def test_in_kernel_noise_statistics_and_determinism():
    """Throughput mode: Philox noise generated inside the FIR kernel.  Deterministic for a seed,
    independent of how the batch is sharded (utterance_offset), uniform(-1,1) statistics."""
    B, nF = 4, 64
    ir = torch.zeros(B, nF, 510, device=DEV)
    ir[:, :, 255] = 1.0                      # identity filter: output = the generated noise
    a = ops.ltv_fir(None, ir, P, seed=1234)
    b = ops.ltv_fir(None, ir, P, seed=1234)
    assert torch.equal(a, b)
    c = ops.ltv_fir(None, ir[2:], P, seed=1234, utterance_offset=2)
    assert torch.equal(a[2:], c)             # shard-invariant
    d = ops.ltv_fir(None, ir, P, seed=1235)
    assert not torch.equal(a, d)
    x = a.double().cpu().numpy()
    report.record("philox", mean=x.mean(), var=x.var(), min=x.min(), max=x.max())
    assert abs(x.mean()) < 5e-3 and abs(x.var() - 1 / 3) < 5e-3
    assert x.min() >= -1.0 and x.max() < 1.0
    assert abs(np.corrcoef(x[0, :-1], x[0, 1:])[0, 1]) < 0.02
    assert abs(np.corrcoef(x[0], x[1])[0, 1]) < 0.02


def test_full_size_properties():
    """BASELINE config-2 shape (B=32 x 10 s x 128 harmonics): too large for the per-sample oracle
    in seconds, so check size-independent properties: tiled FIR == generic FIR on the same data,
    linearity of the filter, and agreement of a sampled utterance with the oracle."""
    B, nF, H = 32, 861, 128
    sm = syn.sins_split_map(H, 256, 256)
    f0 = syn.make_f0(B, nF, SR, P)
    dense, ctrls = syn.make_ctrl(B, nF, sm)
    noise = syn.uniform_noise(B, nF * P, 11)
    dc = syn.split_views(dense.to(DEV), sm)
    f0d = f0.to(DEV)
    fp, _ = ops.phase_scan(f0d, P, SR)
    sig, harm, nz = ops.sins_synth(f0d, fp, dc["amplitudes"], dc["group_delay"], dc["noise_magnitude"], P, SR,
                                   noise_in=noise.to(DEV))
    assert torch.isfinite(sig).all()
    assert torch.equal(sig, harm + nz)
    # stage checks at full size
    sinus = ops.sins_bank(f0d, fp, dc["amplitudes"], P, SR)
    ir = ops.ir_build(dc["group_delay"], ops.IR_ALLPASS, SR)
    y_t = ops.ltv_fir(sinus, ir, P)
    y_g = ops.ltv_fir(sinus, ir, P, generic=True)
    e_tg = (y_t - y_g).abs().max().item()
    assert torch.equal(y_t, harm)
    lin = ops.ltv_fir(2.0 * sinus + noise.to(DEV), ir, P) - (2.0 * y_t + ops.ltv_fir(noise.to(DEV), ir, P))
    # one utterance against the oracle
    from oracle import torch_port as tp
    row = 17
    with torch.no_grad():
        ref = tp.sins_forward(f0[row:row + 1], {k: v[row:row + 1] for k, v in ctrls.items()}, SR, P,
                              noise=noise[row:row + 1])
    e_row = util.rms(sig[row:row + 1].cpu() - ref["signal"])
    report.record("full_size", tiled_vs_generic_max=e_tg, linearity_max=lin.abs().max().item(), row_rms=e_row,
                  signal_rms=util.rms(ref["signal"]))
    assert e_tg < 2e-6
    assert lin.abs().max().item() < 5e-5   # inputs of amplitude ~2: fp32 rounding of a 510-term sum
    assert e_row < GATE_RMS


@pytest.mark.parametrize("H", [128, 64, 31, 1])
def test_fused_bank_fir_kernel_equals_split_kernels(H):
    """`fused` evaluates the oscillator bank inside the FFT-domain FIR kernel; `split` runs the stand-alone bank kernel +
    FIR kernel (the default); `spectrum` moves the impulse-response transforms into their own kernel.  Same bank arithmetic, same transforms: the three outputs must agree to round-off (the two
    compilations may contract differently), for full and partial harmonic groups, in-kernel and explicit noise."""
    B, nF = 3, 70
    sm = syn.sins_split_map(H, 256, 256)
    f0 = syn.make_f0(B, nF, SR, P, seed=21, unvoiced_fraction=0.1, sweep_row=1).to(DEV)
    dense = syn.make_ctrl(B, nF, sm, seed=22)[0].to(DEV)
    dc = syn.split_views(dense, sm)
    noise = syn.uniform_noise(B, nF * P, 23).to(DEV)
    fp, _ = ops.phase_scan(f0, P, SR)
    outs = {}
    try:
        for impl in ("split", "fused", "spectrum"):
            ops.set_sins_impl(impl)
            outs[impl] = [ops.sins_synth(f0, fp, dc["amplitudes"], dc["group_delay"], dc["noise_magnitude"], P, SR,
                                         noise_in=noise),
                          ops.sins_synth(f0, fp, dc["amplitudes"], dc["group_delay"], dc["noise_magnitude"], P, SR,
                                         seed=5, utterance_offset=7)]
    finally:
        ops.set_sins_impl("auto")
    worst = 0.0
    for a, b in zip(outs["split"], outs["fused"]):
        for x, y in zip(a, b):
            worst = max(worst, (x - y).abs().max().item())
    report.record("sins_fused_vs_split/H%d" % H, max_diff=worst, identical=worst == 0.0)
    assert worst < 1e-7
    # the spectrum path pairs the impulse responses differently in their transforms: round-off level differences
    worst_s = 0.0
    for a, b in zip(outs["split"], outs["spectrum"]):
        for x, y in zip(a, b):
            worst_s = max(worst_s, (x - y).abs().max().item())
    report.record("sins_spectrum_vs_split/H%d" % H, max_diff=worst_s)
    assert worst_s < 2e-7
