"""Golden-vector case table shared by make_golden.py (generator, runs the LIVE reference in
the build container) and the tests (which only read the committed .npz files).

Inputs are regenerated from seeds with ddsp_svc_b200.synthetic (torch CPU generators);
each .npz also stores a float64 checksum of every input so RNG drift between torch
builds is detected instead of silently comparing different inputs.
"""
import os
from collections import OrderedDict

import torch

from ddsp_svc_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
SR, P = 44100, 512

# name -> dict(kind, B, nF, model kwargs, seeds, f0 options, which outputs are stored)
CASES = OrderedDict()


def _add(name, **kw):
    CASES[name] = kw


# --- Sins -------------------------------------------------------------------------------
_add("sins_b2_f24_h128", kind="sins", B=2, nF=24, H=128, Ma=256, Mn=256, unvoiced=0.1, sweep_row=1,
     store=("signal", "harmonic", "noise", "phase_frames"))
for _h in (1, 31, 32, 33, 64):
    _add("sins_b1_f7_h%d" % _h, kind="sins", B=1, nF=7, H=_h, Ma=256, Mn=256,
         store=("signal", "harmonic", "noise", "phase_frames"))
_add("sins_b3_f1_h128", kind="sins", B=3, nF=1, H=128, Ma=256, Mn=256,
     store=("signal", "harmonic", "noise", "phase_frames"))
_add("sins_b1_f2_h128", kind="sins", B=1, nF=2, H=128, Ma=256, Mn=256,
     store=("signal", "harmonic", "noise", "phase_frames"))
_add("sins_b1_f12_h40_m65_initphase", kind="sins", B=1, nF=12, H=40, Ma=65, Mn=129, initial_phase=True,
     store=("signal", "harmonic", "noise", "phase_frames"))
# BASELINE config 1 shape: 1 utterance, 2 s, 64 harmonics
_add("sins_cfg1_b1_f172_h64", kind="sins", B=1, nF=172, H=64, Ma=256, Mn=256, store=("signal",))
# --- CombSub (old) ----------------------------------------------------------------------
_add("combsub_b2_f24", kind="combsub", B=2, nF=24, Ma=256, Mh=512, Mn=256,
     store=("signal", "harmonic", "noise", "phase_frames"))
_add("combsub_b1_f3_unvoiced", kind="combsub", B=1, nF=3, Ma=256, Mh=512, Mn=256, unvoiced=0.4,
     store=("signal", "harmonic", "noise", "phase_frames"))
# --- CombSubSuperFast -------------------------------------------------------------------
_add("superfast_b2_f24", kind="superfast", B=2, nF=24, win=2048, store=("signal", "phase_frames"))
_add("superfast_b1_f5", kind="superfast", B=1, nF=5, win=2048, store=("signal", "phase_frames"))
_add("superfast_b1_f2_constpad", kind="superfast", B=1, nF=2, win=2048, store=("signal", "phase_frames"))
# --- CombSubFast (1024-point sqrt-Hann frames; the variant the diffusion / reflow vocoders embed) --------
_add("csfast_b2_f24", kind="combsubfast", B=2, nF=24, unvoiced=0.1, sweep_row=1, store=("signal", "phase_frames"))
_add("csfast_b1_f3_unvoiced", kind="combsubfast", B=1, nF=3, unvoiced=0.4, store=("signal", "phase_frames"))
_add("csfast_b1_f1", kind="combsubfast", B=1, nF=1, store=("signal", "phase_frames"))
_add("csfast_b1_f9_initphase", kind="combsubfast", B=1, nF=9, initial_phase=True, store=("signal", "phase_frames"))
# --- SineGen ----------------------------------------------------------------------------
_add("sinegen_b2_f12", kind="sinegen", B=2, nF=12, upp=512, harmonic_num=8, unvoiced=0.25, store=("out",))
_add("sinegen_b1_f3_upp256", kind="sinegen", B=1, nF=3, upp=256, harmonic_num=8, store=("out",))
# SourceModuleHnNSF = SineGen + tanh(Linear(9 -> 1)); the golden also stores the seeded Linear parameters
_add("srcmod_b2_f10", kind="source_module", B=2, nF=10, upp=512, harmonic_num=8, unvoiced=0.3,
     store=("out",), extra=("weight", "bias"))


def path(name):
    return os.path.join(HERE, name + ".npz")


def split_map(case):
    k = case["kind"]
    if k == "sins":
        return syn.sins_split_map(case["H"], case["Ma"], case["Mn"])
    if k == "combsub":
        return syn.combsub_split_map(case["Ma"], case["Mh"], case["Mn"])
    if k == "superfast":
        return syn.superfast_split_map(case["win"])
    if k == "combsubfast":
        return syn.combsubfast_split_map(P)
    return None


def seeds(name):
    base = sum(ord(c) for c in name)
    return {"f0": 1000 + base, "ctrl": 2000 + base, "noise": 3000 + base}


def build_inputs(name):
    """Regenerate the inputs of a case: dict with f0, dense ctrl, ctrl views, noise, ..."""
    case = CASES[name]
    sd = seeds(name)
    B, nF = case["B"], case["nF"]
    out = {"case": case}
    if case["kind"] in ("sinegen", "source_module"):
        upp = case["upp"]
        out["f0"] = syn.make_f0(B, nF, SR, upp, seed=sd["f0"],
                                unvoiced_fraction=case.get("unvoiced", 0.0))[..., 0].contiguous()
        dim = case["harmonic_num"] + 1
        torch.manual_seed(sd["noise"])
        rand_ini = torch.rand(1, 1, dim)
        rand_ini[..., 0] = 0
        out["rand_ini"] = rand_ini
        out["noise"] = torch.randn(B, nF * upp, dim)
        return out
    out["f0"] = syn.make_f0(B, nF, SR, P, seed=sd["f0"], unvoiced_fraction=case.get("unvoiced", 0.0),
                            sweep_row=case.get("sweep_row"))
    dense, views = syn.make_ctrl(B, nF, split_map(case), seed=sd["ctrl"])
    out["dense"], out["ctrls"] = dense, views
    if case["kind"] == "superfast":
        out["noise"] = syn.normal_noise((B, nF * P), sd["noise"])
    else:
        out["noise"] = syn.uniform_noise(B, nF * P, sd["noise"])
    if case.get("initial_phase"):
        g = torch.Generator().manual_seed(sd["noise"] + 1)
        out["initial_phase"] = (torch.rand(B, 1, 1, generator=g) * 6.0 - 3.0)
    return out


def input_checksums(inp):
    cs = {}
    for k in ("f0", "dense", "noise", "rand_ini", "initial_phase"):
        if k in inp and inp[k] is not None:
            t = inp[k].double()
            cs["cs_" + k] = float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64)
                                   .reshape(t.shape).remainder(97.0)).sum())
    return cs
