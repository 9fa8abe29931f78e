"""The oracle against the committed golden vectors (made by the LIVE reference,
tests/golden/make_golden.py).  CPU only.

* torch_port must reproduce the reference outputs essentially bit for bit (same ATen
  operators; a tiny tolerance only absorbs ISA-dependent vectorisation on other hosts);
* closed_form (independent float64 restatement) must agree to the reference's own fp32
  noise floor, which is what makes it usable as ground truth at sizes with no golden.
"""
import numpy as np
import pytest

from tests.golden import cases as G
from tests import util

# per-kind (rms, max) bound of |closed_form - reference|: the reference's own fp32 error
CF_BOUNDS = {"sins": (2e-7, 2e-6), "combsub": (2e-7, 2e-6), "superfast": (5e-6, 2e-4),
             "sinegen": (5e-6, 5e-5), "source_module": (5e-6, 5e-5), "combsubfast": (5e-7, 5e-6)}


@pytest.mark.parametrize("name", list(G.CASES))
def test_port_matches_golden(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    util.check_inputs_match_golden(name, inp, gold)
    out = util.port_outputs(name, inp)
    for key in inp["case"]["store"]:
        got = out[key].numpy()
        ref = gold[key]
        assert got.shape == ref.shape, key
        assert np.abs(got - ref).max() <= 2e-7, (name, key, np.abs(got - ref).max())


@pytest.mark.parametrize("name", list(G.CASES))
def test_closed_form_matches_golden(name):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    out = util.closed_form_outputs(name, inp)
    r_bound, m_bound = CF_BOUNDS[inp["case"]["kind"]]
    for key in inp["case"]["store"]:
        if key == "phase_frames":
            continue
        diff = out[key] - gold[key]
        assert util.rms(diff) <= r_bound, (name, key, util.rms(diff))
        assert np.abs(diff).max() <= m_bound, (name, key, np.abs(diff).max())
