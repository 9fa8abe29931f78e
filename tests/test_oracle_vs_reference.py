"""Pin the oracle on the LIVE reference (only possible in the build container, where
/root/reference exists; skipped on the GPU box).  The port must be bit-identical."""
import contextlib
import io

import pytest
import torch

from oracle import ref_loader
from tests.golden import cases as G
from tests.golden import make_golden
from tests import util

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="live reference not present")


@pytest.mark.parametrize("name", ["sins_b2_f24_h128", "sins_b1_f7_h33", "sins_b1_f12_h40_m65_initphase",
                                  "combsub_b2_f24", "superfast_b2_f24", "superfast_b1_f2_constpad",
                                  "sinegen_b2_f12", "csfast_b2_f24", "csfast_b1_f9_initphase", "srcmod_b2_f10"])
def test_port_bit_identical_to_live_reference(name):
    with contextlib.redirect_stdout(io.StringIO()):
        inp, ref = make_golden.run_reference(name)
    out = util.port_outputs(name, inp)
    for key in inp["case"]["store"]:
        assert torch.equal(out[key], ref[key]), (name, key, (out[key] - ref[key]).abs().max().item())
