"""HostPipeline (chunked upload | kernels | download on three streams) must give exactly the result
of the direct call, for ragged chunkings too."""
import os

import pytest
import torch

from ddsp_svc_b200 import FixedControls, HostPipeline, Sins, synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,chunks", [(8, 3), (5, 4), (2, 8), (6, 1), (8, (1, 3, 3, 1)), (5, (4, 9, 13, 6))])
def test_host_pipeline_matches_direct_call(B, chunks):
    nF, H = 30, 64
    sm = syn.sins_split_map(H, 256, 256)
    f0 = syn.make_f0(B, nF).pin_memory()
    dense = syn.make_ctrl(B, nF, sm)[0].pin_memory()
    noise = syn.uniform_noise(B, nF * 512, 3).to(DEV)
    fixed = FixedControls()
    model = Sins(44100, 512, H, 256, 256, unit2ctrl=fixed).to(DEV)
    out = torch.empty(B, nF * 512).pin_memory()

    def fwd(d, lo, hi):
        fixed.ctrls = syn.split_views(d["dense"], sm)
        return model(None, d["f0"], None, noise=noise[lo:hi])[0]

    with torch.no_grad():
        pipe = HostPipeline(DEV, chunks=chunks)
        pipe.run({"f0": f0, "dense": dense}, fwd, out).synchronize()
        first = out.clone()
        pipe.run({"f0": f0, "dense": dense}, fwd, out).synchronize()      # buffers are reused across runs
        fixed.ctrls = syn.split_views(dense.to(DEV), sm)
        ref = model(None, f0.to(DEV), None, noise=noise)[0].cpu()
    assert torch.equal(first, ref) and torch.equal(out, ref)


@pytest.mark.skipif(os.environ.get("B2D_EXPERIMENTAL") != "1",
                    reason="compute_streams > 1 has not run on hardware yet (set B2D_EXPERIMENTAL=1)")
def test_host_pipeline_two_compute_streams_matches_direct_call():
    """opt-in mode: chunks alternate between two compute streams"""
    B, nF, H = 6, 30, 64
    sm = syn.sins_split_map(H, 256, 256)
    f0 = syn.make_f0(B, nF).pin_memory()
    dense = syn.make_ctrl(B, nF, sm)[0].pin_memory()
    noise = syn.uniform_noise(B, nF * 512, 3).to(DEV)
    fixed = FixedControls()
    model = Sins(44100, 512, H, 256, 256, unit2ctrl=fixed).to(DEV)
    out = torch.empty(B, nF * 512).pin_memory()

    def fwd(d, lo, hi):
        ctrls = syn.split_views(d["dense"], sm)
        return Sins(44100, 512, H, 256, 256, unit2ctrl=FixedControls(ctrls, None))(None, d["f0"], None, noise=noise[lo:hi])[0]

    with torch.no_grad():
        pipe = HostPipeline(DEV, chunks=3, compute_streams=2)
        for _ in range(2):
            pipe.run({"f0": f0, "dense": dense}, fwd, out).synchronize()
        torch.cuda.synchronize()
        fixed.ctrls = syn.split_views(dense.to(DEV), sm)
        ref = model(None, f0.to(DEV), None, noise=noise)[0].cpu()
    assert torch.equal(out, ref)


def test_host_pipeline_requires_pinned_memory():
    pipe = HostPipeline(DEV, chunks=2)
    with pytest.raises(ValueError, match="pinned"):
        pipe.run({"f0": torch.zeros(2, 4, 1)}, lambda d, lo, hi: d["f0"], torch.zeros(2, 4, 1).pin_memory())
