"""Drop-in contract: constructor arguments, buffers / state-dict keys and class dispatch match
the reference.  The parts that need the live reference (Unit2Control,
strict state-dict round trip, patch_reference) run only where /root/reference exists."""
import contextlib
import io
import sys

import pytest
import torch

import ddsp_svc_b200 as pkg
from ddsp_svc_b200 import dropin
from oracle import ref_loader


def test_buffers_and_forward_signature_without_reference():
    import inspect
    m = pkg.Sins(44100, 512, 128, 256, 256, unit2ctrl=pkg.FixedControls())
    assert set(m.state_dict()) == {"sampling_rate", "block_size"}
    assert m.sampling_rate.dtype == torch.int64 and m.sampling_rate.dim() == 0
    sig = inspect.signature(m.forward)
    for name in ("units_frames", "f0_frames", "volume_frames", "spk_id", "spk_mix_dict", "initial_phase", "infer",
                 "max_upsample_dim"):
        assert name in sig.parameters
    c = pkg.CombSub(44100, 512, 256, 512, 256, unit2ctrl=pkg.FixedControls())
    assert set(c.state_dict()) == {"sampling_rate", "block_size"}
    cf = pkg.CombSubFast(44100, 512, unit2ctrl=pkg.FixedControls())
    assert list(cf.state_dict()) == ["sampling_rate", "block_size", "window"] and cf.window.shape == (1024,)
    g = pkg.SineGen(44100, harmonic_num=8)
    assert g.dim == 9 and len(g.state_dict()) == 0
    sm = pkg.SourceModuleHnNSF(44100, harmonic_num=8)
    assert list(sm.state_dict()) == ["l_linear.weight", "l_linear.bias"]
    assert sm.l_linear.weight.shape == (1, 9) and sm.l_sin_gen.dim == 9
    assert "voiced_threshod" in inspect.signature(pkg.SourceModuleHnNSF.__init__).parameters   # reference spelling
    with pytest.raises(ValueError, match="Unknown Model"):
        dropin.build_model(dropin.DotDict({"model": {"type": "Nope"}, "data": {}}))
    # no unit2ctrl argument: the package's own Unit2Control (no dependency on the reference package), same head as the
    # reference's (dense_out -> 640 outputs for Sins 128/256/256)
    own = pkg.Sins(44100, 512, 128, 256, 256, n_unit=768)
    from ddsp_svc_b200.unit2control import Unit2Control
    assert isinstance(own.unit2ctrl, Unit2Control) and own.unit2ctrl.dense_out.weight_v.shape == (640, 256)
    assert {"unit2ctrl.dense_out.weight_g", "unit2ctrl.dense_out.weight_v", "unit2ctrl.dense_out.bias"} <= set(own.state_dict())
    with pytest.raises(ValueError, match="block_size"):
        pkg.Sins(44100, 441, 128, 256, 256, unit2ctrl=pkg.FixedControls())          # unsupported shapes fail at build time


def test_module_refuses_cpu_tensors():
    m = pkg.Sins(44100, 512, 4, 9, 9, unit2ctrl=pkg.FixedControls())
    with pytest.raises(ValueError, match="CUDA"):
        m(None, torch.zeros(1, 3, 1), None)


@pytest.mark.skipif(not ref_loader.available(), reason="live reference not present")
def test_state_dict_round_trip_with_the_reference_classes(tmp_path):
    with contextlib.redirect_stdout(io.StringIO()):
        V, _, RefSineGen = ref_loader.load()
        cases = [
            (V.Sins(44100, 512, 128, 256, 256, n_unit=768, n_spk=1), lambda: pkg.Sins(44100, 512, 128, 256, 256, n_unit=768, n_spk=1)),
            (V.CombSub(44100, 512, 256, 512, 256, n_unit=768, n_spk=2), lambda: pkg.CombSub(44100, 512, 256, 512, 256, n_unit=768, n_spk=2)),
            (V.CombSubSuperFast(44100, 512, 2048, n_unit=768, n_spk=1), lambda: pkg.CombSubSuperFast(44100, 512, 2048, n_unit=768, n_spk=1)),
            (V.CombSubFast(44100, 512, n_unit=768, n_spk=1), lambda: pkg.CombSubFast(44100, 512, n_unit=768, n_spk=1)),
        ]
        for ref_model, make in cases:
            ours = make()                                   # uses the reference's Unit2Control
            sd = ref_model.state_dict()
            assert list(ours.state_dict().keys()) == list(sd.keys())
            ours.load_state_dict(sd)                        # strict, like ddsp/vocoder.py:527
            for k, v in ours.state_dict().items():
                assert torch.equal(v, sd[k]), k
        # load_model: config.yaml next to the checkpoint -> class dispatch -> strict load
        import yaml
        cfg = {"data": {"sampling_rate": 44100, "block_size": 512, "encoder_out_channels": 768},
               "model": {"type": "CombSubSuperFast", "win_length": 2048, "n_spk": 1}}
        (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
        torch.save({"global_step": 1, "model": cases[2][0].state_dict()}, tmp_path / "model_1.pt")
        model, args = dropin.load_model(str(tmp_path / "model_1.pt"), device="cpu")
        assert isinstance(model, pkg.CombSubSuperFast) and args.model.win_length == 2048
        # patch / unpatch
        saved = pkg.patch_reference()
        try:
            assert V.Sins is pkg.Sins and V.CombSubSuperFast is pkg.CombSubSuperFast
            assert V.CombSubFast is pkg.CombSubFast
            import nsf_hifigan.models as nsf
            assert nsf.SineGen is pkg.SineGen and nsf.SourceModuleHnNSF is pkg.SourceModuleHnNSF
            model2, _ = V.load_model(str(tmp_path / "model_1.pt"), device="cpu")   # the REFERENCE's loader
            assert isinstance(model2, pkg.CombSubSuperFast)
        finally:
            pkg.unpatch_reference(saved)
        assert V.Sins is not pkg.Sins and nsf.SourceModuleHnNSF is not pkg.SourceModuleHnNSF
        ref_sm = nsf.SourceModuleHnNSF(44100, harmonic_num=8)          # state dict of the reference class loads strictly
        ours_sm = pkg.SourceModuleHnNSF(44100, harmonic_num=8)
        assert list(ours_sm.state_dict()) == list(ref_sm.state_dict())
        ours_sm.load_state_dict(ref_sm.state_dict())
