#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unit2control.py tests/test_gpu_acceptance.py -q > gpurun_out/pytest_u2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_u2c.log; tail -30 gpurun_out/pytest_u2c.log | cut -c1-220
python - <<'PY'
# Unit2Control: this package's module (three GEMM precisions) vs the reference's class, all on the GPU, B = 32 x 861 frames
import contextlib, io, copy, torch
from oracle import ref_loader
from ddsp_svc_b200.unit2control import Unit2Control
if ref_loader.available():
    with contextlib.redirect_stdout(io.StringIO()):
        ref_loader.load()
    from ddsp.unit2control import Unit2Control as Ref
    dev = "cuda:0"
    for name, kw, splits in (("pcmer/sins", {}, {"a": 128, "b": 256, "c": 256}),
                             ("naive/superfast", dict(use_naive_v2=True, use_conv_stack=True), {"a": 1025, "b": 1025, "c": 1025, "d": 1025})):
        torch.manual_seed(0)
        ref = Ref(768, 1, splits, **kw).eval()
        B, T = 32, 861
        u = torch.randn(B, T, 768); f0 = 200 + 100 * torch.rand(B, T, 1); ph = torch.rand(B, T, 1); vo = torch.rand(B, T, 1)
        with torch.no_grad():
            dense = torch.cat(list(ref(u, f0, ph, vo)[0].values()), -1)
        ref = ref.to(dev)
        ours = Unit2Control(768, 1, splits, **kw).to(dev).eval(); ours.load_state_dict(ref.state_dict())
        u, f0, ph, vo = (t.to(dev) for t in (u, f0, ph, vo))
        models = [("reference_eager(torch defaults)", ref)]
        for mode in ("3xtf32", "fp32", "tf32"):
            m = copy.copy(ours); m.gemm_precision = mode; m.__dict__["_packed"] = None
            models.append(("b200_" + mode, m))
        if not kw:
            m = copy.copy(ours); m.fused_attention = False; m.__dict__["_packed"] = None
            models.append(("b200_3xtf32_unfused_attention", m))
        for tag, m in models:
            with torch.no_grad():
                got = torch.cat(list(m(u, f0, ph, vo)[0].values()), -1).cpu()
                err = float((got - dense).pow(2).mean().sqrt() / dense.pow(2).mean().sqrt())
                for _ in range(3): m(u, f0, ph, vo)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5): m(u, f0, ph, vo)
                b.record(); b.synchronize()
            print("unit2control %-16s %-40s %7.3f ms   controls rel rms err vs reference CPU fp32 %.2e" % (name, tag, a.elapsed_time(b) / 5, err))
PY
