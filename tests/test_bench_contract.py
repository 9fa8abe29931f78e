"""bench.py's JSON contract on the paths that run without a GPU: the reference (CPU) arm prints one
line with the required keys; the workload table and the algorithmic-byte model match SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_match_the_survey():
    nF = 861
    T = nF * 512
    per = lambda w: bench.algorithmic_bytes(bench.WORKLOADS[w], nF) / (bench.WORKLOADS[w]["B"] * T)
    assert abs(per("sins") - 17.008) < 2e-3            # 3 outputs + 641 control words per frame
    assert abs(per("superfast") - 36.039) < 2e-3
    assert abs(per("combsub") - 20.008) < 2e-3
    assert abs(per("sinegen") - 36.008) < 2e-3
    assert abs(per("srcmod") - 4.008) < 2e-3           # fused tanh(Linear(9 -> 1)): one value per sample
    assert abs(per("combsubfast") - 16.031) < 2e-3     # 3 x 513 control words per frame + 1 output
    assert abs(per("mel") - 5.0) < 1e-6                # waveform in, 128 mels per 512-sample hop out
    assert abs(per("maskmul") - 8.008) < 2e-3          # read + write in place, one mask value per frame


def test_cpu_arm_runs_every_workload_kind():
    """the oracle port behind --impl reference / cpu_baseline accepts every workload (tiny sizes)"""
    import torch
    from ddsp_svc_b200 import synthetic as syn
    for name, w in bench.WORKLOADS.items():
        w = dict(w, B=1, sec=0.05)
        nF = max(2, syn.n_frames_for(w["sec"], bench.SR, bench.P))
        if w["kind"] == "mel":
            nF = 8                                       # one 2048-sample frame needs more than 0.05 s
        f0, ctrls = bench.oracle_inputs(w, 1, nF)
        with torch.no_grad():
            out = bench.oracle_forward(w, f0, ctrls)
        key = "signal" if "signal" in out else "out"
        assert out[key].shape[2 if w["kind"] == "mel" else 1] == (nF if w["kind"] == "mel" else nF * bench.P), name


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sins_cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
