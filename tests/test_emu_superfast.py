"""csrc/superfast.cu's main KERNEL SOURCE executed on the CPU (tests/emu/host_emu.h) against the live-reference
goldens.  The kernel is validated on hardware (tests/test_gpu_superfast.py); the emulation exists so that changes to
the shared FFT code can be checked -- and race-checked under ThreadSanitizer -- without a GPU.  The frame scan (warp
shuffles) is restated here in numpy with the kernel's fp32 operation order."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import util
from tests.golden import cases as G

HERE = os.path.dirname(os.path.abspath(__file__))
SR, P, WIN = G.SR, G.P, 2048
f32 = np.float32

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")


def frame_par(f0):
    """(s, ds, acc_prev, 0) per frame as superfast_scan_kernel computes them (fp32 steps, fp64 running sum)"""
    f0 = np.asarray(f0, f32)
    B, nF = f0.shape
    s = (f0 / f32(SR)).astype(f32)
    ds = np.zeros_like(s)
    ds[:, :-1] = s[:, 1:] - s[:, :-1]
    fP, fPm1 = f32(P), f32(P - 1)
    t2 = (((f32(0.5) * ds).astype(f32) * fPm1).astype(f32) * fP).astype(f32)
    last = ((s * fP).astype(f32) + (t2 / fP).astype(f32)).astype(f32)
    adv = (np.fmod((last + f32(0.5)).astype(f32), f32(1.0)) - f32(0.5)).astype(f32)
    run = np.concatenate([np.zeros((B, 1)), np.cumsum(adv.astype(np.float64), axis=1)[:, :-1]], axis=1)
    accp = np.fmod(run.astype(f32), f32(1.0)).astype(f32)
    accp[:, 0] = 0
    return np.ascontiguousarray(np.stack([s, ds, accp, np.zeros_like(s)], axis=-1), f32)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libemu_sf.so")
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
           "-o", so, os.path.join(HERE, "emu", "emu_superfast.cpp")]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    lib = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.emu_superfast.argtypes = [fp, fp, fp, fp, fp, ctypes.c_longlong, fp, ctypes.c_ulonglong, ctypes.c_longlong,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]

    def run(f0, dense, noise, hops=29, seed=0, utt_off=0):
        B, nF = f0.shape[0], f0.shape[1]
        par = frame_par(np.asarray(f0).reshape(B, nF))
        dense = np.ascontiguousarray(dense, f32)
        nz = None if noise is None else np.ascontiguousarray(noise, f32)
        out = np.full((B, nF * P), np.nan, f32)
        n = WIN // 2 + 1
        ptr = lambda a, off=0: ctypes.cast(a.ctypes.data + 4 * off, fp)
        rc = lib.emu_superfast(ptr(par), ptr(dense, 0), ptr(dense, n), ptr(dense, 2 * n), ptr(dense, 3 * n),
                               dense.shape[2], ptr(nz) if nz is not None else None, seed, utt_off, B, nF, hops, ptr(out))
        assert rc == 0
        return out

    return run


@pytest.mark.parametrize("name", [n for n, c in G.CASES.items() if c["kind"] == "superfast"])
@pytest.mark.parametrize("hops", [29, 5])
def test_kernel_source_matches_reference_golden(emu, name, hops):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    got = emu(inp["f0"].numpy(), inp["dense"].numpy(), inp["noise"].numpy(), hops=hops)
    assert not np.isnan(got).any()
    e, m = util.rms(got - gold["signal"]), np.abs(got - gold["signal"]).max()
    assert e < 2e-7 and m < 5e-6, (name, hops, e, m)


def test_in_kernel_noise_is_shard_invariant(emu):
    inp = G.build_inputs("superfast_b2_f24")
    f0, dense = inp["f0"].numpy(), inp["dense"].numpy()
    full = emu(f0, dense, None, seed=3)
    part = emu(f0[1:], dense[1:], None, seed=3, utt_off=1)
    assert np.array_equal(full[1:], part) and np.isfinite(full).all()
