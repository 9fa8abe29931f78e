"""CPU emulations (numpy, float64) of the index logic the CUDA kernels use, checked against the
closed-form oracle.  These do not run device code; they pin the ALGEBRA of the restructured
algorithms (so a GPU parity failure can only be an implementation slip, not a derivation error):

* ltv_fir.cu   : tile f / segment s / "A"-"B" table switch + band fix-up + (i-1)/P recombination
* ir_build.cu  : even/odd-bin split of the inverse real DFT, t <-> M-1-t pairing, causal roll
* sins_bank.cu : harmonic h = a + 16 b factorisation with per-base rotation
"""
import numpy as np
import pytest

from oracle import closed_form as cf


def emulate_fir_tiles(x, ir, P):
    """Mirror of ltv_fir_kernel for one utterance: x [T], ir [nF, L] -> y [T]."""
    nF, L = ir.shape
    T = nF * P
    Mh = L // 2 + 1
    NS = (L + P - 1) // P
    ntiles = nF + (Mh + P - 1) // P
    y = np.zeros(T)
    clamp = lambda j: min(max(j, 0), nF - 1)
    for f in range(ntiles):
        a1 = np.zeros(P)
        a2 = np.zeros(P)
        for s in range(NS):
            g = f - s
            if g < 0 or g > nF:
                continue
            m = g * P - P + np.arange(2 * P)
            xs = np.where((m >= 0) & (m < T), x[np.clip(m, 0, T - 1)], 0.0)
            tau = s * P + np.arange(P)
            take = lambda row: np.where(tau < L, ir[row][np.minimum(tau, L - 1)], 0.0)
            vm, v0, vp = take(clamp(g - 1)), take(clamp(g)), take(clamp(g + 1))
            w = np.arange(P) / P
            eA, eB = vp - v0, v0 - vm
            GA, GB = v0 - w * eA, v0 - w * eB
            for lt in range(P // 8):
                i0 = 8 * lt
                for step in range(P // 4):
                    G, E = (GA, eA) if step < 2 * lt else (GB, eB)
                    for tt in range(4):
                        tp = 4 * step + tt
                        for r in range(8):
                            xv = xs[P - 1 + i0 + r - tp]
                            a1[i0 + r] += xv * G[tp]
                            a2[i0 + r] += xv * E[tp]
                for bb in range(7):
                    tp = i0 + bb
                    dG, dE = GA[tp] - GB[tp], eA[tp] - eB[tp]
                    for r in range(bb + 1, 8):
                        xv = xs[P - 1 + r - bb]
                        a1[i0 + r] += xv * dG
                        a2[i0 + r] += xv * dE
        i = np.arange(P)
        yv = a1 + (i - 1) / P * a2
        n = f * P - Mh + i
        ok = (n >= 0) & (n < T)
        y[n[ok]] = yv[ok]
    return y


@pytest.mark.parametrize("P,M,nF", [(32, 17, 5), (32, 33, 4), (16, 9, 1), (32, 16, 3)])
def test_fir_tiling_matches_definition(P, M, nF):
    rng = np.random.default_rng(P + M + nF)
    L = 2 * (M - 1)
    x = rng.standard_normal(nF * P)
    ir = rng.standard_normal((nF, L))
    want = cf.ltv_fir(x[None], ir[None], P)[0]
    got = emulate_fir_tiles(x, ir, P)
    assert np.abs(got - want).max() < 1e-10


def emulate_ir_even_odd(R, I, M):
    """Mirror of ir_build_kernel's DFT stage for one frame: spectrum (R + jI)[M] -> h[L]."""
    L = 2 * (M - 1)
    Nt = (M - 1) // 2 + 1
    wgt = np.where((np.arange(M) == 0) | (np.arange(M) == M - 1), 1.0, 2.0) / L
    ar, ai = R * wgt, I * wgt
    h = np.full(L, np.nan)
    for t in range(Nt):
        ang = lambda m: 2 * np.pi * ((m * t) % L) / L
        me, mo = np.arange(0, M, 2), np.arange(1, M, 2)
        Ce, Co = (ar[me] * np.cos(ang(me))).sum(), (ar[mo] * np.cos(ang(mo))).sum()
        Se, So = (ai[me] * np.sin(ang(me))).sum(), (ai[mo] * np.sin(ang(mo))).sum()
        tl, th = t, M - 1 - t
        Cl, Ch, Sl, Sh = Ce + Co, Ce - Co, Se + So, So - Se
        if tl <= M - 2: h[M - 1 + tl] = Cl - Sl
        if tl >= 1: h[M - 1 - tl] = Cl + Sl
        if th != tl and th <= M - 2: h[M - 1 + th] = Ch - Sh
        if th != tl and th >= 1: h[M - 1 - th] = Ch + Sh
    return h


@pytest.mark.parametrize("M", [256, 65, 64, 9, 2, 3])
def test_ir_even_odd_split_matches_irfft(M):
    rng = np.random.default_rng(M)
    spec = rng.standard_normal(M) + 1j * rng.standard_normal(M)
    want = cf.impulse_response(spec[None, None], "none")[0, 0]
    got = emulate_ir_even_odd(spec.real, spec.imag, M)
    assert not np.isnan(got).any()
    assert np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("H", [128, 64, 33, 1, 200])
def test_bank_anchor_base_factorisation(H):
    rng = np.random.default_rng(H)
    x = rng.uniform(-0.5, 0.5, 50)
    amp = rng.uniform(0, 1, (50, H))
    want = sum(np.sin(2 * np.pi * h * x) * amp[:, h - 1] for h in range(1, H + 1))
    got = np.zeros(50)
    G = (H + 127) // 128
    for g in range(G):
        for b in range(8):
            hb = 128 * g + 16 * b
            Pb = np.zeros(50)
            Qb = np.zeros(50)
            for a in range(1, 17):
                h = hb + a
                if h > H:
                    continue
                Pb += np.sin(2 * np.pi * a * x) * amp[:, h - 1]
                Qb += np.cos(2 * np.pi * a * x) * amp[:, h - 1]
            r = hb * x - np.rint(hb * x)
            got += Pb * np.cos(2 * np.pi * r) + Qb * np.sin(2 * np.pi * r)
    assert np.abs(got - want).max() < 1e-10
