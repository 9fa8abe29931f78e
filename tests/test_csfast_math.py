"""CPU emulation (numpy) of the algorithm inside combsubfast.cu, checked against the closed-form oracle and the
live-reference goldens: 1024-point Stockham passes (16, 8, 8) with the kernel's index formulas; one complex FFT per
frame carrying window*(comb + j*noise), split by conjugate symmetry; the C2R convention for the DC / Nyquist bins
(imaginary parts ignored); ONE inverse FFT per PAIR of frames (Sa + j Sb) by the swap trick; overlap-add with a
carried half-frame tail; chunks of hops recomputing one boundary frame.  Pins the derivation; the GPU tests then
only have to catch implementation slips."""
import numpy as np
import pytest

from oracle import closed_form as cf
from tests.golden import cases as G
from tests import util

P = 512
N = 2 * P


def stockham_fft(x, radices=(16, 8, 8)):
    """butterfly j of a radix-R pass reads in[j + r N/R], multiplies by exp(-2 pi i r (j % Ns) / (Ns R)), does an
    R-point DFT and writes out[(j // Ns) Ns R + (j % Ns) + r Ns]   (same formulas as superfast.cu)."""
    n = len(x)
    data = np.asarray(x, np.complex128).copy()
    Ns = 1
    for R in radices:
        j = np.arange(n // R)
        k = j % Ns
        r = np.arange(R)[:, None]
        v = data[j[None, :] + r * (n // R)] * np.exp(-2j * np.pi * r * k[None, :] / (Ns * R))
        v = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R) @ v
        out = np.empty_like(data)
        out[((j // Ns) * Ns * R + k)[None, :] + r * Ns] = v
        data = out
        Ns *= R
    return data


def test_stockham_1024_pass_structure_is_a_dft():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.abs(stockham_fft(x) - np.fft.fft(x)).max() < 1e-9
    sw = lambda z: z.imag + 1j * z.real
    assert np.abs(sw(stockham_fft(sw(x))) / N - np.fft.ifft(x)).max() < 1e-12


def emulate_combsubfast(comb, noise, hm, hp, nm, chunk=5):
    """comb/noise [T]; hm/hp/nm [nF, P+1] raw controls -> signal [T], processed like the kernel: chunks of `chunk`
    hops, frames q = h0..h1 in pairs, tail carried between frames."""
    T = len(comb)
    nF = T // P
    w = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N))
    out = np.full(T, np.nan)

    def spectrum(q):
        """filtered half spectrum S[0..P] of frame q (original samples [(q-1)P, (q+1)P), zeros outside)"""
        m = (q - 1) * P + np.arange(N)
        ok = (m >= 0) & (m < T)
        z = np.zeros(N, np.complex128)
        z[ok] = w[ok] * (comb[m[ok]] + 1j * noise[m[ok]])
        Z = stockham_fft(z)
        k = np.arange(P + 1)
        Zk, Zm = Z[k], np.conj(Z[(N - k) % N])
        C, Nz = 0.5 * (Zk + Zm), -0.5j * (Zk - Zm)
        row = min(q, nF - 1)                                    # frame nF reuses row nF-1 (:759,761)
        S = C * np.exp(hm[row]) * np.exp(1j * np.pi * hp[row]) + Nz * (np.exp(nm[row]) / 128.0)
        S[0] = S[0].real                                        # C2R ignores Im of DC and Nyquist
        S[P] = S[P].real
        return S

    for h0 in range(0, nF, chunk):
        h1 = min(h0 + chunk, nF)
        tail = None
        q = h0
        while q <= h1:
            Sa = spectrum(q)
            Sb = spectrum(q + 1) if q + 1 <= h1 else np.zeros(P + 1, np.complex128)
            Y = np.zeros(N, np.complex128)
            k = np.arange(1, P)
            Y[: P + 1] = Sa + 1j * Sb
            Y[N - k] = np.conj(Sa[k]) + 1j * np.conj(Sb[k])
            sw = lambda z: z.imag + 1j * z.real
            y = sw(stockham_fft(sw(Y))) / N                     # real = frame a, imag = frame b
            ya, yb = y.real * w, y.imag * w
            if tail is not None:                                # hop q-1 = tail of frame q-1 + head of frame q
                out[(q - 1) * P:q * P] = tail + ya[:P]
            if q + 1 <= h1:
                out[q * P:(q + 1) * P] = ya[P:] + yb[:P]       # hop q
                tail = yb[P:]
            q += 2
    assert not np.isnan(out).any()
    return out


@pytest.mark.parametrize("name", ["csfast_b2_f24", "csfast_b1_f3_unvoiced", "csfast_b1_f1"])
@pytest.mark.parametrize("chunk", [1, 4, 5, 32])
def test_emulation_matches_closed_form_and_golden(name, chunk):
    inp = G.build_inputs(name)
    gold = util.load_golden(name)
    truth = util.closed_form_outputs(name, inp)
    c = {k: v.numpy().astype(np.float64) for k, v in inp["ctrls"].items()}
    for b in range(inp["case"]["B"]):
        got = emulate_combsubfast(truth["comb"][b], inp["noise"][b].numpy().astype(np.float64),
                                  c["harmonic_magnitude"][b], c["harmonic_phase"][b], c["noise_magnitude"][b], chunk)
        assert np.abs(got - truth["signal"][b]).max() < 1e-12
        assert util.rms(got - gold["signal"][b]) < 5e-7
