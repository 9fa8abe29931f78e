"""CPU emulation (numpy) of the algorithm inside superfast.cu, checked against the closed-form
oracle: Stockham mixed-radix passes (16,16,8) with the kernel's index formulas, one complex FFT
carrying comb + j*noise per frame, split by conjugate symmetry, a single inverse FFT for a PAIR
of frames (Ya + j Yb), reflect padding, overlap-add in a ring and division by the window
envelope.  Pins the derivation; the GPU tests then only have to catch implementation slips."""
import numpy as np
import pytest

from oracle import closed_form as cf

N = 2048


def stockham_fft(x, radices=(16, 16, 8)):
    """Forward DFT by the pass structure of superfast.cu: butterfly j of a radix-R pass reads
    in[j + r*N/R], multiplies by exp(-2 pi i r (j % Ns) / (Ns R)), does an R-point DFT and writes
    out[(j // Ns) * Ns * R + (j % Ns) + r * Ns]."""
    n = len(x)
    data = np.asarray(x, np.complex128).copy()
    Ns = 1
    for R in radices:
        out = np.empty_like(data)
        Wr = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        for j in range(n // R):
            k = j % Ns
            v = np.array([data[j + r * (n // R)] * np.exp(-2j * np.pi * r * k / (Ns * R)) for r in range(R)])
            v = Wr @ v
            base = (j // Ns) * Ns * R + k
            for r in range(R):
                out[base + r * Ns] = v[r]
        data = out
        Ns *= R
    return data


def test_stockham_pass_structure_is_a_dft():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.abs(stockham_fft(x) - np.fft.fft(x)).max() < 1e-9
    # inverse by the swap trick used in the kernel: ifft(x) = swap(fft(swap(x))) / N
    sw = lambda z: z.imag + 1j * z.real
    assert np.abs(sw(stockham_fft(sw(x))) / N - np.fft.ifft(x)).max() < 1e-12


def emulate_superfast(comb, noise, h_src, h_noise, P, chunk=5):
    """comb/noise [T]; h_src/h_noise [nF, N/2+1] complex (frame nF reuses nF-1) -> signal [T]."""
    T = len(comb)
    nF = T // P
    half = N // 2
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N)
    reflect = T > half

    def src(m):
        if m < 0 or m >= T:
            if not reflect:
                return 0.0, 0.0
            m = -m if m < 0 else 2 * (T - 1) - m
        return comb[m], noise[m]

    def frame_spectrum(q):
        z = np.empty(N, np.complex128)
        for i in range(N):
            c, nz = src(q * P - half + i)
            z[i] = win[i] * (c + 1j * nz)
        Z = stockham_fft(z)
        Y = np.empty(half + 1, np.complex128)
        qc = min(q, nF - 1)
        for b in range(half + 1):
            za, zb = Z[b], np.conj(Z[(N - b) % N])
            X = (za + zb) / 2
            d = za - zb
            Nz = complex(d.imag / 2, -d.real / 2)
            Y[b] = X * h_src[qc, b] + Nz * h_noise[qc, b]
        return Y

    out = np.zeros(T)
    sw = lambda z: z.imag + 1j * z.real
    for h0 in range(0, nF, chunk):                      # one CTA per chunk of hops
        h1 = min(h0 + chunk, nF)
        ring = np.zeros(4096)
        qs, qe = max(h0 - 1, 0), min(h1 + 1, nF)
        q = qs
        while q <= qe:
            qa, qb = q, q + 1
            S = np.zeros(N, np.complex128)
            Ya = frame_spectrum(qa)
            S[0], S[half] = Ya[0].real, Ya[half].real
            S[1:half] = Ya[1:half]
            S[half + 1:] = np.conj(Ya[1:half][::-1])
            if qb <= qe:
                Yb = frame_spectrum(qb)
                S[0] += 1j * Yb[0].real
                S[half] += 1j * Yb[half].real
                S[1:half] += 1j * Yb[1:half]
                S[half + 1:] += 1j * np.conj(Yb[1:half][::-1])
            s = sw(stockham_fft(sw(S))) / N
            for i in range(N):
                ring[(qa * P - half + i) % 4096] += s[i].real * win[i]
                if qb <= qe:
                    ring[(qb * P - half + i) % 4096] += s[i].imag * win[i]
            # hops <= qlast-2 are complete; the chunk's last hops complete at the end
            last = min(qb, qe)
            done_upto = last - 2 if last < qe else h1 - 1
            for h in range(max(h0, qa - 2), min(done_upto, h1 - 1) + 1):
                for i in range(P):
                    n = h * P + i
                    env = sum(win[n - qq * P + half] ** 2 for qq in range(h - 1, h + 3) if 0 <= qq <= nF)
                    out[n] = ring[n % 4096] / env
                    ring[n % 4096] = 0.0
            q += 2
    return out


@pytest.mark.parametrize("nF,chunk", [(7, 3), (2, 5), (5, 2)])
def test_superfast_pipeline_matches_oracle(nF, chunk):
    P, sr = 512, 44100
    rng = np.random.default_rng(nF)
    f0 = (150 + 50 * rng.random((1, nF, 1))).astype(np.float32)
    ctrls = {k: rng.standard_normal((1, nF, N // 2 + 1)) * s + m for k, (m, s) in
             {"harmonic_magnitude": (-2, .5), "harmonic_phase": (0, .3), "noise_magnitude": (-3, .5),
              "noise_phase": (0, .3)}.items()}
    noise = rng.standard_normal((1, nF * P))
    want = cf.superfast(f0, ctrls, sr, P, N, noise)
    h_src = np.exp(ctrls["harmonic_magnitude"][0] + 1j * np.pi * ctrls["harmonic_phase"][0])
    h_noise = np.exp(ctrls["noise_magnitude"][0] + 1j * np.pi * ctrls["noise_phase"][0]) / 128
    got = emulate_superfast(want["comb"][0], noise[0], h_src, h_noise, P, chunk)
    assert np.abs(got - want["signal"][0]).max() < 1e-10
