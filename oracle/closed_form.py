"""Float64 numpy restatement of the synthesis path from its closed-form math.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the independent ground
truth used as a tie-breaker: both the reference's fp32 result and the CUDA result
must sit within tolerance of it.  It shares no code path with ``torch_port`` (no
ATen operators; frame-level closed-form phase instead of a per-sample cumsum; the
time-varying FIR as a sum of per-frame linear convolutions instead of FFT
overlap-add) so an indexing error in either shows up as a disagreement.

Formulas follow SURVEY.md appendix A; each function cites the reference lines whose
result it reproduces.  Notation: P = block size, k = frame, j = in-frame index,
v[nF] := v[nF-1] (last frame held).
"""
import numpy as np


def _hold(v):
    return np.concatenate([v, v[..., -1:, :]], axis=-2)


def upsample(v, P):
    """[B, nF, C] -> [B, nF*P, C] linear with last frame held (ddsp/core.py:66-70)."""
    v = np.asarray(v, np.float64)
    ve = _hold(v)
    lam = (np.arange(P) / P)[None, None, :, None]
    out = ve[:, :-1, None, :] * (1 - lam) + ve[:, 1:, None, :] * lam
    return out.reshape(v.shape[0], -1, v.shape[2])


def phase_cycles(f0_frames, sr, P, initial_phase=None):
    """Wrapped phase x[t] in cycles (ddsp/vocoder.py:564-572), closed form per frame.

    inclusive sum of the linearly interpolated f0/sr:
      x[kP+j] = S_k + ((j+1) f_k + (f_{k+1}-f_k) j(j+1)/(2P)) / sr,
      S_k     = sum_{i<k} (P f_i + (f_{i+1}-f_i)(P-1)/2) / sr.
    Returns x wrapped to [-0.5, 0.5] (round-half-even), shape [B, T].
    """
    f = np.asarray(f0_frames, np.float64)[..., 0]
    fe = np.concatenate([f, f[:, -1:]], axis=1)
    d = fe[:, 1:] - fe[:, :-1]
    adv = (P * f + d * (P - 1) / 2.0) / sr
    S = np.concatenate([np.zeros((f.shape[0], 1)), np.cumsum(adv, axis=1)[:, :-1]], axis=1)
    j = np.arange(P, dtype=np.float64)[None, None, :]
    x = S[:, :, None] + ((j + 1) * f[:, :, None] + d[:, :, None] * j * (j + 1) / (2.0 * P)) / sr
    x = x.reshape(f.shape[0], -1)
    if initial_phase is not None:
        x = x + np.asarray(initial_phase, np.float64).reshape(-1, 1) / (2 * np.pi)
    return x - np.rint(x)


def harmonic_amplitudes(c_amp, f0_frames, sr):
    """exp(c)/128 * (1[f0*h < sr/2] + 1e-7)  (ddsp/vocoder.py:580,585; ddsp/core.py:73-77).
    The mask compare is done on the fp32 product like the reference."""
    c = np.asarray(c_amp, np.float64)
    H = c.shape[-1]
    hz = (np.asarray(f0_frames, np.float32) * np.arange(1, H + 1, dtype=np.float32)).astype(np.float32)
    keep = (hz < np.float32(sr / 2)).astype(np.float64) + np.float64(np.float32(1e-7))
    return np.exp(c) / 128.0 * keep


def sinusoid_bank(x_cycles, amp_frames, P):
    """sum_h sin(2 pi h x[t]) * up(A)[t,h]  (ddsp/vocoder.py:586-594)."""
    amp = upsample(amp_frames, P)
    H = amp.shape[-1]
    out = np.zeros(x_cycles.shape, np.float64)
    for h in range(1, H + 1):
        out += np.sin(2 * np.pi * h * x_cycles) * amp[:, :, h - 1]
    return out


def allpass_spectrum(c_gd):
    """exp(j cumsum(pi tanh(c)))  (ddsp/vocoder.py:581,599)."""
    gd = np.pi * np.tanh(np.asarray(c_gd, np.float64))
    return np.exp(1j * np.cumsum(gd, axis=-1))


def impulse_response(spectrum, window="none", half_width=None):
    """irfft -> causal form (ddsp/core.py:254-270).  L = 2(M-1);
    h[tau] = g[(tau - L/2) mod L] * w[tau], with
      none:    w = 1
      hann:    w = 0.5 (1 - cos(2 pi tau / L))                         (core.py:185-237)
      dynamic: u = (tau - L/2)/hw; u := 0 where u > 1; w = 0.5(1+cos(pi u))  (core.py:240-251)
    """
    spectrum = np.asarray(spectrum)
    M = spectrum.shape[-1]
    L = 2 * (M - 1)
    g = np.fft.irfft(spectrum, n=L, axis=-1)
    h = np.roll(g, L // 2, axis=-1)
    tau = np.arange(L, dtype=np.float64)
    if window == "hann":
        h = h * (0.5 * (1 - np.cos(2 * np.pi * tau / L)))
    elif window == "dynamic":
        u = (tau - L // 2) / np.asarray(half_width, np.float64)
        u = np.where(u > 1, 0.0, u)
        h = h * (0.5 * (1 + np.cos(np.pi * u)))
    elif window != "none":
        raise ValueError(window)
    return h


def ltv_fir(audio, ir, P):
    """y = sum_g h_g * (bartlett_g . x), cropped with delay L/2  (ddsp/core.py:120-182).

    Frame g's triangular window peaks at sample g*P; g = 0..nF with h_nF := h_{nF-1}.
    """
    x = np.asarray(audio, np.float64)
    ir = np.asarray(ir, np.float64)
    B, T = x.shape
    nF, L = ir.shape[1], ir.shape[2]
    xp = np.concatenate([np.zeros((B, P)), x, np.zeros((B, P))], axis=1)
    tri = 1.0 - np.abs(np.arange(2 * P) - P) / P
    full = np.zeros((B, T + 2 * P + L), np.float64)   # origin at sample -P
    for b in range(B):
        for g in range(nF + 1):
            seg = xp[b, g * P:g * P + 2 * P] * tri
            if not seg.any():
                continue
            h = ir[b, min(g, nF - 1)]
            full[b, g * P:g * P + 2 * P + L - 1] += np.convolve(seg, h)
    start = P + L // 2
    return full[:, start:start + T]


def sins(f0_frames, ctrls, sr, P, noise, initial_phase=None):
    x = phase_cycles(f0_frames, sr, P, initial_phase)
    x32 = x.astype(np.float32).astype(np.float64)       # the reference rounds x to fp32 (vocoder.py:572)
    amp = harmonic_amplitudes(ctrls["amplitudes"], f0_frames, sr)
    sinusoids = sinusoid_bank(x32, amp, P)
    ir_ap = impulse_response(allpass_spectrum(ctrls["group_delay"]), "none")
    ir_n = impulse_response(np.exp(np.asarray(ctrls["noise_magnitude"], np.float64)) / 128.0, "hann")
    harmonic = ltv_fir(sinusoids, ir_ap, P)
    noise_out = ltv_fir(noise, ir_n, P)
    return {"x": x, "sinusoids": sinusoids, "ir_allpass": ir_ap, "ir_noise": ir_n,
            "harmonic": harmonic, "noise": noise_out, "signal": harmonic + noise_out}


def combsub(f0_frames, ctrls, sr, P, noise, initial_phase=None):
    """(ddsp/vocoder.py:811-862)"""
    x = phase_cycles(f0_frames, sr, P, initial_phase)
    x32 = x.astype(np.float32).astype(np.float64)
    f0_up = upsample(f0_frames, P)[..., 0]
    comb = np.sinc(sr * x32 / (f0_up + 1e-3))
    ir_ap = impulse_response(allpass_spectrum(ctrls["group_delay"]), "none")
    allpassed = ltv_fir(comb, ir_ap, P)
    hw = 1.5 * sr / (np.asarray(f0_frames, np.float64) + 1e-3)
    ir_h = impulse_response(np.exp(np.asarray(ctrls["harmonic_magnitude"], np.float64)), "dynamic", hw)
    harmonic = ltv_fir(allpassed, ir_h, P)
    ir_n = impulse_response(np.exp(np.asarray(ctrls["noise_magnitude"], np.float64)) / 128.0, "hann")
    noise_out = ltv_fir(noise, ir_n, P)
    return {"x": x, "comb": comb, "ir_allpass": ir_ap, "allpassed": allpassed, "ir_harmonic": ir_h,
            "ir_noise": ir_n, "harmonic": harmonic, "noise": noise_out, "signal": harmonic + noise_out}


def combsubfast(f0_frames, ctrls, sr, P, noise, initial_phase=None):
    """(ddsp/vocoder.py:735-786) in float64: frames of 2P at hop P, sqrt-Hann in and out, filter of frame q =
    control row min(q, nF-1), plain overlap-add, cropped by P on both sides."""
    x = phase_cycles(f0_frames, sr, P, initial_phase)
    x32 = x.astype(np.float32).astype(np.float64)           # the reference rounds x to fp32 (:751)
    f0_up = upsample(f0_frames, P)[..., 0]
    comb = np.sinc(sr * x32 / (f0_up + 1e-3))
    B, T = comb.shape
    nF = T // P
    N = 2 * P
    w = np.sqrt(0.5 - 0.5 * np.cos(2 * np.pi * np.arange(N) / N))
    hm = np.asarray(ctrls["harmonic_magnitude"], np.float64)
    hp = np.asarray(ctrls["harmonic_phase"], np.float64)
    nm = np.asarray(ctrls["noise_magnitude"], np.float64)
    h_src = _hold(np.exp(hm + 1j * np.pi * hp))
    h_noise = _hold(np.exp(nm) / 128.0)
    pad = lambda z: np.concatenate([np.zeros((B, P)), np.asarray(z, np.float64), np.zeros((B, P))], axis=1)
    cp, zp = pad(comb), pad(noise)
    out = np.zeros((B, T + 2 * P))
    for q in range(nF + 1):
        seg = slice(q * P, q * P + N)
        spec = np.fft.rfft(cp[:, seg] * w, N) * h_src[:, q] + np.fft.rfft(zp[:, seg] * w, N) * h_noise[:, q]
        out[:, seg] += np.fft.irfft(spec, N) * w
    return {"x": x, "comb": comb, "signal": out[:, P:-P]}


def superfast_phase(f0_frames, sr, P):
    """Wrapped in-frame phase of fast_source_gen in exact arithmetic (ddsp/vocoder.py:639-651)."""
    s = np.asarray(f0_frames, np.float64)[..., 0] / sr
    ds = np.concatenate([s[:, 1:] - s[:, :-1], np.zeros((s.shape[0], 1))], axis=1)
    j = np.arange(P, dtype=np.float64)[None, None, :]
    rad = s[:, :, None] * (j + 1) + 0.5 * ds[:, :, None] * j * (j + 1) / P
    s_up = s[:, :, None] + ds[:, :, None] * j / P
    adv = rad[:, :, -1]
    acc = np.concatenate([np.zeros((s.shape[0], 1)), np.cumsum(adv, axis=1)[:, :-1]], axis=1)
    rad = rad + acc[:, :, None]
    rad = rad - np.rint(rad)
    return rad, s_up


def stft_frames(x, n_fft, hop, pad_mode="reflect"):
    """torch.stft(center=True, periodic Hann) -> [B, n_frames, n_fft//2+1]."""
    half = n_fft // 2
    xp = np.pad(x, ((0, 0), (half, half)), mode=pad_mode)
    n_frames = 1 + (xp.shape[1] - n_fft) // hop
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(n_fft) / n_fft))
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(xp[:, idx] * win, axis=-1)


def istft_frames(spec, n_fft, hop):
    """torch.istft(center=True): OLA(irfft*win) / OLA(win^2), trimmed by n_fft/2."""
    B, n_frames, _ = spec.shape
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(n_fft) / n_fft))
    pieces = np.fft.irfft(spec, n=n_fft, axis=-1) * win
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros((B, total))
    env = np.zeros(total)
    for q in range(n_frames):
        y[:, q * hop:q * hop + n_fft] += pieces[:, q]
        env[q * hop:q * hop + n_fft] += win * win
    half = n_fft // 2
    y, env = y[:, half:total - half], env[half:total - half]
    return y / env


def superfast(f0_frames, ctrls, sr, P, win_length, noise):
    """(ddsp/vocoder.py:653-710)"""
    rad, s_up = superfast_phase(f0_frames, sr, P)
    B = rad.shape[0]
    comb = np.sinc(rad / (s_up + 1e-5)).reshape(B, -1)
    hold = lambda z: np.concatenate([z, z[:, -1:, :]], axis=1)
    c = {k: np.asarray(v, np.float64) for k, v in ctrls.items()}
    h_src = hold(np.exp(c["harmonic_magnitude"] + 1j * np.pi * c["harmonic_phase"]))
    h_noise = hold(np.exp(c["noise_magnitude"] + 1j * np.pi * c["noise_phase"]) / 128.0)
    mode = "reflect" if comb.shape[1] > win_length // 2 else "constant"
    spec = stft_frames(comb, win_length, P, mode) * h_src \
        + stft_frames(np.asarray(noise, np.float64), win_length, P, mode) * h_noise
    return {"comb": comb, "signal": istft_frames(spec, win_length, P)}


def sinegen(f0, upp, sr, rand_ini, noise, sine_amp=0.1, noise_std=0.003, voiced_threshold=0.0):
    """(nsf_hifigan/models.py:134-165) f0 [B, nF]; rand_ini [dim]; noise [B, nF*upp, dim]."""
    f = np.asarray(f0, np.float64)
    B, nF = f.shape
    dim = len(rand_ini)
    step = f / sr
    acc = np.concatenate([np.zeros((B, 1)), np.cumsum(step * upp, axis=1)[:, :-1]], axis=1)
    rad = acc[:, :, None] + step[:, :, None] * np.arange(1, upp + 1)[None, None, :]
    rad = rad.reshape(B, -1, 1) * np.arange(1, dim + 1)[None, None, :] + np.asarray(rand_ini, np.float64)
    uv = np.repeat((f > voiced_threshold).astype(np.float64), upp, axis=1)[:, :, None]
    amp = uv * noise_std + (1 - uv) * sine_amp / 3
    return np.sin(2 * np.pi * rad) * sine_amp * uv + amp * np.asarray(noise, np.float64)
