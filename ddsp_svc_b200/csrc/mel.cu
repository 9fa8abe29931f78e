// K8: log-mel spectrogram front end of the NSF-HiFiGAN vocoder -- replaces STFT.get_mel (nsf_hifigan/nvSTFT.py:73-117),
// the consumer of the synthesizer's waveform in enhancer.py:113, diffusion/vocoder.py:147 and reflow/vocoder.py:125:
//
//   y_pad  = pad(y, (win - hop) / 2 each side, reflect | constant)                     (:97-104)
//   S      = stft(y_pad, n_fft = win = 2048, hop, hann (periodic), center = False)       (:106-107)
//   mag    = sqrt(Re^2 + Im^2 + 1e-9)                                                    (:108)
//   mel    = log(clamp(mel_basis [n_mels, 1025] @ mag, clip_val))                        (:114-115)
//
// for the shape every shipped configuration uses (keyshift 0, speed 1: n_fft = win = 2048; any hop, n_mels <= 128).
// One CTA (128 threads) owns 16 consecutive frames of one utterance and transforms them TWO at a time: a complex
// 2048-point FFT of w (ya + j yb) in shared memory (fft_smem.cuh, radix 16 x 16 x 8) carries both real frames, the two
// spectra are separated by conjugate symmetry, and thread m reduces its mel filter over the bins where it is non-zero
// (the triangular filters are sparse: ~2 050 of 131 200 weights; [lo, hi) per filter comes from the host) for both
// frames.  The 128 x 16 results are staged in shared memory and written as 64-byte row segments of the
// [B, n_mels, n_frames] output.  HBM: the waveform is read once (the 4x frame overlap is served by L1/L2), 4 n_mels /
// hop bytes per input sample are written: 4 + 1 = 5 B per sample for 128 mels at hop 512.
#include "b2d_common.cuh"
#include "fft_smem.cuh"

using namespace b2d_fft;
using b2d_fft_smem::kThreads;
using b2d_fft_smem::padi;

namespace {

constexpr int kN = 2048, kBins = kN / 2 + 1;
constexpr int kPad = b2d_fft_smem::Plan<kN>::kPad, kTw2 = b2d_fft_smem::Plan<kN>::kTw2, kTw3 = b2d_fft_smem::Plan<kN>::kTw3;
constexpr int kFramesPerCta = 16;
constexpr int kMagStride = kBins + 3;                       // 1028 floats per frame of magnitudes
constexpr size_t kSmemBytes = (size_t)kPad * sizeof(float2) + (size_t)(kTw2 + kTw3) * sizeof(float2) +
                              (size_t)2 * kMagStride * sizeof(float) + (size_t)128 * (kFramesPerCta + 1) * sizeof(float);

struct MelParams {
    const float* y;            // [B, T]
    const float* window;       // [2048] hann, periodic (torch.hann_window)
    const float* basis;        // [n_mels, 1025]
    const int* lohi;           // [n_mels, 2] first / one-past-last non-zero bin of each filter
    float* out;                // [B, n_mels, n_frames]
    int T, hop, n_frames, n_mels, pad_left, reflect;
    float clip;
};

__device__ __forceinline__ float sample_at(const float* __restrict__ y, int T, int src, int reflect_mode) {
    if (src >= 0 && src < T) return __ldg(y + src);
    if (!reflect_mode) return 0.f;
    const int r = src < 0 ? -src : 2 * (T - 1) - src;       // torch 'reflect' (pad < T is guaranteed by the host check)
    return __ldg(y + r);
}

__global__ void __launch_bounds__(kThreads, 3) mel_kernel(MelParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* F = reinterpret_cast<float2*>(smem_raw);
    float2* tw2 = F + kPad;
    float2* tw3 = tw2 + kTw2;
    float* mag = reinterpret_cast<float*>(tw3 + kTw3);        // [2][kMagStride]
    float* stage = mag + 2 * kMagStride;                      // [128][kFramesPerCta + 1]
    const int tid = threadIdx.x, b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta, f1 = min(f0 + kFramesPerCta, p.n_frames);
    const float* y = p.y + (size_t)b * p.T;

    b2d_fft_smem::init_twiddles<kN>(tw2, tw3, tid);
    int lo = 0, hi = 0;
    if (tid < p.n_mels) { lo = p.lohi[2 * tid]; hi = p.lohi[2 * tid + 1]; }
    const float* brow = p.basis + (size_t)min(tid, p.n_mels - 1) * kBins;

    for (int fa = f0; fa < f1; fa += 2) {
        const bool has_b = fa + 1 < f1;
        // ---- windowed frames fa (real part) and fa + 1 (imaginary part) ----
        const int s0 = fa * p.hop - p.pad_left;
        if (s0 >= 0 && s0 + p.hop + kN <= p.T && has_b) {
            // interior pair (all but the first / last frames): no bounds logic, all 48 loads of a thread in flight at once
            float va[kN / kThreads], vb[kN / kThreads], w[kN / kThreads];
#pragma unroll
            for (int u = 0; u < kN / kThreads; ++u) {
                const int n = tid + u * kThreads;
                w[u] = __ldg(p.window + n);
                va[u] = __ldg(y + s0 + n);
                vb[u] = __ldg(y + s0 + p.hop + n);
            }
#pragma unroll
            for (int u = 0; u < kN / kThreads; ++u) F[padi(tid + u * kThreads)] = make_float2(w[u] * va[u], w[u] * vb[u]);
        } else {
#pragma unroll 4
            for (int u = 0; u < kN / kThreads; ++u) {
                const int n = tid + u * kThreads;
                const float w = __ldg(p.window + n);
                const float va = sample_at(y, p.T, s0 + n, p.reflect);
                const float vb = has_b ? sample_at(y, p.T, s0 + n + p.hop, p.reflect) : 0.f;
                F[padi(n)] = make_float2(w * va, w * vb);
            }
        }
        __syncthreads();
        b2d_fft_smem::fft_forward<kN, 1, true>(F, tw2, tw3, tid);
        // ---- |A[k]|, |B[k]| from Z = FFT(a + j b):  A = (Z[k] + conj Z[N-k]) / 2,  B = (Z[k] - conj Z[N-k]) / 2j ----
        for (int k = tid; k < kBins; k += kThreads) {
            const float2 zk = F[padi(k)], zm = F[padi((kN - k) & (kN - 1))];
            const float ar = 0.5f * (zk.x + zm.x), ai = 0.5f * (zk.y - zm.y);
            const float br = 0.5f * (zk.y + zm.y), bi = 0.5f * (zm.x - zk.x);
            mag[k] = sqrtf(ar * ar + ai * ai + 1e-9f);
            mag[kMagStride + k] = sqrtf(br * br + bi * bi + 1e-9f);
        }
        __syncthreads();
        // ---- mel projection: thread m over the support of its filter, both frames ----
        if (tid < p.n_mels) {
            float sa = 0.f, sb = 0.f;
            for (int k = lo; k < hi; ++k) {
                const float wgt = __ldg(brow + k);
                sa = fmaf(wgt, mag[k], sa);
                sb = fmaf(wgt, mag[kMagStride + k], sb);
            }
            stage[tid * (kFramesPerCta + 1) + (fa - f0)] = logf(fmaxf(sa, p.clip));
            stage[tid * (kFramesPerCta + 1) + (fa - f0) + 1] = logf(fmaxf(sb, p.clip));
        }
        // the next iteration's loads write F only; mag is rewritten after two more barriers: no barrier needed here
    }
    __syncthreads();
    // ---- write out: row segments of up to 16 frames ----
    const int nfr = f1 - f0;
    for (int i = tid; i < p.n_mels * kFramesPerCta; i += kThreads) {
        const int m = i / kFramesPerCta, c = i - m * kFramesPerCta;
        if (c < nfr) p.out[((size_t)b * p.n_mels + m) * p.n_frames + f0 + c] = stage[m * (kFramesPerCta + 1) + c];
    }
}

}  // namespace

extern "C" int b2d_mel_frames(int n_samples, int n_fft, int win_size, int hop) {
    if (n_samples <= 0 || hop <= 0 || win_size < hop || n_fft < win_size) return 0;
    const int pad_left = (win_size - hop) / 2;
    int pad_right = (win_size - hop + 1) / 2;
    if (win_size - n_samples - pad_left > pad_right) pad_right = win_size - n_samples - pad_left;
    const long long padded = (long long)n_samples + pad_left + pad_right;
    return padded < n_fft ? 0 : (int)(1 + (padded - n_fft) / hop);
}

extern "C" int b2d_mel_spectrogram(const float* audio, const float* window, const float* mel_basis, const int* filter_lohi,
                                   int B, int n_samples, int n_fft, int win_size, int hop, int n_mels, float clip_val,
                                   float* mel, void* stream) {
    if (!audio || !window || !mel_basis || !filter_lohi || !mel) return b2d::fail(B2D_ERR_NULL, "mel_spectrogram: null pointer");
    if (n_fft != kN || win_size != kN)
        return b2d::fail(B2D_ERR_UNSUPPORTED, "mel_spectrogram: only n_fft = win_size = %d is built (keyshift 0), got %d / %d", kN, n_fft, win_size);
    if (B <= 0 || B > 65535 || n_samples <= 0 || hop <= 0 || hop > kN || n_mels <= 0 || n_mels > 128)
        return b2d::fail(B2D_ERR_SHAPE, "mel_spectrogram: bad shape B=%d T=%d hop=%d n_mels=%d (n_mels <= 128)", B, n_samples, hop, n_mels);
    MelParams p;
    p.y = audio; p.window = window; p.basis = mel_basis; p.lohi = filter_lohi; p.out = mel;
    p.T = n_samples; p.hop = hop; p.n_mels = n_mels; p.clip = clip_val;
    p.pad_left = (win_size - hop) / 2;
    int pad_right = (win_size - hop + 1) / 2;
    if (win_size - n_samples - p.pad_left > pad_right) pad_right = win_size - n_samples - p.pad_left;
    p.reflect = pad_right < n_samples ? 1 : 0;                  // nvSTFT.py:99-102
    if (p.reflect && (p.pad_left >= n_samples || pad_right >= n_samples))
        return b2d::fail(B2D_ERR_SHAPE, "mel_spectrogram: reflect padding needs more than %d samples", p.pad_left);
    p.n_frames = b2d_mel_frames(n_samples, n_fft, win_size, hop);
    if (p.n_frames <= 0) return b2d::fail(B2D_ERR_SHAPE, "mel_spectrogram: signal too short");
    cudaError_t e = cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return b2d::fail((int)e, "mel_spectrogram: smem attr: %s", cudaGetErrorString(e));
    dim3 grid((p.n_frames + kFramesPerCta - 1) / kFramesPerCta, B);
    mel_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(p);
    return b2d::check_launch("mel_spectrogram");
}
