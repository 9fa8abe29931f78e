// 1024-point complex FFT in shared memory for a 128-thread CTA: three in-place Stockham passes (radix 16, 8, 8)
// over NBATCH independent transforms stored back to back, data padded by one slot per 16 so every pass is
// bank-conflict free.  Used by combsubfast.cu and ltv_fir_fft.cu; index formulas pinned by tests/test_csfast_math.py
// and by the host emulation of both kernels (tests/emu/).
#pragma once
#include "fft_regs.cuh"

namespace b2d_fft1024 {
using namespace b2d_fft;

constexpr int kN = 1024, kThreads = 128;
constexpr int kPad = kN + kN / 16;      // complex slots of one padded FFT buffer
constexpr int kTw2 = 7 * 16;            // exp(-2 pi i r k / 128), r = 1..7, k < 16
constexpr int kTw3 = 128;               // exp(-2 pi i k / 1024), k < 128

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }   // one pad slot per 16: conflict-free passes

// One radix-R Stockham pass over NBATCH independent 1024-point FFTs stored back to back (FFT g at buf + g*kPad),
// in place.
//   butterfly j: v[r] = in[j + r N/R] * exp(-2 pi i r (j % NS) / (NS R));  DFT_R;  out[(j/NS) NS R + j%NS + r NS] = v[r]
// TW: 0 none (NS = 1), 1 full table tw[(r-1) NS + k], 2 powers of tw[k] = exp(-2 pi i k / (NS R))
// The batch is processed in stages of 128 butterflies (= one FFT for R = 8, two for R = 16).  A stage reads, hits a
// barrier, then writes; different stages touch different FFTs, so stage s+1 may start reading while other threads
// still write stage s, and only ONE butterfly per thread is live (R complex registers, not R x batch).
// Barriers per pass: stages + 1.
template <int R, int NS, int TW, int NBATCH>
__device__ __forceinline__ void fft_pass(float2* buf, const float2* __restrict__ tw, int tid) {
    constexpr int NB = kN / R;                                   // butterflies per FFT
    constexpr int TOTAL = NB * NBATCH;
    constexpr int STAGES = (TOTAL + kThreads - 1) / kThreads;
    static_assert(kThreads % NB == 0, "a stage must hold whole FFTs");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
        const int idx = tid + s * kThreads;
        const bool active = idx < TOTAL;
        const int g = idx / NB, j = idx % NB, k = j % NS;
        float2* fft = buf + g * kPad;
        float2 v[R];
        if (active) {
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = fft[padi(j + r * NB)];
            if (TW == 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) v[r] = cmul(v[r], tw[(r - 1) * NS + k]);
            } else if (TW == 2) {
                const float2 w1 = tw[k];
                const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
                const float2 w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
                v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[3] = cmul(v[3], w3); v[4] = cmul(v[4], w4);
                v[5] = cmul(v[5], w5); v[6] = cmul(v[6], w6); v[7] = cmul(v[7], w7);
            }
            Dft<R>::run(v);
        }
        __syncthreads();
        if (active) {
            const int base = (j / NS) * NS * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) fft[padi(base + r * NS)] = v[r];
        }
    }
    __syncthreads();
}

template <int NBATCH>
__device__ __forceinline__ void fft1024(float2* buf, const float2* tw2, const float2* tw3, int tid) {
    fft_pass<16, 1, 0, NBATCH>(buf, nullptr, tid);
    fft_pass<8, 16, 1, NBATCH>(buf, tw2, tid);
    fft_pass<8, 128, 2, NBATCH>(buf, tw3, tid);
}


// twiddle tables of passes 2 and 3 (call with all threads, then __syncthreads)
__device__ __forceinline__ void init_twiddles(float2* tw2, float2* tw3, int tid) {
    for (int i = tid; i < kTw2; i += kThreads) {
        const int r = i / 16 + 1, k = i % 16;
        float sn, cs; sincospif(-2.0f * (float)(r * k) / 128.0f, &sn, &cs);
        tw2[i] = make_float2(cs, sn);
    }
    for (int i = tid; i < kTw3; i += kThreads) {
        float sn, cs; sincospif(-2.0f * (float)i / 1024.0f, &sn, &cs);
        tw3[i] = make_float2(cs, sn);
    }
}

}  // namespace b2d_fft1024
