// Small complex DFTs held in registers: the butterflies of the shared-memory Stockham FFTs in superfast.cu
// (2048 points) and combsubfast.cu (1024 points).
#pragma once
#ifndef B2D_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace b2d_fft {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }  // a * (-j)

// forward DFT of 2^n points, natural order in and out (recursive decimation in time, unrolled)
template <int R> struct Dft;
template <> struct Dft<1> { static __device__ __forceinline__ void run(float2*) {} };
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(float2* v) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    }
};
template <int R, int K> __device__ __forceinline__ float2 twid(float2 o) {  // o * exp(-2 pi i K / R)
    if (K == 0) return o;
    if (4 * K == R) return mul_mj(o);
    if (8 * K == R) return make_float2(0.70710678118654752f * (o.x + o.y), 0.70710678118654752f * (o.y - o.x));
    if (8 * K == 3 * R) return make_float2(0.70710678118654752f * (o.y - o.x), -0.70710678118654752f * (o.x + o.y));
    // remaining cases: R = 16, K in {1,3,5,7}
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f;
    const float c = (K == 1) ? c1 : (K == 3) ? s1 : (K == 5) ? -s1 : -c1;
    const float s = (K == 1) ? s1 : (K == 3) ? c1 : (K == 5) ? c1 : s1;
    return make_float2(fmaf(o.x, c, o.y * s), fmaf(o.y, c, -o.x * s));  // (c - j s) * o
}
template <int R, int K> struct Comb {
    static __device__ __forceinline__ void run(const float2* e, const float2* o, float2* v) {
        const float2 t = twid<R, K>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        Comb<R, K + 1>::run(e, o, v);
    }
};
template <int R> struct Comb<R, R / 2> { static __device__ __forceinline__ void run(const float2*, const float2*, float2*) {} };
template <int R> struct Dft {
    static __device__ __forceinline__ void run(float2* v) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<R / 2>::run(e);
        Dft<R / 2>::run(o);
        Comb<R, 0>::run(e, o, v);
    }
};

}  // namespace b2d_fft
