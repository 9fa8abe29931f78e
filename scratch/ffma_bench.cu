// microbenchmark: sustained FFMA vs FFMA2 (fma.rn.f32x2) rate on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float a[16], x = s * threadIdx.x, y = s + 1.f;
    float2 a2[16];
    for (int i = 0; i < 16; ++i) { a[i] = i; a2[i] = make_float2(i, -i); }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
        } else {
            float2 xx = make_float2(x, x), yy = make_float2(y, -y);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int i = 0; i < 16; ++i) { ffma2(a2[i], xx, yy); }
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += (MODE == 0) ? a[i] : (a2[i].x + a2[i].y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) for (int bps = 2; bps <= 8; bps *= 2) {
        int iters = 20000;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * bps, 256>>>(out, iters, 1e-3f); else k<1><<<148 * bps, 256>>>(out, iters, 1e-3f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)148 * bps * 256 * iters * 64 * (mode ? 2 : 1);
        printf("%s blocks/SM=%d warps/SM=%d: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", mode ? "FFMA2" : "FFMA ", bps, bps * 8, ms, fma / ms / 1e9, fma / ms / 1e3 / 148 / 1.965e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
