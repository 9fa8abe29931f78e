// register-only microbenchmark of the FIR inner step: a1[r] += W[3+r-tt]*G[tt]; a2[r] += W[..]*E[tt]
// MODE 0: scalar FFMA as in ltv_fir.cu;  MODE 1: FFMA2 with (a1,a2)/(G,E)/(x,x) pairs
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void ffma2(u64& d, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }
template <int MODE>
__global__ void __launch_bounds__(128, 6) k(const float4* __restrict__ in, float* out, int iters) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    const float4* px = sm + threadIdx.x;        // conflict-free consecutive float4
    const float4* pt = sm + 512;                // broadcast
    float a1[8], a2[8]; u64 acc[8];
    for (int r = 0; r < 8; ++r) { a1[r] = a2[r] = 0.f; acc[r] = pk(0.f, 0.f); }
    float4 A = px[1], B = px[2], C;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const float4 t0 = pt[(2 * (it * 6 + u)) & 255], t1 = pt[(2 * (it * 6 + u) + 1) & 255];
            C = px[(it * 6 + u) & 127];
            const float W[12] = {C.x, C.y, C.z, C.w, A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w};
            if (MODE == 0) {
                const float G[4] = {t0.x, t0.z, t1.x, t1.z}, E[4] = {t0.y, t0.w, t1.y, t1.w};
#pragma unroll
                for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                    for (int r = 0; r < 8; ++r) { float xv = W[3 + r - tt]; a1[r] = fmaf(xv, G[tt], a1[r]); a2[r] = fmaf(xv, E[tt], a2[r]); }
            } else {
                u64 X[12];
#pragma unroll
                for (int j = 0; j < 12; ++j) X[j] = pk(W[j], W[j]);
                const u64 GE[4] = {pk(t0.x, t0.y), pk(t0.z, t0.w), pk(t1.x, t1.y), pk(t1.z, t1.w)};
#pragma unroll
                for (int tt = 0; tt < 4; ++tt)
#pragma unroll
                    for (int r = 0; r < 8; ++r) ffma2(acc[r], X[3 + r - tt], GE[tt]);
            }
            B = A; A = C;
        }
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { r += a1[i] + a2[i]; float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); r += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float4* in; float* out; cudaMalloc(&in, 1024 * 16); cudaMemset(in, 0, 1024 * 16); cudaMalloc(&out, 148 * 6 * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        int iters = 4000; float ms;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 6, 128, 16384>>>(in, out, iters); else k<1><<<148 * 6, 128, 16384>>>(in, out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)148 * 6 * 128 * iters * 6 * 64;
        printf("%s: %.3f ms  %.2f TFMA/s = %.1f%% of 37.2 peak\n", mode ? "FFMA2 pattern" : "FFMA pattern ", ms, fma / ms / 1e9, fma / ms / 1e9 / 37.23 * 100);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
