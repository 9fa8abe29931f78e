// tcgen05 kind::tf32 microtests (sm_100a): (1) plain K-major no-swizzle GEMM 128x8x32,
// (2) Hankel A operand as an overlapping-descriptor view of a linear signal, (3) fp32->tf32 conversion mode
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: A given as [128][K] matrix, staged canonically.  mode 1: A[r][k] = xs[4 r + k] (Hankel view of xs).
__global__ void __launch_bounds__(128) k(const float* __restrict__ Ain, const float* __restrict__ Bin, int K, int mode,
                                       float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* As = reinterpret_cast<float*>(sm);                 // mode 0: [K/4][128][4]; mode 1: linear xs[4*127 + K]
    float* Bs = As + 128 * K + 1024;                          // [K/4][8][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (mode == 0) {
        for (int i = tid; i < 128 * K; i += 128) { int r = i / K, kk = i % K; As[((kk >> 2) * 128 + r) * 4 + (kk & 3)] = Ain[i]; }
    } else {
        for (int i = tid; i < 4 * 127 + K; i += 128) As[i] = Ain[i];
    }
    for (int i = tid; i < 8 * K; i += 128) { int n = i / K, kk = i % K; Bs[((kk >> 2) * 8 + n) * 4 + (kk & 3)] = Bin[i]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 17) | (8u << 24);   // F32 acc, TF32 x TF32, K-major both, N=8, M=128
    if (tid == 0) {
        for (int s = 0; s < K / 8; ++s) {
            uint64_t da, db;
            if (mode == 0) da = make_desc(smem_u32(As) + s * 2 * 2048, 2048, 128);
            else           da = make_desc(smem_u32(As) + s * 32, 16, 128);          // rows 16 B apart, k-chunks 16 B apart
            db = make_desc(smem_u32(Bs) + s * 2 * 128, 128, 128);
            mma_tf32(tmem_d, da, db, idesc, s > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // wait for the MMAs
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[tid * 8 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
#include <cstring>
int main() {
    const int K = 32;
    std::vector<float> A(128 * K), B(8 * K), xs(4 * 127 + K), D(128 * 8);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    float *dA, *dB, *dD; cudaMalloc(&dA, 4 * (128 * K + 2048)); cudaMalloc(&dB, 4 * 8 * K); cudaMalloc(&dD, 4 * 1024);
    size_t smem = (128 * K + 1024 + 8 * K) * 4 + 256;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int test = 0; test < 3; ++test) {
        const int mode = (test == 1) ? 1 : 0;
        for (auto& v : A) v = (test == 2) ? rnd() : tf32_trunc(rnd());
        for (auto& v : B) v = (test == 2) ? rnd() : tf32_trunc(rnd());
        for (auto& v : xs) v = tf32_trunc(rnd());
        if (mode == 1) for (int r = 0; r < 128; ++r) for (int kk = 0; kk < K; ++kk) A[r * K + kk] = xs[4 * r + kk];
        cudaMemcpy(dA, mode == 1 ? xs.data() : A.data(), 4 * (mode == 1 ? xs.size() : A.size()), cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), 4 * B.size(), cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, 4096);
        k<<<1, 128, smem>>>(dA, dB, K, mode, dD);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D.data(), dD, 4096, cudaMemcpyDeviceToHost);
        double e_exact = 0, e_tr = 0, e_rn = 0;
        for (int r = 0; r < 128; ++r) for (int n = 0; n < 8; ++n) {
            double s = 0, st = 0, sr = 0;
            for (int kk = 0; kk < K; ++kk) {
                s += (double)A[r * K + kk] * B[n * K + kk];
                st += (double)tf32_trunc(A[r * K + kk]) * tf32_trunc(B[n * K + kk]);
                sr += (double)tf32_rn(A[r * K + kk]) * tf32_rn(B[n * K + kk]);
            }
            e_exact = fmax(e_exact, fabs(D[r * 8 + n] - s)); e_tr = fmax(e_tr, fabs(D[r * 8 + n] - st)); e_rn = fmax(e_rn, fabs(D[r * 8 + n] - sr));
        }
        printf("test %d (%s): cuda=%s  max|D-exact|=%.3e  max|D-trunc_model|=%.3e  max|D-rn_model|=%.3e   D[0..3]=%g %g %g %g\n", test,
               test == 0 ? "plain canonical" : test == 1 ? "Hankel overlapping descriptor" : "unrounded inputs", cudaGetErrorString(e), e_exact, e_tr, e_rn,
               D[0], D[1], D[2], D[3]);
    }
    return 0;
}
