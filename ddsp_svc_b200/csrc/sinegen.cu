// K6: NSF-HiFiGAN SineGen f0 excitation  (reference nsf_hifigan/models.py:134-165).
//
//   out[b, t, h] = sine_amp * sin(2 pi ((h+1) rad[t] + rho_h)) * uv[t] + noise_amp[t] * eps[b,t,h]
//   rad[k*upp + j] = (f0[k]/sr) (j+1) + acc[k-1],   acc = per-frame wrapped phase advance
//   uv = 1[f0[k] > thr],  noise_amp = uv*noise_std + (1-uv)*sine_amp/3,  eps ~ N(0,1)
//
// The reference evaluates everything in fp32 (piecewise-constant f0, fmod-wrapped frame advance
// accumulated by a cumsum that -- on CPU, the oracle -- accumulates in fp64 and emits fp32); the
// operation order below mirrors it so results agree to the last few ulps of the sin argument.
//
// Two launches: (1) a per-utterance frame scan (n_frames values); (2) the streaming kernel.
// (2) is write-bound (dim*4 = 36 B per audio sample): a CTA owns 128 consecutive samples x dim
// harmonics, each thread computes the `dim` values of one sample into shared memory, and the tile
// is written back as coalesced 128-bit stores.  Noise is either read (parity mode, same tiling)
// or generated in-kernel: Philox4x32-10 + Box-Muller, keyed by (seed, utterance, flat index).
#include "b2d_common.cuh"

namespace {

constexpr int kScanThreads = 256;
constexpr int kTile = 128;   // samples per CTA
constexpr int kMaxDim = 16;

// acc_prev[b,k] = fp32( fmod( fp32( sum_{i<k} adv_i ), 1 ) ),  adv_i = fmod(s_i*upp + 0.5, 1) - 0.5
__global__ void __launch_bounds__(kScanThreads)
sinegen_scan_kernel(const float* __restrict__ f0, int nF, int upp, float sr, float* __restrict__ acc_prev) {
    const int b = blockIdx.x;
    const float* f = f0 + (size_t)b * nF;
    const int per = (nF + kScanThreads - 1) / kScanThreads;
    const int k0 = min(nF, (int)threadIdx.x * per), k1 = min(nF, k0 + per);
    const float fupp = (float)upp;
    auto adv = [&](int k) {
        const float s = __fdiv_rn(f[k], sr);
        const float last = __fmul_rn(s, fupp);                       // rad[k, upp-1]
        return __fsub_rn(fmodf(__fadd_rn(last, 0.5f), 1.0f), 0.5f);  // (:139)
    };
    double local = 0.0;
    for (int k = k0; k < k1; ++k) local += (double)adv(k);
    __shared__ double warp_tot[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double run = incl - local;
    for (int w = 0; w < warp; ++w) run += warp_tot[w];
    for (int k = k0; k < k1; ++k) {
        // exclusive: what the reference adds to frame k is fmod(cumsum[k-1], 1)   (:140-141)
        acc_prev[(size_t)b * nF + k] = (k == 0) ? 0.0f : fmodf((float)run, 1.0f);
        run += (double)adv(k);
    }
}

struct SgParams {
    const float* f0;
    const float* acc_prev;
    const float* rand_ini;  // [dim]
    const float* noise_in;  // [B, T, dim] or nullptr
    float* out;             // [B, T, dim]
    int nF, upp, dim;
    int upp_shift;          // log2(upp) if upp is a power of two, else -1
    float sr, sine_amp;
    float namp_voiced, namp_unvoiced;   // uv*noise_std + (1-uv)*sine_amp/3 for uv = 1 / 0 (fp32, reference op order)
    float thr;
    unsigned long long seed;
    long long utt_off;
};

// sin(arg) for |arg| up to a few thousand rad: two-term Cody-Waite reduction, then the SFU.
__device__ __forceinline__ float sin_reduced(float arg) {
    const float k = rintf(arg * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, arg);   // 2pi hi (fp32)
    r = fmaf(k, 1.7484555e-7f, r);                  // -(2pi - hi)
    return __sinf(r);
}

// DIM > 0: compile-time number of harmonics (fully unrolled, no predication); DIM == 0: runtime p.dim <= 16
template <int DIM>
__global__ void __launch_bounds__(kTile) sinegen_kernel(SgParams p) {
    extern __shared__ __align__(16) float tile[];  // [kTile * dim]
    constexpr int MAXD = DIM > 0 ? DIM : kMaxDim;
    const int dim = DIM > 0 ? DIM : p.dim;
    const int b = blockIdx.y;
    const int T = p.nF * p.upp;
    const int t0 = blockIdx.x * kTile;
    const int nt = min(kTile, T - t0);
    const int tid = threadIdx.x;
    const size_t base = ((size_t)b * T + t0) * dim;
    const int nflat = nt * dim;

    if (p.noise_in) {  // stage the noise tile (coalesced)
        const float* src = p.noise_in + base;
        if ((base & 3) == 0 && (nflat & 3) == 0) {
            for (int i = tid; i < (nflat >> 2); i += kTile)
                reinterpret_cast<float4*>(tile)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        } else {
            for (int i = tid; i < nflat; i += kTile) tile[i] = src[i];
        }
        __syncthreads();
    }

    if (tid < nt) {
        const int t = t0 + tid;
        int k, j;
        if (p.upp_shift >= 0) { k = t >> p.upp_shift; j = t & (p.upp - 1); }
        else { k = t / p.upp; j = t - k * p.upp; }
        const float f = p.f0[(size_t)b * p.nF + k];
        const float s = __fdiv_rn(f, p.sr);                                             // f0 / sr
        const float rad = __fadd_rn(__fmul_rn(s, (float)(j + 1)), p.acc_prev[(size_t)b * p.nF + k]);  // (:138,141)
        const bool voiced = f > p.thr;
        const float namp = voiced ? p.namp_voiced : p.namp_unvoiced;                    // (:162)
        const float samp_uv = voiced ? p.sine_amp : 0.0f;   // (sin*sine_amp)*uv == sin*(sine_amp*uv) for uv in {0,1}
        float* row = tile + tid * dim;
        float eps[MAXD + 3];
        if (p.noise_in) {
#pragma unroll
            for (int h = 0; h < MAXD; ++h) if (h < dim) eps[h] = row[h];
        } else {
            // 4 normals per Philox call; counter = (sample index, call index), key = seed,
            // stream = utterance.  ceil(dim/4) calls per sample (3 for dim = 9; 3 normals unused).
            const unsigned long long utt = (unsigned long long)(p.utt_off + b);
#pragma unroll
            for (int c = 0; c < (MAXD + 3) / 4; ++c) {
                if (4 * c < dim) {
                    uint4 r = b2d::philox4x32_10(make_uint4((uint32_t)t, 0x51e6e000u + c, (uint32_t)utt, (uint32_t)(utt >> 32)),
                                                 make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
                    // Box-Muller: u1,u3 in (0,1], u2,u4 in [0,1);  sqrt(a) = a * rsqrt(a), a > 0
                    const float u1 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
                    const float u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
                    const float u3 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
                    const float u4 = (float)(r.w >> 8) * (1.0f / 16777216.0f);
                    const float a1 = fmaxf(-2.0f * __logf(u1), 1e-30f), a3 = fmaxf(-2.0f * __logf(u3), 1e-30f);
                    const float m1 = a1 * rsqrtf(a1), m2 = a3 * rsqrtf(a3);
                    float s1, c1, s2, c2;
                    __sincosf(B2D_TWO_PI_F * u2, &s1, &c1);
                    __sincosf(B2D_TWO_PI_F * u4, &s2, &c2);
                    eps[4 * c + 0] = m1 * c1; eps[4 * c + 1] = m1 * s1;
                    eps[4 * c + 2] = m2 * c2; eps[4 * c + 3] = m2 * s2;
                }
            }
        }
#pragma unroll
        for (int h = 0; h < MAXD; ++h) {
            if (h < dim) {
                const float theta = __fadd_rn(__fmul_rn(rad, (float)(h + 1)), __ldg(p.rand_ini + h));   // (:143,146)
                const float sn = sin_reduced(__fmul_rn(B2D_TWO_PI_F, theta));                             // (:147)
                row[h] = __fadd_rn(__fmul_rn(sn, samp_uv), __fmul_rn(namp, eps[h]));                      // (:159,163-164)
            }
        }
    }
    __syncthreads();
    float* dst = p.out + base;
    if ((base & 3) == 0 && (nflat & 3) == 0) {
        for (int i = tid; i < (nflat >> 2); i += kTile)
            b2d::st_global_v4(dst + 4 * i, reinterpret_cast<const float4*>(tile)[i]);
    } else {
        for (int i = tid; i < nflat; i += kTile) dst[i] = tile[i];
    }
}

}  // namespace

extern "C" int b2d_sinegen(const float* f0, const float* rand_ini, const float* noise_in, uint64_t seed,
                           int64_t utterance_offset, int B, int n_frames, int upp, int dim,
                           double sampling_rate, float sine_amp, float noise_std, float voiced_threshold,
                           float* acc_workspace, float* out, void* stream) {
    if (!f0 || !rand_ini || !acc_workspace || !out) return b2d::fail(B2D_ERR_NULL, "sinegen: null pointer");
    if (B <= 0 || n_frames <= 0 || upp <= 0 || dim <= 0) return b2d::fail(B2D_ERR_SHAPE, "sinegen: bad shape");
    if (dim > kMaxDim) return b2d::fail(B2D_ERR_UNSUPPORTED, "sinegen: dim %d > %d", dim, kMaxDim);
    if (B > 65535) return b2d::fail(B2D_ERR_UNSUPPORTED, "sinegen: batch %d > 65535", B);
    if (!b2d::aligned16(out) || (noise_in && !b2d::aligned16(noise_in)))
        return b2d::fail(B2D_ERR_ALIGN, "sinegen: out / noise_in must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    sinegen_scan_kernel<<<B, kScanThreads, 0, st>>>(f0, n_frames, upp, (float)sampling_rate, acc_workspace);
    int rc = b2d::check_launch("sinegen_scan");
    if (rc) return rc;
    SgParams p;
    p.f0 = f0; p.acc_prev = acc_workspace; p.rand_ini = rand_ini; p.noise_in = noise_in; p.out = out;
    p.nF = n_frames; p.upp = upp; p.dim = dim;
    p.sr = (float)sampling_rate; p.sine_amp = sine_amp; p.thr = voiced_threshold;
    p.upp_shift = -1;
    for (int sft = 0; sft < 30; ++sft) if ((1 << sft) == upp) p.upp_shift = sft;
    // noise_amp = uv*noise_std + (1-uv)*sine_amp/3 in fp32 with the reference's operation order (:162)
    p.namp_voiced = (1.0f * noise_std) + ((0.0f * sine_amp) / 3.0f);
    p.namp_unvoiced = (0.0f * noise_std) + ((1.0f * sine_amp) / 3.0f);
    p.seed = seed; p.utt_off = utterance_offset;
    const long long T = (long long)n_frames * upp;
    const dim3 grid((unsigned)((T + kTile - 1) / kTile), B);
    const size_t smem = kTile * dim * sizeof(float);
    if (dim == 9) sinegen_kernel<9><<<grid, kTile, smem, st>>>(p);
    else if (dim == 1) sinegen_kernel<1><<<grid, kTile, smem, st>>>(p);
    else sinegen_kernel<0><<<grid, kTile, smem, st>>>(p);
    return b2d::check_launch("sinegen");
}
