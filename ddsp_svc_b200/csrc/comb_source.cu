// K5a: band-limited pulse train ("comb-tooth") source of the old CombSub synthesizer.
// Replaces ddsp/vocoder.py:819-829,839-840:
//   x   = wrapped fp64 cumulative phase (cycles), rounded to fp32       (same scan as Sins)
//   f0u = linearly upsampled f0 (fp32)                                   (ddsp/core.py:66-70)
//   comb[t] = sinc(sr * x / (f0u + 1e-3)),  sinc(z) = sin(pi z)/(pi z)   (torch.sinc, fp32)
// The sinc argument amplifies phase error by sr/f0 (up to ~680x), so the phase is evaluated in
// fp64 from the frame scan and rounded once, and the fp32 operation order of the reference is
// kept.  One thread per 4 consecutive samples, 128-bit stores.
#include "b2d_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
comb_source_kernel(const float* __restrict__ f0, const double* __restrict__ frame_phase, int nF, int P,
                   double inv_sr, float sr, int round_fp32, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int T = nF * P;
    const int off = (blockIdx.x * blockDim.x + threadIdx.x) << 2;
    if (off >= T) return;
    const int k = off / P, j = off - k * P;
    const float f0k = f0[(size_t)b * nF + k], f0n = f0[(size_t)b * nF + min(k + 1, nF - 1)];
    const double fk = (double)f0k, dk = (double)f0n - fk, S = frame_phase[(size_t)b * nF + k];
    const double inv2P = 0.5 / (double)P;
    const float invP = 1.0f / (float)P;
    float v[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const double jj = (double)(j + s);
        double x = S + ((jj + 1.0) * fk + dk * (jj * (jj + 1.0)) * inv2P) * inv_sr;
        if (round_fp32) x = (double)(float)x;
        x -= rint(x);
        const float x32 = (float)x;
        const float lam1 = (float)(j + s) * invP, lam0 = 1.0f - lam1;
        const float f0u = __fadd_rn(__fmul_rn(lam0, f0k), __fmul_rn(lam1, f0n));     // upsample
        const float z = __fdiv_rn(__fmul_rn(sr, x32), __fadd_rn(f0u, 1e-3f));         // (:839)
        const float pz = __fmul_rn(B2D_PI_F, z);
        v[s] = (z == 0.0f) ? 1.0f : __fdiv_rn(sinf(pz), pz);
    }
    b2d::st_global_v4(out + (size_t)b * T + off, make_float4(v[0], v[1], v[2], v[3]));
}

}  // namespace

extern "C" int b2d_comb_source(const float* f0_frames, const double* frame_phase, int B, int n_frames, int block,
                               double sampling_rate, int round_fp32, float* comb, void* stream) {
    if (!f0_frames || !frame_phase || !comb) return b2d::fail(B2D_ERR_NULL, "comb_source: null pointer");
    if (B <= 0 || n_frames <= 0 || block <= 0) return b2d::fail(B2D_ERR_SHAPE, "comb_source: bad shape");
    if (block % 4 != 0) return b2d::fail(B2D_ERR_UNSUPPORTED, "comb_source: block size %d must be a multiple of 4", block);
    if (B > 65535) return b2d::fail(B2D_ERR_UNSUPPORTED, "comb_source: batch %d > 65535", B);
    if (!b2d::aligned16(comb)) return b2d::fail(B2D_ERR_ALIGN, "comb_source: output must be 16-byte aligned");
    const long long quads = (long long)n_frames * block / 4;
    comb_source_kernel<<<dim3((unsigned)((quads + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(
        f0_frames, frame_phase, n_frames, block, 1.0 / sampling_rate, (float)sampling_rate, round_fp32, comb);
    return b2d::check_launch("comb_source");
}
