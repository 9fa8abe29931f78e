// K1: frame-rate exciter phase.  Replaces the per-sample fp64 cumsum of the upsampled f0
// (reference ddsp/vocoder.py:564-575) by its closed form: the sum over one frame of the
// linearly interpolated f0 is P*f_k + (f_{k+1}-f_k)(P-1)/2, so only an n_frames-long fp64
// exclusive scan per utterance is needed; the in-frame part is evaluated per sample by the
// synthesis kernels.  One CTA per utterance: thread-serial chunk sums + warp-shuffle scan.
#include "b2d_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
phase_scan_kernel(const float* __restrict__ f0, const float* __restrict__ init_phase, int nF, int P,
                  double sr, int round_fp32, double* __restrict__ frame_phase,
                  float* __restrict__ phase_frames) {
    const int b = blockIdx.x;
    const float* f = f0 + (size_t)b * nF;
    const int per = (nF + kThreads - 1) / kThreads;
    const int k0 = min(nF, (int)threadIdx.x * per), k1 = min(nF, k0 + per);
    const double half_pm1 = 0.5 * (double)(P - 1);

    double local = 0.0;
    for (int k = k0; k < k1; ++k) {
        double fk = (double)f[k], fn = (double)f[min(k + 1, nF - 1)];
        local += ((double)P * fk + (fn - fk) * half_pm1) / sr;
    }
    // block-wide exclusive scan of `local`
    __shared__ double warp_tot[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        double up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    double run = base + (incl - local);
    if (init_phase) run += (double)init_phase[b] / 2.0 / 3.14159265358979323846;

    for (int k = k0; k < k1; ++k) {
        double fk = (double)f[k], fn = (double)f[min(k + 1, nF - 1)];
        frame_phase[(size_t)b * nF + k] = run;
        double x0 = run + fk / sr;  // first sample of the frame (inclusive scan)
        if (round_fp32) x0 = (double)(float)x0;
        x0 -= rint(x0);
        phase_frames[(size_t)b * nF + k] = (float)x0 * B2D_TWO_PI_F;
        run += ((double)P * fk + (fn - fk) * half_pm1) / sr;
    }
}

}  // namespace

extern "C" int b2d_phase_scan(const float* f0_frames, const float* initial_phase, int B, int n_frames,
                              int block, double sampling_rate, int round_fp32, double* frame_phase,
                              float* phase_frames, void* stream) {
    if (!f0_frames || !frame_phase || !phase_frames) return b2d::fail(B2D_ERR_NULL, "phase_scan: null pointer");
    if (B <= 0 || n_frames <= 0 || block <= 0 || !(sampling_rate > 0))
        return b2d::fail(B2D_ERR_SHAPE, "phase_scan: bad shape B=%d nF=%d block=%d", B, n_frames, block);
    phase_scan_kernel<<<B, kThreads, 0, (cudaStream_t)stream>>>(f0_frames, initial_phase, n_frames, block,
                                                                 sampling_rate, round_fp32, frame_phase,
                                                                 phase_frames);
    return b2d::check_launch("phase_scan");
}
