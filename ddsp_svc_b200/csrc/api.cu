// C-ABI plumbing of libb200ddsp: version, thread-local error string, and the drivers that
// chain the kernels of one synthesizer on the caller's stream.
#include <stdarg.h>
#include <string.h>

#include "b2d_common.cuh"

namespace b2d {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int ltv_fir_launch(const float* x1, const float* ir1, int taps1, float* y1, const float* x2, const float* ir2,
                   int taps2, float* y2, const float* addend, float* mix, uint64_t seed, int64_t utt_off, int B,
                   int nF, int P, cudaStream_t st);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace b2d

namespace b2d { bool g_fft_packed = false; }

extern "C" int b2d_set_fft_arith(int packed) {
    if (packed != 0 && packed != 1) return b2d::fail(B2D_ERR_UNSUPPORTED, "set_fft_arith: %d not in {0, 1}", packed);
    b2d::g_fft_packed = packed != 0;
    return 0;
}

extern "C" int b2d_version(void) { return B2D_VERSION; }
extern "C" const char* b2d_last_error(void) { return b2d::err_buf(); }

// ---------------------------------------------------------------------------------------
// Sins: bank -> all-pass IR -> noise IR -> two FIRs + mix        (ddsp/vocoder.py:580-611)
// workspace: sinusoids [B,T] | ir_allpass [B,nF,2(Ma-1)] | ir_noise [B,nF,2(Mn-1)]
// ---------------------------------------------------------------------------------------
extern "C" size_t b2d_sins_workspace_bytes(int B, int n_frames, int block, int n_mag_allpass, int n_mag_noise) {
    if (B <= 0 || n_frames <= 0 || block <= 0 || n_mag_allpass < 2 || n_mag_noise < 2) return 0;
    const size_t BT = (size_t)B * n_frames * block, BF = (size_t)B * n_frames;
    return b2d::align_up(BT * 4, 256) + b2d::align_up(BF * 2 * (n_mag_allpass - 1) * 4, 256) +
           b2d::align_up(BF * 2 * (n_mag_noise - 1) * 4, 256);
}

extern "C" int b2d_sins_synth(const float* f0_frames, const double* frame_phase, const float* c_amp,
                              const float* c_group_delay, const float* c_noise, int64_t ctrl_stride,
                              const float* noise_in, uint64_t seed, int64_t utterance_offset,
                              const float* dft_tables_allpass, const float* dft_tables_noise, int B, int n_frames,
                              int block, int n_harmonics, int n_mag_allpass, int n_mag_noise,
                              double sampling_rate, int round_fp32, float* signal, float* harmonic,
                              float* noise_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace) return b2d::fail(B2D_ERR_NULL, "sins_synth: null workspace");
    const size_t need = b2d_sins_workspace_bytes(B, n_frames, block, n_mag_allpass, n_mag_noise);
    if (need == 0) return b2d::fail(B2D_ERR_SHAPE, "sins_synth: bad shape");
    if (workspace_bytes < need) return b2d::fail(B2D_ERR_WORKSPACE, "sins_synth: workspace %zu < %zu bytes", workspace_bytes, need);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return b2d::fail(B2D_ERR_ALIGN, "sins_synth: workspace must be 256-byte aligned");
    const size_t BT = (size_t)B * n_frames * block, BF = (size_t)B * n_frames;
    const int La = 2 * (n_mag_allpass - 1), Ln = 2 * (n_mag_noise - 1);
    char* ws = static_cast<char*>(workspace);
    float* sinus = reinterpret_cast<float*>(ws);
    float* ir_ap = reinterpret_cast<float*>(ws + b2d::align_up(BT * 4, 256));
    float* ir_n = reinterpret_cast<float*>(ws + b2d::align_up(BT * 4, 256) + b2d::align_up(BF * La * 4, 256));

    int rc = b2d_sins_bank(f0_frames, frame_phase, c_amp, ctrl_stride, B, n_frames, block, n_harmonics,
                           sampling_rate, round_fp32, sinus, stream);
    if (rc) return rc;
    rc = b2d_ir_build(c_group_delay, ctrl_stride, B2D_IR_ALLPASS, nullptr, dft_tables_allpass, B, n_frames,
                      n_mag_allpass, sampling_rate, ir_ap, stream);
    if (rc) return rc;
    rc = b2d_ir_build(c_noise, ctrl_stride, B2D_IR_MAG_HANN, nullptr, dft_tables_noise, B, n_frames, n_mag_noise,
                      sampling_rate, ir_n, stream);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (La == Ln && block % 256 == 0) {
        return b2d::ltv_fir_launch(sinus, ir_ap, La, harmonic, noise_in, ir_n, Ln, noise_out, nullptr, signal, seed,
                                   utterance_offset, B, n_frames, block, st);
    }
    if (block % 256 != 0)
        return b2d::fail(B2D_ERR_UNSUPPORTED, "sins_synth: block size %d must be a multiple of 256", block);
    // different tap counts: two launches, the second adds the first's output
    if (!noise_out) return b2d::fail(B2D_ERR_UNSUPPORTED, "sins_synth: noise_out required when n_mag_allpass != n_mag_noise");
    rc = b2d::ltv_fir_launch(noise_in, ir_n, Ln, noise_out, nullptr, nullptr, 0, nullptr, nullptr, nullptr, seed,
                             utterance_offset, B, n_frames, block, st);
    if (rc) return rc;
    return b2d::ltv_fir_launch(sinus, ir_ap, La, harmonic, nullptr, nullptr, 0, nullptr, noise_out, signal, seed,
                               utterance_offset, B, n_frames, block, st);
}

// ---------------------------------------------------------------------------------------
// CombSub (old): comb source -> all-pass FIR -> dynamic-window harmonic FIR, + noise FIR
// (ddsp/vocoder.py:834-862).
// workspace: comb [B,T] | allpassed [B,T] | noise [B,T] | ir_ap | ir_h | ir_n
// ---------------------------------------------------------------------------------------
extern "C" int b2d_comb_source(const float*, const double*, int, int, int, double, int, float*, void*);

extern "C" size_t b2d_combsub_workspace_bytes(int B, int n_frames, int block, int n_mag_allpass,
                                              int n_mag_harmonic, int n_mag_noise) {
    if (B <= 0 || n_frames <= 0 || block <= 0 || n_mag_allpass < 2 || n_mag_harmonic < 2 || n_mag_noise < 2) return 0;
    const size_t BT = (size_t)B * n_frames * block, BF = (size_t)B * n_frames;
    return 3 * b2d::align_up(BT * 4, 256) + b2d::align_up(BF * 2 * (n_mag_allpass - 1) * 4, 256) +
           b2d::align_up(BF * 2 * (n_mag_harmonic - 1) * 4, 256) + b2d::align_up(BF * 2 * (n_mag_noise - 1) * 4, 256);
}

extern "C" int b2d_combsub_synth(const float* f0_frames, const double* frame_phase, const float* c_group_delay,
                                 const float* c_harmonic, const float* c_noise, int64_t ctrl_stride,
                                 const float* noise_in, uint64_t seed, int64_t utterance_offset,
                                 const float* dft_tables_allpass, const float* dft_tables_harmonic,
                                 const float* dft_tables_noise, int B, int n_frames, int block,
                                 int n_mag_allpass, int n_mag_harmonic, int n_mag_noise, double sampling_rate,
                                 int round_fp32, float* signal, float* harmonic, float* noise_out, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    if (!workspace) return b2d::fail(B2D_ERR_NULL, "combsub_synth: null workspace");
    const size_t need = b2d_combsub_workspace_bytes(B, n_frames, block, n_mag_allpass, n_mag_harmonic, n_mag_noise);
    if (need == 0) return b2d::fail(B2D_ERR_SHAPE, "combsub_synth: bad shape");
    if (workspace_bytes < need) return b2d::fail(B2D_ERR_WORKSPACE, "combsub_synth: workspace %zu < %zu bytes", workspace_bytes, need);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return b2d::fail(B2D_ERR_ALIGN, "combsub_synth: workspace must be 256-byte aligned");
    if (block % 256 != 0) return b2d::fail(B2D_ERR_UNSUPPORTED, "combsub_synth: block size %d must be a multiple of 256", block);
    const size_t BT = (size_t)B * n_frames * block, BF = (size_t)B * n_frames;
    const int La = 2 * (n_mag_allpass - 1), Lh = 2 * (n_mag_harmonic - 1), Ln = 2 * (n_mag_noise - 1);
    char* ws = static_cast<char*>(workspace);
    const size_t sBT = b2d::align_up(BT * 4, 256);
    float* comb = reinterpret_cast<float*>(ws);
    float* allp = reinterpret_cast<float*>(ws + sBT);
    float* nbuf = noise_out ? noise_out : reinterpret_cast<float*>(ws + 2 * sBT);
    float* ir_ap = reinterpret_cast<float*>(ws + 3 * sBT);
    float* ir_h = reinterpret_cast<float*>(ws + 3 * sBT + b2d::align_up(BF * La * 4, 256));
    float* ir_n = reinterpret_cast<float*>(ws + 3 * sBT + b2d::align_up(BF * La * 4, 256) + b2d::align_up(BF * Lh * 4, 256));
    cudaStream_t st = (cudaStream_t)stream;

    int rc = b2d_comb_source(f0_frames, frame_phase, B, n_frames, block, sampling_rate, round_fp32, comb, stream);
    if (rc) return rc;
    rc = b2d_ir_build(c_group_delay, ctrl_stride, B2D_IR_ALLPASS, nullptr, dft_tables_allpass, B, n_frames,
                      n_mag_allpass, sampling_rate, ir_ap, stream);
    if (rc) return rc;
    rc = b2d_ir_build(c_harmonic, ctrl_stride, B2D_IR_MAG_DYNAMIC, f0_frames, dft_tables_harmonic, B, n_frames,
                      n_mag_harmonic, sampling_rate, ir_h, stream);
    if (rc) return rc;
    rc = b2d_ir_build(c_noise, ctrl_stride, B2D_IR_MAG_HANN, nullptr, dft_tables_noise, B, n_frames, n_mag_noise,
                      sampling_rate, ir_n, stream);
    if (rc) return rc;
    // all-pass on the comb and the noise filter: one launch when the tap counts agree
    if (La == Ln) {
        rc = b2d::ltv_fir_launch(comb, ir_ap, La, allp, noise_in, ir_n, Ln, nbuf, nullptr, nullptr, seed,
                                 utterance_offset, B, n_frames, block, st);
        if (rc) return rc;
    } else {
        rc = b2d::ltv_fir_launch(comb, ir_ap, La, allp, nullptr, nullptr, 0, nullptr, nullptr, nullptr, seed,
                                 utterance_offset, B, n_frames, block, st);
        if (rc) return rc;
        rc = b2d::ltv_fir_launch(noise_in, ir_n, Ln, nbuf, nullptr, nullptr, 0, nullptr, nullptr, nullptr, seed,
                                 utterance_offset, B, n_frames, block, st);
        if (rc) return rc;
    }
    // harmonic magnitude filter on the all-passed comb; signal = harmonic + noise
    return b2d::ltv_fir_launch(allp, ir_h, Lh, harmonic, nullptr, nullptr, 0, nullptr, nbuf, signal, seed,
                               utterance_offset, B, n_frames, block, st);
}
