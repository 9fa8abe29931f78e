// CPU execution of csrc/ltv_fir_fft.cu's kernel source (see host_emu.h).  Built by tests/test_emu_ltv_fir_fft.py.
#define B2D_HOST_EMU 1
#include "host_emu.h"
#include "../../ddsp_svc_b200/csrc/ltv_fir_fft.cu"

namespace { alignas(16) unsigned char smem_raw[1 << 18]; }   // the kernel's `extern __shared__` array

extern "C" int emu_ltv_fir_fft(const float* x1, const float* ir1, int L1, float* y1, const float* x2, const float* ir2,
                               int L2, float* y2, const float* addend, float* mix, unsigned long long seed,
                               long long utt_off, int B, int nF, int G) {
    static_assert(fir_fft_smem<2048, 2>() <= sizeof(smem_raw), "shared-memory emulation buffer too small");
    FftFirParams p;
    p.job[0] = {x1, ir1, y1, L1};
    p.job[1] = {x2, ir2, y2, ir2 ? L2 : L1};
    p.addend = addend; p.mix = mix; p.seed = seed; p.utt_off = utt_off; p.nF = nF; p.G = G;
    const unsigned gx = (unsigned)((nF + G - 1) / G);
    const int tmax = ir2 ? (L1 > L2 ? L1 : L2) : L1;
    if (tmax > 1024) return -4;
    if (tmax <= 512) {             // same size selection as b2d::ltv_fir_fft_launch
        if (ir2) emu::launch(gx, (unsigned)B, kThreads, [&] { ltv_fir_fft_kernel<1024, 2, false>(p); });
        else emu::launch(gx, (unsigned)B, kThreads, [&] { ltv_fir_fft_kernel<1024, 1, false>(p); });
    } else {
        if (ir2) emu::launch(gx, (unsigned)B, kThreads, [&] { ltv_fir_fft_kernel<2048, 2, false>(p); });
        else emu::launch(gx, (unsigned)B, kThreads, [&] { ltv_fir_fft_kernel<2048, 1, false>(p); });
    }
    return 0;
}

// spectrum path: ir_spectrum_kernel (taps -> packed spectra, once per frame) + the SPEC variant of the FIR kernel
extern "C" int emu_ltv_fir_fft_spec(const float* x1, const float* ir1, int L1, float* y1, const float* x2, const float* ir2,
                                    int L2, float* y2, float* mix, unsigned long long seed, long long utt_off, int B, int nF,
                                    int G, float* spec1, float* spec2) {
    if (L1 > 512 || L2 > 512) return -4;
    IrSpecParams sp;
    sp.ir[0] = ir1; sp.ir[1] = ir2; sp.L[0] = L1; sp.L[1] = L2; sp.nF = nF;
    sp.spec[0] = reinterpret_cast<float2*>(spec1); sp.spec[1] = reinterpret_cast<float2*>(spec2);
    for (unsigned z = 0; z < 2; ++z) {                        // the emulator's grid is 2-D: one launch per job
        IrSpecParams one = sp;
        one.ir[0] = sp.ir[z]; one.spec[0] = sp.spec[z]; one.L[0] = sp.L[z];
        emu::launch((unsigned)((nF + 7) / 8), (unsigned)B, kThreads, [&] { ir_spectrum_kernel<1024, false>(one); });
    }
    FftFirParams p;
    p.job[0] = {x1, spec1, y1, L1};
    p.job[1] = {x2, spec2, y2, L2};
    p.addend = nullptr; p.mix = mix; p.seed = seed; p.utt_off = utt_off; p.nF = nF; p.G = G;
    emu::launch((unsigned)((nF + G - 1) / G), (unsigned)B, kThreads, [&] { ltv_fir_fft_kernel<1024, 2, false, 0, true>(p); });
    return 0;
}
