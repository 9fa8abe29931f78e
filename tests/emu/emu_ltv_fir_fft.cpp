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
