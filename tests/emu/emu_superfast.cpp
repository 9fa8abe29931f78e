// CPU execution of csrc/superfast.cu's main kernel source (see host_emu.h).  Built by tests/test_emu_superfast.py.
// The frame scan (warp shuffles) is not emulated: the caller passes frame_par = (s, ds, acc_prev, 0) per frame.
#define B2D_HOST_EMU 1
#include "host_emu.h"
#include "../../ddsp_svc_b200/csrc/superfast.cu"

namespace { alignas(16) unsigned char smem_raw[1 << 17]; }   // the kernel's `extern __shared__` array

extern "C" int emu_superfast(const float* frame_par, const float* hm, const float* hp, const float* nm, const float* np_,
                             long long stride, const float* noise_in, unsigned long long seed, long long utt_off, int B,
                             int nF, int G, float* out) {
    static_assert(kSmemBytes <= sizeof(smem_raw), "shared-memory emulation buffer too small");
    SfParams p;
    p.f0 = nullptr; p.frame_par = reinterpret_cast<const float4*>(frame_par);
    p.c_hm = hm; p.c_hp = hp; p.c_nm = nm; p.c_np = np_; p.ctrl_stride = stride; p.noise_in = noise_in; p.out = out;
    p.nF = nF; p.P = 512; p.G = G; p.seed = seed; p.utt_off = utt_off;
    emu::launch((unsigned)((nF + G - 1) / G), (unsigned)B, kThreads, [&] { superfast_kernel<false>(p); });
    return 0;
}
