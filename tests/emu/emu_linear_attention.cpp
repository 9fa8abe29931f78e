// CPU execution of csrc/linear_attention.cu's kernel source (see host_emu.h).  Built by tests/test_emu_linear_attention.py.
#define B2D_HOST_EMU 1
#include "host_emu.h"
#include "../../ddsp_svc_b200/csrc/linear_attention.cu"

namespace { alignas(16) unsigned char smem_raw[1 << 17]; }   // the kernel's `extern __shared__` array

extern "C" int emu_linear_attention(const float* qf, const float* kf, const float* v, float* out, int B, int H, int T, int J,
                                    float eps) {
    static_assert(kLaSmemFloats * sizeof(float) <= sizeof(smem_raw), "shared-memory emulation buffer too small");
    if (J > kLaJmax) return -4;
    LinAttnParams p;
    p.qf = qf; p.kf = kf; p.v = v; p.out = out; p.T = T; p.J = J; p.H = H; p.eps = eps;
    emu::launch((unsigned)(B * H), 1u, kLaThreads, [&] { u2c_linear_attention_kernel(p); });
    return 0;
}
