// ThreadSanitizer driver for the emulated kernels (built by tests/test_emu_tsan.py with -fsanitize=thread).
// A CUDA shared-memory race (missing / misplaced __syncthreads) is a data race between the std::threads of host_emu.h,
// which TSan reports; `racy` is the negative control proving that it does.
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#if defined(TSAN_FIRFFT)
#include "emu_ltv_fir_fft.cpp"
#elif defined(TSAN_CSFAST)
#include "emu_combsubfast.cpp"
#elif defined(TSAN_SUPERFAST)
#include "emu_superfast.cpp"
#elif defined(TSAN_LINATTN)
#include "emu_linear_attention.cpp"
#else
#define B2D_HOST_EMU 1
#include "host_emu.h"
namespace {
float shared_buf[128];
void racy_kernel(float* out, bool with_barrier) {
    const int tid = threadIdx.x;
    shared_buf[tid] = (float)tid;
    if (with_barrier) __syncthreads();
    out[tid] = shared_buf[(tid + 1) & 127];      // reads the neighbour's slot
}
}  // namespace
#endif

int main(int argc, char** argv) {
    std::mt19937 rng(1);
    std::normal_distribution<float> nd(0.f, 1.f);
    auto fill = [&](std::vector<float>& v, float scale, float shift = 0.f) { for (auto& e : v) e = nd(rng) * scale + shift; };
    double s = 0;
#if defined(TSAN_FIRFFT)
    const int B = 1, nF = 7, L = 510, T = nF * 512;
    std::vector<float> x1(B * T), x2(B * T), ir1(B * nF * L), ir2(B * nF * L), y1(B * T), y2(B * T), mix(B * T), add(B * T);
    fill(x1, 1.f); fill(x2, 1.f); fill(ir1, 0.05f); fill(ir2, 0.05f); fill(add, 1.f);
    emu_ltv_fir_fft(x1.data(), ir1.data(), L, y1.data(), x2.data(), ir2.data(), L, y2.data(), add.data(), mix.data(), 1, 0, B, nF, 4);
    emu_ltv_fir_fft(x1.data(), ir1.data(), L, y1.data(), nullptr, ir2.data(), 254, y2.data(), nullptr, mix.data(), 1, 0, B, nF, 32);
    emu_ltv_fir_fft(x1.data(), ir1.data(), L, y1.data(), nullptr, nullptr, 0, nullptr, nullptr, mix.data(), 1, 0, B, nF, 2);
    {   // 2048-point instance (1022 taps)
        std::vector<float> irl(B * nF * 1022);
        fill(irl, 0.03f);
        emu_ltv_fir_fft(x1.data(), irl.data(), 1022, y1.data(), nullptr, nullptr, 0, nullptr, nullptr, mix.data(), 1, 0, B, nF, 4);
    }
    for (float v : mix) s += v;
#elif defined(TSAN_CSFAST)
    const int B = 1, nF = 6, T = nF * 512, C = 3 * 513;
    std::vector<float> comb(B * T), noise(B * T), dense(B * nF * C), out(B * T);
    fill(comb, 1.f); fill(noise, 1.f); fill(dense, 0.3f, -1.f);
    emu_combsubfast(comb.data(), dense.data(), dense.data() + 513, dense.data() + 1026, C, noise.data(), 1, 0, B, nF, 4, out.data());
    emu_combsubfast(comb.data(), dense.data(), dense.data() + 513, dense.data() + 1026, C, nullptr, 1, 0, B, nF, 32, out.data());
    for (float v : out) s += v;
#elif defined(TSAN_LINATTN)
    const int B = 1, H = 2, T = 37, J = 266;
    std::vector<float> qf(B * H * T * J), kf(B * H * T * J), v(B * H * T * 64), out(B * T * H * 64);
    fill(qf, 0.1f, 0.5f); fill(kf, 0.1f, 0.5f); fill(v, 1.f);
    emu_linear_attention(qf.data(), kf.data(), v.data(), out.data(), B, H, T, J, 1e-8f);
    for (float e : out) s += e;
#elif defined(TSAN_SUPERFAST)
    const int B = 1, nF = 9, T = nF * 512, C = 4 * 1025;
    std::vector<float> par(B * nF * 4), noise(B * T), dense(B * nF * C), out(B * T);
    for (int k = 0; k < nF; ++k) { par[4 * k] = 0.005f + 0.0001f * k; par[4 * k + 1] = k + 1 < nF ? 0.0001f : 0.f; par[4 * k + 2] = 0.1f * k - (int)(0.1f * k); par[4 * k + 3] = 0.f; }
    fill(noise, 1.f); fill(dense, 0.3f, -1.f);
    emu_superfast(par.data(), dense.data(), dense.data() + 1025, dense.data() + 2050, dense.data() + 3075, C, noise.data(), 1, 0, B, nF, 5, out.data());
    emu_superfast(par.data(), dense.data(), dense.data() + 1025, dense.data() + 2050, dense.data() + 3075, C, nullptr, 1, 0, B, nF, 29, out.data());
    for (float v : out) s += v;
#else
    const bool with_barrier = argc > 1 && std::string(argv[1]) == "ok";
    std::vector<float> out(128);
    emu::launch(1, 1, 128, [&] { racy_kernel(out.data(), with_barrier); });
    for (float v : out) s += v;
#endif
    std::printf("done %g\n", s);
    return 0;
}
