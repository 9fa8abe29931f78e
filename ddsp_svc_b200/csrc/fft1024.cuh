// 1024-point instance of fft_smem.cuh under the names combsubfast.cu uses.
#pragma once
#include "fft_smem.cuh"

namespace b2d_fft1024 {
using namespace b2d_fft;
using b2d_fft_smem::kThreads;
using b2d_fft_smem::padi;

constexpr int kN = 1024;
constexpr int kPad = b2d_fft_smem::Plan<1024>::kPad;
constexpr int kTw2 = b2d_fft_smem::Plan<1024>::kTw2;
constexpr int kTw3 = b2d_fft_smem::Plan<1024>::kTw3;

template <int NBATCH>
__device__ __forceinline__ void fft1024(float2* buf, const float2* tw2, const float2* tw3, int tid) {
    b2d_fft_smem::fft_forward<1024, NBATCH>(buf, tw2, tw3, tid);
}
__device__ __forceinline__ void init_twiddles(float2* tw2, float2* tw3, int tid) {
    b2d_fft_smem::init_twiddles<1024>(tw2, tw3, tid);
}

}  // namespace b2d_fft1024
