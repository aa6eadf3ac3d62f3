// Shared helpers for libddrl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ddrl_b200.h"

namespace ddrl {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DDRL_REQUIRE(cond, code, ...)            \
    do {                                          \
        if (!(cond)) {                            \
            ::ddrl::set_error(__VA_ARGS__);       \
            return (code);                        \
        }                                         \
    } while (0)

#define DDRL_CHECK_LAUNCH(name)                                                        \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            ::ddrl::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return DDRL_E_CUDA;                                                        \
        }                                                                              \
        ::ddrl::count_launch();                                                        \
    } while (0)

// FCNet flat-parameter offsets (checkpoint order, include/ddrl_b200.h).
struct FcOffsets {
    int W1, b1, Wv1, bv1, W2, b2, Wv2, bv2, Wo, bo, Wvo, bvo, NP;
};
__host__ __device__ inline FcOffsets fc_offsets(int D, int A) {
    FcOffsets o;
    const int H = DDRL_HIDDEN;
    int p = 0;
    o.W1 = p;  p += D * H;
    o.b1 = p;  p += H;
    o.Wv1 = p; p += D * H;
    o.bv1 = p; p += H;
    o.W2 = p;  p += H * H;
    o.b2 = p;  p += H;
    o.Wv2 = p; p += H * H;
    o.bv2 = p; p += H;
    o.Wo = p;  p += H * 2 * A;
    o.bo = p;  p += 2 * A;
    o.Wvo = p; p += H;
    o.bvo = p; p += 1;
    o.NP = p;
    return o;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ddrl
