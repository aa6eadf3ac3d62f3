// Fused tail of one SGD step, executed by the SAME kernel that produced the per-CTA partial gradients:
//   barrier (G CTAs of a policy) -> fixed-order reduction of the partials, each CTA owning a contiguous slice of the
//   parameters (+ its share of ||g||^2) -> barrier -> tf.clip_by_global_norm + TF1 Adam on the slice, packed weight
//   images kept in step -> the last CTA of the grid advances beta powers / step counter and re-arms the barriers.
// It replaces the separate grad_reduce and clip_adam launches of the single-GPU path (two launch gaps per step and a
// redundant full-gradient read per CTA).  All CTAs of the grid must be co-resident: the callers launch at most one
// CTA per SM; every spin is bounded and reports through *status instead of hanging.
// Arithmetic and summation order are identical to grad_reduce_kernel + clip_adam_kernel (bit-identical results).
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "fcnet_layout.cuh"
#include "fcnet_tc_layout.cuh"

namespace ddrl {

struct SgdTail {          // all device pointers; theta == nullptr disables the tail
    float *theta, *m, *v, *beta_pow, *grad, *gnorm_out, *img;
    unsigned char* tc_img;
    double* step_stats;
    int32_t* step_ctr;
    unsigned int* bar;    // [4*P + 4] zero-initialised counters: per policy {A, B}, then the done ticket
    float* sq;            // [P][G] partial sums of squares
    float lr, beta1, beta2, eps, clip;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one thread: arrive and wait until `target` CTAs have arrived; false on timeout
__device__ __forceinline__ bool grid_group_barrier(unsigned int* ctr, unsigned int target) {
    __threadfence();
    atomicAdd(ctr, 1u);
    for (unsigned int i = 0; i < 40000000u; ++i) {
        if (ld_acquire_u32(ctr) >= target) return true;
        __nanosleep(40);
    }
    return false;
}

// Called by every thread of every CTA after the partial gradient (and stat partial) of this CTA has been written.
// smem_red: >= 40 floats of shared memory.  Returns false if a barrier timed out.
__device__ __forceinline__ bool sgd_step_tail(const SgdTail& t, const float* __restrict__ grad_part,
                                              const double* __restrict__ stat_part, int p, int P, int bx, int G, int NP,
                                              int step, int D, int A, float* smem_red) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const int NPs = (NP + 3) & ~3;
    bool ok = true;
    __syncthreads();
    if (tid == 0) smem_red[39] = grid_group_barrier(t.bar + 4 * p, (unsigned)G) ? 1.f : 0.f;
    __syncthreads();
    ok = ok && smem_red[39] != 0.f;
    // ---- reduce this CTA's slice of the parameters over the G partials (fixed order), partial ||g||^2 ----------------
    const int S = (NP + G - 1) / G, j0 = bx * S, j1 = min(NP, j0 + S);
    float ss = 0.f;
    for (int j = j0 + tid; j < j1; j += nt) {
        const float* g = grad_part + (int64_t)p * G * NPs + j;
        float s = 0.f;
#pragma unroll 4
        for (int i = 0; i < G; ++i) s += __ldcg(g + (int64_t)i * NPs);
        t.grad[(int64_t)p * NP + j] = s;
        ss = fmaf(s, s, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) smem_red[warp] = ss;
    if (bx == 0 && tid < DDRL_NSTAT && stat_part && t.step_stats) {
        double s = 0.0;
        for (int i = 0; i < G; ++i) s += __ldcg(stat_part + ((int64_t)p * G + i) * DDRL_NSTAT + tid);
        t.step_stats[((int64_t)step * P + p) * DDRL_NSTAT + tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += smem_red[w];
        t.sq[p * G + bx] = s;
        smem_red[39] = grid_group_barrier(t.bar + 4 * p + 1, (unsigned)G) ? 1.f : 0.f;
    }
    __syncthreads();
    ok = ok && smem_red[39] != 0.f;
    // ---- global norm (every CTA, same fixed order), clip, TF1 Adam on the slice -------------------------------------
    if (warp == 0) {
        float s = 0.f;
        for (int i = lane; i < G; i += 32) s += __ldcg(t.sq + p * G + i);
        s = warp_sum(s);
        if (lane == 0) {
            const float norm = sqrtf(s);
            smem_red[32] = t.clip > 0.f ? t.clip * fminf(1.f / norm, 1.f / t.clip) : 1.f;
            if (bx == 0 && t.gnorm_out) t.gnorm_out[p] = norm;
        }
    }
    __syncthreads();
    const float scale = smem_red[32];
    const float b1p = __ldcg(t.beta_pow + p * 2), b2p = __ldcg(t.beta_pow + p * 2 + 1);
    const float alpha = t.lr * sqrtf(1.f - b2p) / (1.f - b1p);
    for (int j = j0 + tid; j < j1; j += nt) {
        const int64_t k = (int64_t)p * NP + j;
        const float gj = t.grad[k] * scale;
        float mj = t.m[k], vj = t.v[k];
        mj += (gj - mj) * (1.f - t.beta1);
        vj += (gj * gj - vj) * (1.f - t.beta2);
        t.m[k] = mj;
        t.v[k] = vj;
        const float tnew = t.theta[k] - (mj * alpha) / (sqrtf(vj) + t.eps);
        t.theta[k] = tnew;
        if (t.img) {
            const FcSmem L = fc_smem(D, A, false);
            const FcOffsets o = fc_offsets(D, A);
            int p0, p1;
            fc_img_pos(L, o, D, A, j, p0, p1);
            float* im = t.img + (int64_t)p * L.x;
            im[p0] = tnew;
            if (p1 >= 0) im[p1] = tnew;
        }
        if (t.tc_img) {
            const TcImg L = tc_img(D, A);
            const FcOffsets o = fc_offsets(D, A);
            bool f16;
            int p0, p1;
            tc_img_pos(L, o, D, A, j, f16, p0, p1);
            unsigned char* im = t.tc_img + (int64_t)p * L.bytes;
            if (f16) {
                const float ws = tnew * 256.f;
                const __half hi = __float2half_rn(ws);
                *reinterpret_cast<__half*>(im + p0) = hi;
                *reinterpret_cast<__half*>(im + p1) = __float2half_rn(ws - __half2float(hi));
            } else {
                *reinterpret_cast<float*>(im + p0) = tnew;
            }
        }
    }
    // ---- done ticket: the last CTA of the grid advances the optimizer clocks and re-arms the barriers ---------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(t.bar + 4 * P, 1u) == total - 1) {
            for (int q = 0; q < P; ++q) {
                t.beta_pow[q * 2] *= t.beta1;
                t.beta_pow[q * 2 + 1] *= t.beta2;
                t.bar[4 * q] = 0u;
                t.bar[4 * q + 1] = 0u;
            }
            if (t.step_ctr) *t.step_ctr += 1;
            t.bar[4 * P] = 0u;
            __threadfence();
        }
    }
    return ok;
}

}  // namespace ddrl
