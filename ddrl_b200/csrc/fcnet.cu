// FCNet (models/fcnet_glorot_uniform_init.py) grouped per-leg forward and the fused
// forward + PPO-loss + backward minibatch kernel.  FP32 FMA path ("parity mode": accurate tanhf/expf,
// fixed summation order => bit-reproducible run to run).
//
// Tiling (both kernels): a CTA of 256 threads owns the weights of ONE policy in shared memory and
// walks over 64-row tiles.  The policy and value branches have identical shapes, so they are run as
// one 128-wide network: layer 1 is x[64,D] * [W1|Wv1][D,128], layer 2 is block diagonal.  Every
// thread owns a 4-row x 8-column register tile (ty = tid/16 -> rows, tx = tid%16 -> columns;
// tx < 8 is the policy branch), operands come from shared memory as 128-bit loads
// (32 FMA per 3 LDS.128).  Activations stay in shared memory (row-major, stride 132 floats) and are
// overwritten in place by their gradients on the way back; weight gradients accumulate in registers
// across all tiles of the CTA and leave as one per-CTA partial (reduced in fixed order by
// ddrl_grad_reduce => deterministic).
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "ppo_loss.cuh"

namespace ddrl {

constexpr int H = DDRL_HIDDEN;   // 64
constexpr int HC = 2 * H;        // 128: policy | value concatenated
constexpr int TM = 64;           // rows per tile
constexpr int NT = 256;          // threads per CTA
constexpr int LDH = HC + 4;      // activation row stride (pad keeps 128-bit row reads conflict free)
constexpr int LDT = HC + 4;      // stride of the transposed layer-2 weights
constexpr int LDD = 20;          // stride of the per-row head outputs / head gradients (2A+1 <= 17)
constexpr int MAXHEAD = (H * (2 * DDRL_MAX_ACT + 1) + NT - 1) / NT;  // 5 head-gradient items / thread

struct FcSmem {
    int W1c, b1c, W2c, b2c, W2Tc, Wo, bo, Wvo, bvo, x, h1, h2, out, dl, red, norm, total;
};

__host__ __device__ inline FcSmem fc_smem(int D, int A, bool train, bool has_norm) {
    const int Dp = (D + 3) & ~3;
    FcSmem s;
    int p = 0;
    s.W1c = p;  p += Dp * HC;
    s.b1c = p;  p += HC;
    s.W2c = p;  p += H * HC;
    s.b2c = p;  p += HC;
    s.W2Tc = p; p += train ? H * LDT : 0;
    s.Wo = p;   p += H * 2 * A;
    s.bo = p;   p += ((2 * A + 3) & ~3);
    s.Wvo = p;  p += H;
    s.bvo = p;  p += 4;
    s.x = p;    p += TM * Dp;
    s.h1 = p;   p += TM * LDH;
    s.h2 = p;   p += TM * LDH;
    s.out = p;  p += TM * LDD;
    s.dl = p;   p += TM * LDD;
    p = (p + 1) & ~1;
    s.red = p;  p += 2 * DDRL_NSTAT;                       // doubles
    s.norm = p; p += has_norm ? 2 * 2 * ((D + 1) & ~1) : 0;  // doubles: mean[D], inv[D]
    s.total = p;
    return s;
}

// ---- weights: global flat (checkpoint order) -> shared (concatenated / transposed) --------------
__device__ void load_weights(float* sm, const FcSmem& L, const float* __restrict__ th, int D, int A, bool train) {
    const FcOffsets o = fc_offsets(D, A);
    const int Dp = (D + 3) & ~3;
    const int tid = threadIdx.x;
    for (int i = tid; i < Dp * H; i += NT) {
        const int k = i >> 6, c = i & 63;
        const bool in = k < D;
        sm[L.W1c + k * HC + c] = in ? th[o.W1 + i] : 0.f;
        sm[L.W1c + k * HC + H + c] = in ? th[o.Wv1 + i] : 0.f;
    }
    for (int i = tid; i < H * H; i += NT) {
        const int k = i >> 6, c = i & 63;
        const float w = th[o.W2 + i], wv = th[o.Wv2 + i];
        sm[L.W2c + k * HC + c] = w;
        sm[L.W2c + k * HC + H + c] = wv;
        if (train) {  // W2T[c][k]: row = output index, column = input index
            sm[L.W2Tc + c * LDT + k] = w;
            sm[L.W2Tc + c * LDT + H + k] = wv;
        }
    }
    for (int i = tid; i < H; i += NT) {
        sm[L.b1c + i] = th[o.b1 + i];
        sm[L.b1c + H + i] = th[o.bv1 + i];
        sm[L.b2c + i] = th[o.b2 + i];
        sm[L.b2c + H + i] = th[o.bv2 + i];
        sm[L.Wvo + i] = th[o.Wvo + i];
    }
    for (int i = tid; i < H * 2 * A; i += NT) sm[L.Wo + i] = th[o.Wo + i];
    if (tid < 2 * A) sm[L.bo + tid] = th[o.bo + tid];
    if (tid == 0) sm[L.bvo] = th[o.bvo];
}

// ---- register-tile micro kernels -------------------------------------------------------------
// acc[i][j] += sum_k A[i*lda + k] * B[k*ldb + j],  k in [0,K), K % 4 == 0.
__device__ __forceinline__ void mm_nn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                      int K, float (&acc)[4][8]) {
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + i * lda + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 b0 = *reinterpret_cast<const float4*>(B + (k + kk) * ldb);
            const float4 b1 = *reinterpret_cast<const float4*>(B + (k + kk) * ldb + 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, b0.x, acc[i][0]);
                acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av, b0.z, acc[i][2]);
                acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                acc[i][4] = fmaf(av, b1.x, acc[i][4]);
                acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                acc[i][6] = fmaf(av, b1.z, acc[i][6]);
                acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
        }
    }
}

// acc[i][j] += sum_r A[r*lda + i] * B[r*ldb + j],  r = r0, r0+rs, ... < r1   (weight gradients).
__device__ __forceinline__ void mm_tn(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                                      int r0, int r1, int rs, float (&acc)[4][8]) {
#pragma unroll 4
    for (int r = r0; r < r1; r += rs) {
        const float4 a = *reinterpret_cast<const float4*>(A + r * lda);
        const float4 b0 = *reinterpret_cast<const float4*>(B + r * ldb);
        const float4 b1 = *reinterpret_cast<const float4*>(B + r * ldb + 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i][0] = fmaf(av[i], b0.x, acc[i][0]);
            acc[i][1] = fmaf(av[i], b0.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b0.z, acc[i][2]);
            acc[i][3] = fmaf(av[i], b0.w, acc[i][3]);
            acc[i][4] = fmaf(av[i], b1.x, acc[i][4]);
            acc[i][5] = fmaf(av[i], b1.y, acc[i][5]);
            acc[i][6] = fmaf(av[i], b1.z, acc[i][6]);
            acc[i][7] = fmaf(av[i], b1.w, acc[i][7]);
        }
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// ---- forward of one tile: x (smem) -> h1, h2 (smem) -> out[r][0..2A] = logits, out[r][2A] = value --
__device__ __forceinline__ void forward_tile(float* sm, const FcSmem& L, int D, int A, int nrows) {
    const int Dp = (D + 3) & ~3;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const bool live = ty * 4 < nrows;  // rows of this thread hold data (8-row granularity per warp)
    float acc[4][8];
    if (live) {
        zero_acc(acc);
        mm_nn(sm + L.x + ty * 4 * Dp, Dp, sm + L.W1c + tx * 8, HC, Dp, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o0, o1;
            const float* b = sm + L.b1c + tx * 8;
            o0.x = tanhf(acc[i][0] + b[0]); o0.y = tanhf(acc[i][1] + b[1]);
            o0.z = tanhf(acc[i][2] + b[2]); o0.w = tanhf(acc[i][3] + b[3]);
            o1.x = tanhf(acc[i][4] + b[4]); o1.y = tanhf(acc[i][5] + b[5]);
            o1.z = tanhf(acc[i][6] + b[6]); o1.w = tanhf(acc[i][7] + b[7]);
            float* dst = sm + L.h1 + (ty * 4 + i) * LDH + tx * 8;
            *reinterpret_cast<float4*>(dst) = o0;
            *reinterpret_cast<float4*>(dst + 4) = o1;
        }
    }
    __syncthreads();
    if (live) {
        zero_acc(acc);
        const int br = (tx >> 3) * H;  // branch column offset in h1
        mm_nn(sm + L.h1 + ty * 4 * LDH + br, LDH, sm + L.W2c + tx * 8, HC, H, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o0, o1;
            const float* b = sm + L.b2c + tx * 8;
            o0.x = tanhf(acc[i][0] + b[0]); o0.y = tanhf(acc[i][1] + b[1]);
            o0.z = tanhf(acc[i][2] + b[2]); o0.w = tanhf(acc[i][3] + b[3]);
            o1.x = tanhf(acc[i][4] + b[4]); o1.y = tanhf(acc[i][5] + b[5]);
            o1.z = tanhf(acc[i][6] + b[6]); o1.w = tanhf(acc[i][7] + b[7]);
            float* dst = sm + L.h2 + (ty * 4 + i) * LDH + tx * 8;
            *reinterpret_cast<float4*>(dst) = o0;
            *reinterpret_cast<float4*>(dst + 4) = o1;
        }
    }
    __syncthreads();
    // heads: item = (row r, output o); o == 2A is the value head reading the value branch of h2.
    const int A2 = 2 * A;
    for (int item = tid; item < TM * (A2 + 1); item += NT) {
        const int r = item & (TM - 1), o = item >> 6;
        if (r >= nrows) continue;
        const bool isv = o == A2;
        const float* hrow = sm + L.h2 + r * LDH + (isv ? H : 0);
        const float* w = isv ? sm + L.Wvo : sm + L.Wo + o;
        const int ws = isv ? 1 : A2;
        float s = 0.f;
#pragma unroll 4
        for (int k = 0; k < H; k += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hrow + k);
            s = fmaf(hv.x, w[(k + 0) * ws], s);
            s = fmaf(hv.y, w[(k + 1) * ws], s);
            s = fmaf(hv.z, w[(k + 2) * ws], s);
            s = fmaf(hv.w, w[(k + 3) * ws], s);
        }
        sm[L.out + r * LDD + o] = s + (isv ? sm[L.bvo] : sm[L.bo + o]);
    }
    __syncthreads();
}

// ---- x tile: global -> shared, optional MeanStdFilter normalisation ---------------------------------
__device__ __forceinline__ void load_x_tile(float* sm, const FcSmem& L, const float* __restrict__ obs,
                                            float* __restrict__ obs_out, const double* snorm, float clip,
                                            int64_t row0, int nrows, int D) {
    const int Dp = (D + 3) & ~3;
    const int n = nrows * D;
    const float* src = obs + row0 * D;
    for (int i = threadIdx.x; i < n; i += NT) {
        const int r = i / D, d = i - r * D;
        float v = src[i];
        if (snorm) {
            v = (float)(((double)v - snorm[d]) * snorm[((D + 1) & ~1) + d]);
            if (clip > 0.f) v = fminf(fmaxf(v, -clip), clip);
        }
        if (obs_out) obs_out[row0 * D + i] = v;
        sm[L.x + r * Dp + d] = v;
    }
}

// =================================================================================================
// K1: grouped forward (+filter-normalise prologue, +DiagGaussian sample/logp epilogue)
// =================================================================================================
__global__ void __launch_bounds__(NT, 1)
fcnet_forward_kernel(const float* __restrict__ theta, const float* __restrict__ obs, const double* __restrict__ norm,
                     float clip, int64_t R, int D, int A, float* __restrict__ obs_out, float* __restrict__ logits,
                     float* __restrict__ value, const float* __restrict__ eps, float* __restrict__ action,
                     float* __restrict__ logp) {
    extern __shared__ __align__(16) float sm[];
    const int p = blockIdx.y;
    const FcSmem L = fc_smem(D, A, false, norm != nullptr);
    const FcOffsets o = fc_offsets(D, A);
    const int tid = threadIdx.x;
    const int A2 = 2 * A;
    for (int i = tid; i < TM * ((D + 3) & ~3); i += NT) sm[L.x + i] = 0.f;
    load_weights(sm, L, theta + (int64_t)p * o.NP, D, A, false);
    double* snorm = nullptr;
    if (norm) {
        snorm = reinterpret_cast<double*>(sm + L.norm);
        const int Dd = (D + 1) & ~1;
        for (int i = tid; i < D; i += NT) {
            snorm[i] = norm[(int64_t)p * 2 * D + i];
            snorm[Dd + i] = norm[(int64_t)p * 2 * D + D + i];
        }
    }
    __syncthreads();
    const float* obs_p = obs + (int64_t)p * R * D;
    float* obs_out_p = obs_out ? obs_out + (int64_t)p * R * D : nullptr;
    const int64_t ntiles = (R + TM - 1) / TM;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * TM;
        const int nrows = (int)min((int64_t)TM, R - row0);
        load_x_tile(sm, L, obs_p, obs_out_p, snorm, clip, row0, nrows, D);
        __syncthreads();
        forward_tile(sm, L, D, A, nrows);
        // epilogue ---------------------------------------------------------------------------------
        if (eps && tid < nrows) {
            const int r = tid;
            const int64_t gr = (int64_t)p * R + row0 + r;
            float sz2 = 0.f, sls = 0.f;
            for (int i = 0; i < A; ++i) {
                const float mu = sm[L.out + r * LDD + i], ls = sm[L.out + r * LDD + A + i];
                const float sd = expf(ls);
                const float a = mu + sd * eps[gr * A + i];
                const float z = (a - mu) / sd;
                sz2 = fmaf(z, z, sz2);
                sls += ls;
                sm[L.dl + r * LDD + i] = a;
            }
            logp[gr] = -0.5f * sz2 - 0.5f * kLog2Pi * (float)A - sls;
        }
        if (eps) __syncthreads();
        if (logits)
            for (int i = tid; i < nrows * A2; i += NT) {
                const int r = i / A2, c = i - r * A2;
                logits[((int64_t)p * R + row0) * A2 + i] = sm[L.out + r * LDD + c];
            }
        if (value && tid < nrows) value[(int64_t)p * R + row0 + tid] = sm[L.out + tid * LDD + A2];
        if (eps)
            for (int i = tid; i < nrows * A; i += NT) {
                const int r = i / A, c = i - r * A;
                action[((int64_t)p * R + row0) * A + i] = sm[L.dl + r * LDD + c];
            }
        __syncthreads();
    }
}

// =================================================================================================
// K2/K6: fused forward + PPO loss + backward for one minibatch (all policies), per-CTA partial grads
// =================================================================================================
struct TrainArgs {
    const float *theta, *obs, *actions, *old_logits, *old_logp, *vf_preds, *adv, *vtarg, *ext_dlogits, *ext_dvalue;
    int64_t R;
    int D, A, MB;
    const int32_t* mb_perm;
    int64_t perm_stride;
    const int32_t* step_ctr;
    const float* kl_coeff;
    ddrl_ppo_hyper hp;
    float* grad_part;
    double* stat_part;
};

__global__ void __launch_bounds__(NT, 1) fcnet_train_kernel(const TrainArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int p = blockIdx.y, G = gridDim.x, bx = blockIdx.x;
    const int D = a.D, A = a.A, A2 = 2 * A, Dp = (D + 3) & ~3;
    const FcSmem L = fc_smem(D, A, true, false);
    const FcOffsets o = fc_offsets(D, A);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31, warp = tid >> 5;
    const bool ext = a.ext_dlogits != nullptr;

    // minibatch and this CTA's row range --------------------------------------------------------
    const int step = a.step_ctr ? *a.step_ctr : 0;
    const int mb = a.mb_perm ? a.mb_perm[(int64_t)p * a.perm_stride + step] : step;
    const int64_t mb0 = (int64_t)mb * a.MB;
    const int64_t mb1 = min(mb0 + a.MB, a.R);
    const int rpc = (((a.MB + G - 1) / G) + 7) & ~7;
    const int64_t cr0 = min(mb0 + (int64_t)bx * rpc, mb1), cr1 = min(cr0 + rpc, mb1);

    for (int i = tid; i < TM * Dp; i += NT) sm[L.x + i] = 0.f;
    if (cr1 > cr0) load_weights(sm, L, a.theta + (int64_t)p * o.NP, D, A, true);
    __syncthreads();

    // gradient accumulators (registers, live across tiles) -----------------------------------
    float gW2[4][8], gW1[4][8], gHead[MAXHEAD];
    float gb1 = 0.f, gb2 = 0.f, gbo = 0.f;
    zero_acc(gW2);
    zero_acc(gW1);
#pragma unroll
    for (int i = 0; i < MAXHEAD; ++i) gHead[i] = 0.f;
    double st[DDRL_NSTAT];
#pragma unroll
    for (int i = 0; i < DDRL_NSTAT; ++i) st[i] = 0.0;

    // dW1 row-split: ngd d-groups of 4 input features; the 16 ty-groups are shared by nsplit replicas
    const int ngd = Dp >> 2;
    const int nsplit = max(1, 16 / ngd);
    const int rsplit = ty / ngd, dq = ty - rsplit * ngd;
    const bool w1_live = rsplit < nsplit;

    const float klc = ext ? 0.f : a.kl_coeff[p];
    const float* obs_p = a.obs + (int64_t)p * a.R * D;

    for (int64_t row0 = cr0; row0 < cr1; row0 += TM) {
        const int nrows = (int)min((int64_t)TM, cr1 - row0);
        load_x_tile(sm, L, obs_p, nullptr, nullptr, 0.f, row0, nrows, D);
        __syncthreads();
        forward_tile(sm, L, D, A, nrows);

        // ---- per-row loss gradient -> dl[r][0..2A) = dL/dlogits, dl[r][2A] = dL/dvalue ------------
        if (tid < TM) {
            const int r = tid;
            double s[DDRL_NSTAT];
#pragma unroll
            for (int i = 0; i < DDRL_NSTAT; ++i) s[i] = 0.0;
            if (r < nrows) {
                const int64_t gr = (int64_t)p * a.R + row0 + r;
                float* dl = sm + L.dl + r * LDD;
                const float* out = sm + L.out + r * LDD;
                if (ext) {
                    for (int i = 0; i < A2; ++i) dl[i] = a.ext_dlogits[gr * A2 + i];
                    dl[A2] = a.ext_dvalue[gr];
                } else {
                    ppo_row_loss(out, A, a.actions + gr * A, a.old_logits + gr * A2, a.old_logp[gr], a.vf_preds[gr],
                                 a.adv[gr], a.vtarg[gr], klc, a.hp, dl, s);
                }
            }
            if (!ext) {
#pragma unroll
                for (int i = 0; i < DDRL_NSTAT; ++i) s[i] = warp_sum(s[i]);
                double* red = reinterpret_cast<double*>(sm + L.red);
                if (warp == 1 && lane == 0)
#pragma unroll
                    for (int i = 0; i < DDRL_NSTAT; ++i) red[i] = s[i];
                if (tid == 0)
#pragma unroll
                    for (int i = 0; i < DDRL_NSTAT; ++i) st[i] += s[i];
            }
        }
        __syncthreads();
        if (tid == 0 && !ext) {
            const double* red = reinterpret_cast<const double*>(sm + L.red);
#pragma unroll
            for (int i = 0; i < DDRL_NSTAT; ++i) st[i] += red[i];
        }

        // ---- B1: head weight gradients  gWo[k][o] += sum_r h2[r][k] dl[r][o] ---------------------------
#pragma unroll
        for (int i = 0; i < MAXHEAD; ++i) {
            const int item = tid + i * NT;
            if (item < H * (A2 + 1)) {
                const int k = item & 63, oo = item >> 6;
                const float* hcol = sm + L.h2 + (oo == A2 ? H : 0) + k;
                const float* dcol = sm + L.dl + oo;
                float s = gHead[i];
                for (int r = 0; r < nrows; ++r) s = fmaf(hcol[r * LDH], dcol[r * LDD], s);
                gHead[i] = s;
            }
        }
        if (tid <= A2) {
            float s = gbo;
            for (int r = 0; r < nrows; ++r) s += sm[L.dl + r * LDD + tid];
            gbo = s;
        }
        __syncthreads();

        // ---- B2: dz2 = (dl . Wo^T) * (1 - h2^2), in place over h2 -------------------------------------
        if (ty * 4 < nrows) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty * 4 + i;
                const float* dl = sm + L.dl + r * LDD;
                float* hrow = sm + L.h2 + r * LDH + tx * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = tx * 8 + j, k = c & 63;
                    float s = 0.f;
                    if (tx < 8) {
                        const float* w = sm + L.Wo + k * A2;
                        for (int q = 0; q < A2; ++q) s = fmaf(dl[q], w[q], s);
                    } else {
                        s = dl[A2] * sm[L.Wvo + k];
                    }
                    const float h = hrow[j];
                    hrow[j] = s * (1.f - h * h);
                }
            }
        }
        __syncthreads();

        // ---- B3: gW2[k][c] += sum_r h1[r][br+k] dz2[r][c];  gb2[c] += sum_r dz2[r][c] -------------------
        mm_tn(sm + L.h1 + (tx >> 3) * H + ty * 4, LDH, sm + L.h2 + tx * 8, LDH, 0, nrows, 1, gW2);
        if (tid < HC) {
            float s = gb2;
            for (int r = 0; r < nrows; ++r) s += sm[L.h2 + r * LDH + tid];
            gb2 = s;
        }
        // ---- B4: dz1 = (dz2 . W2^T) * (1 - h1^2)  (registers), then in place over h1 --------------------
        float dz1[4][8];
        const bool live = ty * 4 < nrows;
        if (live) {
            zero_acc(dz1);
            mm_nn(sm + L.h2 + ty * 4 * LDH + (tx >> 3) * H, LDH, sm + L.W2Tc + tx * 8, LDT, H, dz1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float* hrow = sm + L.h1 + (ty * 4 + i) * LDH + tx * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = hrow[j];
                    dz1[i][j] *= (1.f - h * h);
                }
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float* dst = sm + L.h1 + (ty * 4 + i) * LDH + tx * 8;
                *reinterpret_cast<float4*>(dst) = make_float4(dz1[i][0], dz1[i][1], dz1[i][2], dz1[i][3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(dz1[i][4], dz1[i][5], dz1[i][6], dz1[i][7]);
            }
        }
        __syncthreads();
        // ---- B5: gW1[d][c] += sum_r x[r][d] dz1[r][c];  gb1[c] += sum_r dz1[r][c] ------------------------
        if (w1_live) mm_tn(sm + L.x + dq * 4, Dp, sm + L.h1 + tx * 8, LDH, rsplit, nrows, nsplit, gW1);
        if (tid < HC) {
            float s = gb1;
            for (int r = 0; r < nrows; ++r) s += sm[L.h1 + r * LDH + tid];
            gb1 = s;
        }
        __syncthreads();
    }

    // ---- write the per-CTA partial gradient (flat checkpoint order) -------------------------------------
    float* gp = a.grad_part + ((int64_t)p * G + bx) * o.NP;
    {   // W2 / Wv2
        const int base = (tx < 8 ? o.W2 : o.Wv2) + (tx & 7) * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) gp[base + (ty * 4 + i) * H + j] = gW2[i][j];
    }
    if (tid < HC) {
        gp[(tid < H ? o.b1 : o.bv1) + (tid & 63)] = gb1;
        gp[(tid < H ? o.b2 : o.bv2) + (tid & 63)] = gb2;
    }
#pragma unroll
    for (int i = 0; i < MAXHEAD; ++i) {
        const int item = tid + i * NT;
        if (item < H * (A2 + 1)) {
            const int k = item & 63, oo = item >> 6;
            gp[oo == A2 ? o.Wvo + k : o.Wo + k * A2 + oo] = gHead[i];
        }
    }
    if (tid <= A2) gp[tid == A2 ? o.bvo : o.bo + tid] = gbo;
    // W1 / Wv1: combine the row-split replicas in fixed order through shared memory (aliases h1/h2)
    __syncthreads();
    float* scratch = sm + L.h1;  // nsplit * Dp * HC <= 2 * TM * LDH floats
    if (w1_live) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) scratch[(rsplit * Dp + dq * 4 + i) * HC + tx * 8 + j] = gW1[i][j];
    }
    __syncthreads();
    for (int i = tid; i < D * HC; i += NT) {
        const int d = i >> 7, c = i & 127;
        float s = scratch[i];
        for (int q = 1; q < nsplit; ++q) s += scratch[q * Dp * HC + i];
        gp[(c < H ? o.W1 : o.Wv1) + d * H + (c & 63)] = s;
    }
    if (tid == 0 && a.stat_part) {
        double* sp = a.stat_part + ((int64_t)p * G + bx) * DDRL_NSTAT;
#pragma unroll
        for (int i = 0; i < DDRL_NSTAT; ++i) sp[i] = st[i];
    }
}

}  // namespace ddrl

using namespace ddrl;

static int g_num_sms = 0;
static int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) g_num_sms = 148;
    }
    return g_num_sms;
}

extern "C" int ddrl_fcnet_num_params(int D, int A) {
    if (D < 1 || D > DDRL_MAX_OBS || A < 1 || A > DDRL_MAX_ACT) return DDRL_E_UNSUPPORTED_SHAPE;
    return fc_offsets(D, A).NP;
}

extern "C" int ddrl_fcnet_forward(const float* theta, const float* obs, const double* norm, float clip, int P,
                                  int64_t R, int D, int A, float* obs_out, float* logits, float* value,
                                  const float* eps, float* action, float* logp, void* stream) {
    DDRL_REQUIRE(theta && obs && P >= 1 && R >= 0, DDRL_E_BADARG, "fcnet_forward: null theta/obs or bad P/R");
    DDRL_REQUIRE(D >= 1 && D <= DDRL_MAX_OBS && A >= 1 && A <= DDRL_MAX_ACT, DDRL_E_UNSUPPORTED_SHAPE,
                 "fcnet_forward: unsupported D=%d A=%d (D<=%d, A<=%d, hiddens [64,64], tanh)", D, A, DDRL_MAX_OBS,
                 DDRL_MAX_ACT);
    DDRL_REQUIRE(!eps || (action && logp), DDRL_E_BADARG, "fcnet_forward: eps given without action/logp outputs");
    if (R == 0) return DDRL_OK;
    const FcSmem L = fc_smem(D, A, false, norm != nullptr);
    const size_t smem = (size_t)L.total * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(fcnet_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess) {
            set_error("fcnet_forward: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
            return DDRL_E_CUDA;
        }
        attr_set = true;
    }
    const int64_t ntiles = (R + TM - 1) / TM;
    const int per_policy = (int)std::min<int64_t>(ntiles, std::max(1, num_sms() / P));
    dim3 grid(per_policy, P);
    fcnet_forward_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(theta, obs, norm, clip, R, D, A, obs_out, logits,
                                                                   value, eps, action, logp);
    DDRL_CHECK_LAUNCH("fcnet_forward");
    return DDRL_OK;
}

extern "C" int ddrl_ppo_train_step(const float* theta, const float* obs, const float* actions,
                                   const float* old_logits, const float* old_logp, const float* vf_preds,
                                   const float* adv, const float* vtarg, const float* ext_dlogits,
                                   const float* ext_dvalue, int P, int64_t R, int D, int A, int MB,
                                   const int32_t* mb_perm, int64_t perm_stride, const int32_t* step_ctr,
                                   const float* kl_coeff, const ddrl_ppo_hyper* hyper, int ctas_per_policy,
                                   float* grad_part, double* stat_part, void* stream) {
    DDRL_REQUIRE(theta && obs && grad_part && P >= 1 && R >= 1 && MB >= 1 && ctas_per_policy >= 1, DDRL_E_BADARG,
                 "ppo_train_step: null pointer or bad P/R/MB/ctas");
    DDRL_REQUIRE(D >= 1 && D <= DDRL_MAX_OBS && A >= 1 && A <= DDRL_MAX_ACT, DDRL_E_UNSUPPORTED_SHAPE,
                 "ppo_train_step: unsupported D=%d A=%d", D, A);
    const bool ext = ext_dlogits != nullptr;
    DDRL_REQUIRE(ext ? (ext_dvalue != nullptr)
                     : (actions && old_logits && old_logp && vf_preds && adv && vtarg && kl_coeff && hyper),
                 DDRL_E_BADARG, "ppo_train_step: missing batch arrays for the %s path", ext ? "external-gradient" : "PPO");
    TrainArgs a;
    a.theta = theta; a.obs = obs; a.actions = actions; a.old_logits = old_logits; a.old_logp = old_logp;
    a.vf_preds = vf_preds; a.adv = adv; a.vtarg = vtarg; a.ext_dlogits = ext_dlogits; a.ext_dvalue = ext_dvalue;
    a.R = R; a.D = D; a.A = A; a.MB = MB; a.mb_perm = mb_perm; a.perm_stride = perm_stride; a.step_ctr = step_ctr;
    a.kl_coeff = kl_coeff;
    if (hyper) a.hp = *hyper; else a.hp = ddrl_ppo_hyper{0.f, 0.f, 0.f, 0.f, 1.f};
    a.grad_part = grad_part; a.stat_part = stat_part;
    const FcSmem L = fc_smem(D, A, true, false);
    const size_t smem = (size_t)L.total * sizeof(float);
    DDRL_REQUIRE(smem <= 227 * 1024, DDRL_E_UNSUPPORTED_SHAPE, "ppo_train_step: shared memory %zu > 227 KB", smem);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(fcnet_train_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess) {
            set_error("ppo_train_step: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
            return DDRL_E_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(ctas_per_policy, P);
    fcnet_train_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(a);
    DDRL_CHECK_LAUNCH("ppo_train_step");
    return DDRL_OK;
}
