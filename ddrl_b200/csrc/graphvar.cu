// Graph-layer variants of the reference that are present but not wired into a model:
//   MPNN2 (models/gcn.py:96-150), GAT1 (models/gcn.py:153-206), graph_ops.symm_norm (models/graph_ops.py:3-11) and
//   graph_ops.segment_softmax (models/graph_ops.py:23-26).  Forward only, FP32, 4-node leg graphs (N = 4), one CTA of
//   256 threads = (node, unit) pairs per sample, grid-stride over the batch; weights are read through L1.
#include <algorithm>

#include "common.cuh"

namespace ddrl {

constexpr int GV_N = 4;        // nodes of the leg graph
constexpr int GV_MAXF = 64;    // input features per node
constexpr int GV_MAXU = 64;    // units
constexpr int GV_NT = GV_N * GV_MAXU;

__device__ __forceinline__ float gv_act(float v, int act) { return act == 1 ? tanhf(v) : v; }

// MPNN2: e_(s->r) = [x_s, x_r] W_msg;  m_r = mean over incoming edges (0 if none);  y = act([x, m] W_upd + b)
__global__ void __launch_bounds__(GV_NT) mpnn2_forward_kernel(const float* __restrict__ x, const float* __restrict__ adj,
                                                              const float* __restrict__ Wmsg, const float* __restrict__ Wupd,
                                                              const float* __restrict__ bias, int64_t B, int F, int U, int act,
                                                              float* __restrict__ y) {
    __shared__ float xs[GV_N][GV_MAXF];
    __shared__ float ms[GV_N][GV_MAXU];
    __shared__ float as[GV_N][GV_N];
    const int tid = threadIdx.x, n = tid / GV_MAXU, u = tid % GV_MAXU;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < GV_N * F; i += GV_NT) xs[i / F][i % F] = x[b * GV_N * F + i];
        if (tid < GV_N * GV_N) as[tid / GV_N][tid % GV_N] = adj[b * GV_N * GV_N + tid];
        __syncthreads();
        if (u < U) {
            float own = 0.f;                       // receiver half of every incoming message: x_n W_msg[F:2F]
            for (int f = 0; f < F; ++f) own = fmaf(xs[n][f], __ldg(Wmsg + (size_t)(F + f) * U + u), own);
            float sum = 0.f;
            int cnt = 0;
            for (int s = 0; s < GV_N; ++s) {
                if (as[s][n] != 0.f) {             // edge s -> n
                    float e = own;
                    for (int f = 0; f < F; ++f) e = fmaf(xs[s][f], __ldg(Wmsg + (size_t)f * U + u), e);
                    sum += e;
                    ++cnt;
                }
            }
            ms[n][u] = cnt ? sum / (float)cnt : 0.f;
        }
        __syncthreads();
        if (u < U) {
            float v = bias ? bias[u] : 0.f;
            for (int f = 0; f < F; ++f) v = fmaf(xs[n][f], __ldg(Wupd + (size_t)f * U + u), v);
            for (int k = 0; k < U; ++k) v = fmaf(ms[n][k], __ldg(Wupd + (size_t)(F + k) * U + u), v);
            y[(b * GV_N + n) * U + u] = gv_act(v, act);
        }
    }
}

// GAT1: self loops added; x' = x W_pre; a_(s->r) = leaky_relu_0.2(w_att . [x'_s, x'_r]); softmax over the edges of each
// receiver (no max subtraction, like the reference); Att[s][r] scattered by (sender, receiver); y_s = sum_r Att[s][r] x'_r
// (the reference multiplies the [sender][receiver] matrix from the left: row = sender).
__global__ void __launch_bounds__(GV_NT) gat1_forward_kernel(const float* __restrict__ x, const float* __restrict__ adj,
                                                             const float* __restrict__ Wpre, const float* __restrict__ watt,
                                                             const float* __restrict__ bias, int64_t B, int F, int U, int act,
                                                             float* __restrict__ y) {
    __shared__ float xs[GV_N][GV_MAXF];
    __shared__ float xp[GV_N][GV_MAXU];
    __shared__ float as[GV_N][GV_N];
    __shared__ float ps[GV_N], pr[GV_N];          // w_att[:U] . x'_n  and  w_att[U:] . x'_n
    __shared__ float att[GV_N][GV_N];
    const int tid = threadIdx.x, n = tid / GV_MAXU, u = tid % GV_MAXU, lane = tid & 31;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < GV_N * F; i += GV_NT) xs[i / F][i % F] = x[b * GV_N * F + i];
        if (tid < GV_N * GV_N) {
            const int s = tid / GV_N, r = tid % GV_N;
            as[s][r] = fminf(1.f, adj[b * GV_N * GV_N + tid] + (s == r ? 1.f : 0.f));
        }
        __syncthreads();
        float v = 0.f;
        if (u < U)
            for (int f = 0; f < F; ++f) v = fmaf(xs[n][f], __ldg(Wpre + (size_t)f * U + u), v);
        if (u < U) xp[n][u] = v;
        // the two halves of the attention logit per node: node n = warps 2n, 2n+1 (64 units) -> fixed-order two-warp sums
        float a0 = u < U ? v * __ldg(watt + u) : 0.f, a1 = u < U ? v * __ldg(watt + U + u) : 0.f;
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        __shared__ float part[GV_NT / 32][2];
        if (lane == 0) { part[tid >> 5][0] = a0; part[tid >> 5][1] = a1; }
        __syncthreads();
        if (tid < GV_N) {
            ps[tid] = part[2 * tid][0] + part[2 * tid + 1][0];
            pr[tid] = part[2 * tid][1] + part[2 * tid + 1][1];
        }
        __syncthreads();
        if (tid < GV_N) {                          // receiver r = tid: softmax over its incoming edges
            const int r = tid;
            float e[GV_N], sum = 0.f;
            for (int s = 0; s < GV_N; ++s) {
                float a = ps[s] + pr[r];
                a = a > 0.f ? a : 0.2f * a;        // tf.nn.leaky_relu default alpha
                e[s] = as[s][r] != 0.f ? expf(a) : 0.f;
                sum += e[s];
            }
            for (int s = 0; s < GV_N; ++s) att[s][r] = as[s][r] != 0.f ? e[s] / sum : 0.f;
        }
        __syncthreads();
        if (u < U) {
            float o = 0.f;
            for (int r = 0; r < GV_N; ++r) o = fmaf(att[n][r], xp[r][u], o);
            if (bias) o += bias[u];
            y[(b * GV_N + n) * U + u] = gv_act(o, act);
        }
    }
}

// ---- backward passes (reference: tf.gradients through models/gcn.py:96-206) -----------------------------
// One CTA of 256 threads = (node n, unit u) walks its samples: the forward is recomputed, the gradient w.r.t. the inputs is
// written per sample, the weight gradients are accumulated in REGISTERS (thread (q = tid / 64, u) owns the rows r = q mod 4
// of column u) over all samples of the CTA and leave once as a per-CTA partial [G][NPs] (reduce with ddrl_grad_reduce(P = 1));
// fixed order, no atomics.  Flat order of the partial: MPNN2 [W_msg (2F x U) | W_upd ((F + U) x U) | b (U)],
// GAT1 [W_pre (F x U) | w_att (2U) | b (U)].
constexpr int GV_RPT = (2 * GV_MAXF) / 4;      // rows per thread of a (<= 128)-row matrix

__global__ void __launch_bounds__(GV_NT) mpnn2_backward_kernel(const float* __restrict__ x, const float* __restrict__ adj,
                                                               const float* __restrict__ Wmsg, const float* __restrict__ Wupd,
                                                               const float* __restrict__ bias, const float* __restrict__ dy,
                                                               int64_t B, int F, int U, int act, float* __restrict__ dx,
                                                               float* __restrict__ grad_part) {
    __shared__ float xs[GV_N][GV_MAXF];
    __shared__ float ms[GV_N][GV_MAXU];      // mean message
    __shared__ float dp[GV_N][GV_MAXU];      // d pre
    __shared__ float dm[GV_N][GV_MAXU];      // d mean message
    __shared__ float ds[GV_N][GV_MAXU];      // d sender half, per SENDER node
    __shared__ float dw[GV_N][GV_MAXU];      // d receiver ("own") half, per node
    __shared__ float as[GV_N][GV_N];
    __shared__ int cn[GV_N];
    const int tid = threadIdx.x, n = tid / GV_MAXU, u = tid % GV_MAXU;
    float gm[GV_RPT], gu[GV_RPT], gb = 0.f;
#pragma unroll
    for (int i = 0; i < GV_RPT; ++i) { gm[i] = 0.f; gu[i] = 0.f; }
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < GV_N * F; i += GV_NT) xs[i / F][i % F] = x[b * GV_N * F + i];
        if (tid < GV_N * GV_N) as[tid / GV_N][tid % GV_N] = adj[b * GV_N * GV_N + tid];
        __syncthreads();
        if (tid < GV_N) {
            int c = 0;
            for (int s = 0; s < GV_N; ++s) c += as[s][tid] != 0.f ? 1 : 0;
            cn[tid] = c;
        }
        if (u < U) {      // forward: mean message of node n
            float own = 0.f;
            for (int f = 0; f < F; ++f) own = fmaf(xs[n][f], __ldg(Wmsg + (size_t)(F + f) * U + u), own);
            float sum = 0.f;
            int cnt = 0;
            for (int s = 0; s < GV_N; ++s) {
                if (as[s][n] != 0.f) {
                    float e = own;
                    for (int f = 0; f < F; ++f) e = fmaf(xs[s][f], __ldg(Wmsg + (size_t)f * U + u), e);
                    sum += e;
                    ++cnt;
                }
            }
            ms[n][u] = cnt ? sum / (float)cnt : 0.f;
        }
        __syncthreads();
        if (u < U) {      // forward: pre-activation -> d pre
            float v = bias ? bias[u] : 0.f;
            for (int f = 0; f < F; ++f) v = fmaf(xs[n][f], __ldg(Wupd + (size_t)f * U + u), v);
            for (int k = 0; k < U; ++k) v = fmaf(ms[n][k], __ldg(Wupd + (size_t)(F + k) * U + u), v);
            const float yv = gv_act(v, act);
            const float g = dy[(b * GV_N + n) * U + u] * (act == 1 ? 1.f - yv * yv : 1.f);
            dp[n][u] = g;
            gb += g;      // summed over the four node threads of column u at the end
        }
        __syncthreads();
        if (u < U) {      // d mean message of node n, unit k = u
            float v = 0.f;
            for (int j = 0; j < U; ++j) v = fmaf(dp[n][j], __ldg(Wupd + (size_t)(F + u) * U + j), v);
            dm[n][u] = v;
        }
        __syncthreads();
        if (u < U) {
            dw[n][u] = cn[n] > 0 ? dm[n][u] : 0.f;
            float v = 0.f;      // node n as SENDER: sum over its receivers r of d m_r / cnt_r
            for (int r = 0; r < GV_N; ++r)
                if (as[n][r] != 0.f) v += dm[r][u] / (float)cn[r];
            ds[n][u] = v;
        }
        __syncthreads();
        if (dx) {      // thread (n, f = u): d x_n[f]
            if (u < F) {
                float v = 0.f;
                for (int j = 0; j < U; ++j) {
                    v = fmaf(dp[n][j], __ldg(Wupd + (size_t)u * U + j), v);
                    v = fmaf(dw[n][j], __ldg(Wmsg + (size_t)(F + u) * U + j), v);
                    v = fmaf(ds[n][j], __ldg(Wmsg + (size_t)u * U + j), v);
                }
                dx[(b * GV_N + n) * F + u] = v;
            }
        }
        if (u < U) {      // weight gradients: rows r = n + 4 i of column u
#pragma unroll
            for (int i = 0; i < GV_RPT; ++i) {
                const int r = n + 4 * i;
                if (r < F + U) {
                    float a = 0.f;
#pragma unroll
                    for (int q = 0; q < GV_N; ++q) a = fmaf(r < F ? xs[q][r] : ms[q][r - F], dp[q][u], a);
                    gu[i] += a;
                }
                if (r < 2 * F) {
                    float a = 0.f;
#pragma unroll
                    for (int q = 0; q < GV_N; ++q) a = fmaf(r < F ? xs[q][r] : xs[q][r - F], r < F ? ds[q][u] : dw[q][u], a);
                    gm[i] += a;
                }
            }
        }
    }
    const int NPw = 2 * F * U + (F + U) * U + U, NPs = (NPw + 3) & ~3;
    float* gp = grad_part + (int64_t)blockIdx.x * NPs;
    __syncthreads();
    if (u < U) {
#pragma unroll
        for (int i = 0; i < GV_RPT; ++i) {
            const int r = n + 4 * i;
            if (r < 2 * F) gp[r * U + u] = gm[i];
            if (r < F + U) gp[2 * F * U + r * U + u] = gu[i];
        }
        dp[n][u] = gb;
    }
    __syncthreads();
    if (n == 0 && u < U) gp[2 * F * U + (F + U) * U + u] = (dp[0][u] + dp[1][u]) + (dp[2][u] + dp[3][u]);
    if (tid < NPs - NPw) gp[NPw + tid] = 0.f;
}

__global__ void __launch_bounds__(GV_NT) gat1_backward_kernel(const float* __restrict__ x, const float* __restrict__ adj,
                                                              const float* __restrict__ Wpre, const float* __restrict__ watt,
                                                              const float* __restrict__ bias, const float* __restrict__ dy,
                                                              int64_t B, int F, int U, int act, float* __restrict__ dx,
                                                              float* __restrict__ grad_part) {
    __shared__ float xs[GV_N][GV_MAXF];
    __shared__ float xp[GV_N][GV_MAXU];
    __shared__ float dq[GV_N][GV_MAXU];      // d o (gradient at the attention output), then d x'
    __shared__ float dxp[GV_N][GV_MAXU];
    __shared__ float as[GV_N][GV_N], att[GV_N][GV_N], zz[GV_N][GV_N], datt[GV_N][GV_N], dz[GV_N][GV_N];
    __shared__ float ps[GV_N], pr[GV_N], dps[GV_N], dpr[GV_N];
    __shared__ float part[GV_NT / 32][2];
    const int tid = threadIdx.x, n = tid / GV_MAXU, u = tid % GV_MAXU, lane = tid & 31;
    float gw[GV_RPT / 2], ga0 = 0.f, ga1 = 0.f, gb = 0.f;      // W_pre rows r = n + 4 i (F <= 64 -> 16 rows); w_att halves; bias
#pragma unroll
    for (int i = 0; i < GV_RPT / 2; ++i) gw[i] = 0.f;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < GV_N * F; i += GV_NT) xs[i / F][i % F] = x[b * GV_N * F + i];
        if (tid < GV_N * GV_N) {
            const int s = tid / GV_N, r = tid % GV_N;
            as[s][r] = fminf(1.f, adj[b * GV_N * GV_N + tid] + (s == r ? 1.f : 0.f));
        }
        __syncthreads();
        float v = 0.f;
        if (u < U)
            for (int f = 0; f < F; ++f) v = fmaf(xs[n][f], __ldg(Wpre + (size_t)f * U + u), v);
        if (u < U) xp[n][u] = v;
        float a0 = u < U ? v * __ldg(watt + u) : 0.f, a1 = u < U ? v * __ldg(watt + U + u) : 0.f;
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) { part[tid >> 5][0] = a0; part[tid >> 5][1] = a1; }
        __syncthreads();
        if (tid < GV_N) {
            ps[tid] = part[2 * tid][0] + part[2 * tid + 1][0];
            pr[tid] = part[2 * tid][1] + part[2 * tid + 1][1];
        }
        __syncthreads();
        if (tid < GV_N) {      // receiver r = tid: softmax over its incoming edges
            const int r = tid;
            float e[GV_N], sum = 0.f;
            for (int s = 0; s < GV_N; ++s) {
                const float z = ps[s] + pr[r];
                zz[s][r] = z;
                const float a = z > 0.f ? z : 0.2f * z;
                e[s] = as[s][r] != 0.f ? expf(a) : 0.f;
                sum += e[s];
            }
            for (int s = 0; s < GV_N; ++s) att[s][r] = as[s][r] != 0.f ? e[s] / sum : 0.f;
        }
        __syncthreads();
        if (u < U) {      // forward output of node n -> d o
            float o = 0.f;
            for (int r = 0; r < GV_N; ++r) o = fmaf(att[n][r], xp[r][u], o);
            if (bias) o += bias[u];
            const float yv = gv_act(o, act);
            const float g = dy[(b * GV_N + n) * U + u] * (act == 1 ? 1.f - yv * yv : 1.f);
            dq[n][u] = g;
            gb += g;
        }
        __syncthreads();
        if (tid < GV_N * GV_N) {      // d att[s][r] = do_s . x'_r
            const int s = tid / GV_N, r = tid % GV_N;
            float a = 0.f;
            for (int j = 0; j < U; ++j) a = fmaf(dq[s][j], xp[r][j], a);
            datt[s][r] = a;
        }
        __syncthreads();
        if (tid < GV_N * GV_N) {      // softmax over the senders of receiver r, then leaky relu
            const int s = tid / GV_N, r = tid % GV_N;
            float dot = 0.f;
            for (int q = 0; q < GV_N; ++q) dot = fmaf(att[q][r], datt[q][r], dot);
            const float da = att[s][r] * (datt[s][r] - dot);
            dz[s][r] = as[s][r] != 0.f ? da * (zz[s][r] > 0.f ? 1.f : 0.2f) : 0.f;
        }
        __syncthreads();
        if (tid < GV_N) {
            float a = 0.f, c = 0.f;
            for (int q = 0; q < GV_N; ++q) { a += dz[tid][q]; c += dz[q][tid]; }
            dps[tid] = a;      // node as sender
            dpr[tid] = c;      // node as receiver
        }
        __syncthreads();
        if (u < U) {      // d x'_n[u] = sum_s att[s][n] do_s[u] + dps_n w_att[u] + dpr_n w_att[U + u]
            float g = 0.f;
            for (int s = 0; s < GV_N; ++s) g = fmaf(att[s][n], dq[s][u], g);
            g = fmaf(dps[n], __ldg(watt + u), g);
            g = fmaf(dpr[n], __ldg(watt + U + u), g);
            dxp[n][u] = g;
            if (n == 0) {      // w_att gradients of unit u: fixed node order
                float s0 = 0.f, s1 = 0.f;
                for (int q = 0; q < GV_N; ++q) { s0 = fmaf(dps[q], xp[q][u], s0); s1 = fmaf(dpr[q], xp[q][u], s1); }
                ga0 += s0;
                ga1 += s1;
            }
        }
        __syncthreads();
        if (dx && u < F) {
            float g = 0.f;
            for (int j = 0; j < U; ++j) g = fmaf(dxp[n][j], __ldg(Wpre + (size_t)u * U + j), g);
            dx[(b * GV_N + n) * F + u] = g;
        }
        if (u < U) {
#pragma unroll
            for (int i = 0; i < GV_RPT / 2; ++i) {
                const int r = n + 4 * i;
                if (r < F) {
                    float a = 0.f;
#pragma unroll
                    for (int q = 0; q < GV_N; ++q) a = fmaf(xs[q][r], dxp[q][u], a);
                    gw[i] += a;
                }
            }
        }
    }
    const int NPw = F * U + 2 * U + U, NPs = (NPw + 3) & ~3;
    float* gp = grad_part + (int64_t)blockIdx.x * NPs;
    __syncthreads();
    if (u < U) {
#pragma unroll
        for (int i = 0; i < GV_RPT / 2; ++i) {
            const int r = n + 4 * i;
            if (r < F) gp[r * U + u] = gw[i];
        }
        if (n == 0) { gp[F * U + u] = ga0; gp[F * U + U + u] = ga1; }
        dq[n][u] = gb;
    }
    __syncthreads();
    if (n == 0 && u < U) gp[F * U + 2 * U + u] = (dq[0][u] + dq[1][u]) + (dq[2][u] + dq[3][u]);
    if (tid < NPs - NPw) gp[NPw + tid] = 0.f;
}

// symm_norm: D^-1/2 A D^-1/2, D = row sums (zero degree -> inf * 0 = NaN like the TF expression)
__global__ void symm_norm_kernel(const float* __restrict__ adj, int64_t B, int N, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N * N) return;
    const int64_t b = i / (N * N);
    const int r = (int)((i / N) % N), c = (int)(i % N);
    float dr = 0.f, dc = 0.f;
    for (int k = 0; k < N; ++k) {
        dr += adj[(b * N + r) * N + k];
        dc += adj[(b * N + c) * N + k];
    }
    out[i] = (1.f / sqrtf(dr)) * adj[i] * (1.f / sqrtf(dc));
}

// segment_softmax: exp(data) / (sum of exp(data) over the entries with the same segment id)
__global__ void segsm_sum_kernel(const float* __restrict__ data, const int32_t* __restrict__ seg, int64_t E, int C,
                                 int64_t S, float* __restrict__ sums, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * C) return;
    const int64_t e = i / C;
    const int32_t s = seg[e];
    if (s < 0 || s >= S) { *bad = 1; return; }
    atomicAdd(sums + (int64_t)s * C + (i % C), expf(data[i]));
}
__global__ void segsm_div_kernel(const float* __restrict__ data, const int32_t* __restrict__ seg, int64_t E, int C,
                                 int64_t S, const float* __restrict__ sums, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E * C) return;
    const int32_t s = seg[i / C];
    if (s < 0 || s >= S) return;
    out[i] = expf(data[i]) / sums[(int64_t)s * C + (i % C)];
}

}  // namespace ddrl

using namespace ddrl;

static int gv_check(const void* x, const void* adj, const void* W0, const void* W1, const void* y, int64_t B, int F, int U,
                    int act, const char* who) {
    DDRL_REQUIRE(x && adj && W0 && W1 && y && B >= 0, DDRL_E_BADARG, "%s: null pointer or bad B", who);
    DDRL_REQUIRE(F >= 1 && F <= GV_MAXF && U >= 1 && U <= GV_MAXU && (act == 0 || act == 1), DDRL_E_UNSUPPORTED_SHAPE,
                 "%s: unsupported F=%d U=%d act=%d (F, U <= 64; act 0 none, 1 tanh)", who, F, U, act);
    return DDRL_OK;
}

extern "C" int ddrl_mpnn2_forward(const float* x, const float* adj, const float* W_msg, const float* W_upd, const float* b,
                                  int64_t B, int F, int U, int act, float* y, void* stream) {
    const int rc = gv_check(x, adj, W_msg, W_upd, y, B, F, U, act, "mpnn2_forward");
    if (rc != DDRL_OK) return rc;
    if (B == 0) return DDRL_OK;
    mpnn2_forward_kernel<<<(int)std::min<int64_t>(B, 148 * 8), GV_NT, 0, (cudaStream_t)stream>>>(x, adj, W_msg, W_upd, b, B, F, U, act, y);
    DDRL_CHECK_LAUNCH("mpnn2_forward");
    return DDRL_OK;
}

extern "C" int ddrl_gat1_forward(const float* x, const float* adj, const float* W_pre, const float* w_att, const float* b,
                                 int64_t B, int F, int U, int act, float* y, void* stream) {
    const int rc = gv_check(x, adj, W_pre, w_att, y, B, F, U, act, "gat1_forward");
    if (rc != DDRL_OK) return rc;
    if (B == 0) return DDRL_OK;
    gat1_forward_kernel<<<(int)std::min<int64_t>(B, 148 * 8), GV_NT, 0, (cudaStream_t)stream>>>(x, adj, W_pre, w_att, b, B, F, U, act, y);
    DDRL_CHECK_LAUNCH("gat1_forward");
    return DDRL_OK;
}

extern "C" int ddrl_symm_norm(const float* adj, int64_t B, int N, float* out, void* stream) {
    DDRL_REQUIRE(adj && out && B >= 0 && N >= 1 && N <= 64, DDRL_E_BADARG, "symm_norm: null pointer or bad B/N");
    if (B == 0) return DDRL_OK;
    const int64_t n = B * N * N;
    symm_norm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(adj, B, N, out);
    DDRL_CHECK_LAUNCH("symm_norm");
    return DDRL_OK;
}

extern "C" int ddrl_segment_softmax(const float* data, const int32_t* segment_ids, int64_t E, int C, int64_t num_segments,
                                    float* sums_ws, int* bad_id, float* out, void* stream) {
    DDRL_REQUIRE(data && segment_ids && sums_ws && bad_id && out && E >= 0 && C >= 1 && num_segments >= 1, DDRL_E_BADARG,
                 "segment_softmax: null pointer or bad shape");
    if (E == 0) return DDRL_OK;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(sums_ws, 0, (size_t)num_segments * C * sizeof(float), st);
    cudaMemsetAsync(bad_id, 0, sizeof(int), st);
    const unsigned nb = (unsigned)((E * C + 255) / 256);
    segsm_sum_kernel<<<nb, 256, 0, st>>>(data, segment_ids, E, C, num_segments, sums_ws, bad_id);
    DDRL_CHECK_LAUNCH("segment_softmax(sum)");
    segsm_div_kernel<<<nb, 256, 0, st>>>(data, segment_ids, E, C, num_segments, sums_ws, out);
    DDRL_CHECK_LAUNCH("segment_softmax(div)");
    return DDRL_OK;
}

extern "C" int ddrl_mpnn2_backward(const float* x, const float* adj, const float* W_msg, const float* W_upd, const float* b,
                                   const float* dy, int64_t B, int F, int U, int act, int ctas, float* dx, float* grad_part,
                                   void* stream) {
    const int rc = gv_check(x, adj, W_msg, W_upd, grad_part, B, F, U, act, "mpnn2_backward");
    if (rc != DDRL_OK) return rc;
    DDRL_REQUIRE(dy && ctas >= 1, DDRL_E_BADARG, "mpnn2_backward: null dy or bad ctas");
    mpnn2_backward_kernel<<<ctas, GV_NT, 0, (cudaStream_t)stream>>>(x, adj, W_msg, W_upd, b, dy, B, F, U, act, dx, grad_part);
    DDRL_CHECK_LAUNCH("mpnn2_backward");
    return DDRL_OK;
}

extern "C" int ddrl_gat1_backward(const float* x, const float* adj, const float* W_pre, const float* w_att, const float* b,
                                  const float* dy, int64_t B, int F, int U, int act, int ctas, float* dx, float* grad_part,
                                  void* stream) {
    const int rc = gv_check(x, adj, W_pre, w_att, grad_part, B, F, U, act, "gat1_backward");
    if (rc != DDRL_OK) return rc;
    DDRL_REQUIRE(dy && ctas >= 1, DDRL_E_BADARG, "gat1_backward: null dy or bad ctas");
    gat1_backward_kernel<<<ctas, GV_NT, 0, (cudaStream_t)stream>>>(x, adj, W_pre, w_att, b, dy, B, F, U, act, dx, grad_part);
    DDRL_CHECK_LAUNCH("gat1_backward");
    return DDRL_OK;
}
