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
