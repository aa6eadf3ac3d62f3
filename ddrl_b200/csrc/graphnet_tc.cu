// GraphNet SGD step on tensor cores (sm_100a): ONE persistent kernel runs forward + PPO loss + backward + the fused
// gradient-reduce / [NVLink all-reduce] / clip / TF1-Adam tail (sgd_tail.cuh) for every optimizer step of a launch.
//
// Reference math: models/graph_net.py:10-45 (hyper-network leg encoder, MPNN, linear head), models/gcn.py:57-94 (MPNN:
// act(X W_upd + mean_{s->r}(X_s W_msg))), models/shared_graphnet_glorot_uniform_init.py:21-58 (actor GraphNet(2A) + critic
// GraphNet(1)); loss: RLlib 1.0.1 ppo_tf_policy.PPOLoss (ppo_loss.cuh).
//
// Work split.  grid = (Gn, 2): blockIdx.y = net (0 actor, 1 critic), the Gn CTAs of a net share the minibatch rows; a CTA
// walks its rows in tiles of 64.  Per tile:
//   pairs     the (row, node) pairs the output depends on: the controlled node and its in-neighbours (<= 4 per row)
//   encoder   FMA + MUFU pipes (NOT a contraction: the generated [19, 64] matrix passes through tanh, 1216 tanh per pair —
//             the MUFU pipe bounds the step): lane (h, f-quarter) keeps its 5 encoder columns {W_e[:, f*64+h], b_e} in
//             registers, w = tanh(e . W_e + b_e), x[h] = tanh(sum_f leg[f] w[f, h]) with two shuffles; no block barrier
//             inside the pair loop
//   MPNN      tcgen05 (kind::f16, fp16 hi/lo operand split x 3 products, FP32 accumulation in TMEM, M = 64):
//             Ypre = Xc W_upd + Xmean W_msg -> tanh -> head Y W_out -> PPO loss -> dY = dl W_out^T ->
//             dXc = dYpre W_upd^T, dXmean = dYpre W_msg^T, and the weight gradients as K = rows GEMMs
//             gW_upd = Xc^T dYpre, gW_msg = Xmean^T dYpre, gW_out = Y^T dl (accumulated in TMEM over the CTA's tiles)
//   encoder^T recompute w, dw = dx_pre[h] leg[f] (1 - w^2); gb_e += dw, gW_e[k] += e[k] dw in registers (thread-owned
//             columns, accumulated over all pairs of the CTA; no atomics)
// then one partial gradient per CTA pair (actor CTA i and critic CTA i fill the two halves of row i) and the fused tail.
// All sums have a fixed order: bit-reproducible.
#include <algorithm>

#include "tc_common.cuh"

namespace ddrl {

constexpr int GT_ROWS = 64;        // rows per tile = UMMA M
constexpr int GT_PAIRS = 256;      // (row, node) pairs per tile (<= 4 per row)
constexpr int GT_NT = 512;         // 16 warps
constexpr int GT_F = DDRL_GN_FEATS, GT_E = DDRL_GN_ENC_IN, GT_S = GT_F + GT_E, GT_N = DDRL_GN_NODES, GT_H = DDRL_HIDDEN;
constexpr int GT_LEG = 32;         // staged leg features per pair: 4 lane groups x 8 floats (5 used: f = 5 g + k; zero pads)
constexpr int GT_KC = 5;           // encoder columns per lane
constexpr int GT_XS = 68;          // row stride of XQ in floats (64 + 4: rows of consecutive pairs fall into different banks)

// TMEM columns (M = 64 accumulators: row m lives in lane 32 * (m / 16) + m % 16)
constexpr int GC_YP = 0, GC_HO = 64, GC_DY = 96, GC_DXC = 160, GC_DXM = 224, GC_GWU = 288, GC_GWM = 352, GC_GWO = 416,
              GC_COLS = 512;

struct GnTcOffsets { int We, be, Wm, Wu, Wo, bo, NP; };
__host__ __device__ inline GnTcOffsets gnt_offsets(int O) {
    GnTcOffsets o;
    int p = 0;
    o.We = p; p += GT_E * GT_F * GT_H;
    o.be = p; p += GT_F * GT_H;
    o.Wm = p; p += GT_H * GT_H;
    o.Wu = p; p += GT_H * GT_H;
    o.Wo = p; p += GT_H * O;
    o.bo = p; p += O;
    o.NP = p;
    return o;
}

struct GnTcSmem {
    int Wu[2], Wm[2], WoT[2], bout, Xc[2], Xm[2], Y[2], DL[2], XQ, sE, sLeg, pinfo, rowmeta, red, bar, total;
};
__host__ __device__ inline GnTcSmem gnt_smem() {
    GnTcSmem s;
    int p = 0;
    for (int h = 0; h < 2; ++h) { s.Wu[h] = p; p += GT_H * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.Wm[h] = p; p += GT_H * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.WoT[h] = p; p += TC_NO * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.Xc[h] = p; p += GT_ROWS * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.Xm[h] = p; p += GT_ROWS * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.Y[h] = p; p += GT_ROWS * GT_H * 2; }
    for (int h = 0; h < 2; ++h) { s.DL[h] = p; p += GT_ROWS * TC_NO * 2; }
    s.XQ = p;      p += GT_PAIRS * GT_XS * 4;      // x of every pair (forward), then d x_pre (backward); scratch outside the tiles
    s.sE = p;      p += GT_PAIRS * GT_E * 4;
    s.sLeg = p;    p += GT_PAIRS * GT_LEG * 4;
    s.pinfo = p;   p += GT_PAIRS * 4;             // row | node << 8
    s.rowmeta = p; p += GT_ROWS * 8 * 4;          // per row: pair of the controlled node, #senders, sender pairs [4], #pairs, offset
    s.bout = p;    p += 16 * 4;
    s.red = p;     p += 1024;
    s.bar = p;     p += 32;
    s.total = p;
    return s;
}

struct GnTcArgs {
    const float* theta;
    const int32_t* node_idx;
    const float *state, *adj, *actions, *old_logits, *old_logp, *vf_preds, *adv, *vtarg;
    int64_t R;
    int A, MB;
    const int32_t* mb_perm;
    const int32_t* step_ctr;
    const float* kl_coeff;
    ddrl_ppo_hyper hp;
    float* grad_part;      // [Gn][NPs]: row i = partial of (actor CTA i | critic CTA i)
    double* stat_part;     // [2 Gn][DDRL_NSTAT]: row i filled by BOTH CTAs i (disjoint entries); rows >= Gn stay zero
    int* status;
    SgdTail tail;
};

// Encoder: the weights are pre-multiplied by 2 log2(e), so tanh(pre) = 1 - 2 r with r = 1 / (2^pre' + 1) — no |x| / copysign
// (2^pre' overflows to +inf -> r = 0 -> 1; underflows to 0 -> r = 1 -> -1), and 1 - tanh^2 = 4 r (1 - r).
constexpr float kTanhIn = 2.885390081777927f;
__device__ __forceinline__ float gnt_rexp(float pre_scaled) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(pre_scaled));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.f));
    return r;
}

// Packed FP32 pairs (sm_100a FFMA2 / FADD2 / FMUL2: one issue slot for two FP32 operations — the encoder is issue bound).
__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ float2 f2b(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 gnt_rexp2(float2 pre_scaled) {      // r = 1 / (2^pre' + 1), both halves
    float2 t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(pre_scaled.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(pre_scaled.y));
    t = __fadd2_rn(t, f2b(1.f));
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(t.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(t.y));
    return r;
}
// pre' of a column pair: b + e . W   (e broadcast to both halves)
__device__ __forceinline__ float2 gnt_pre2(const float4& e, const float2 (&W)[4], float2 b) {
    float2 q = __ffma2_rn(f2b(e.x), W[0], b);
    q = __ffma2_rn(f2b(e.y), W[1], q);
    q = __ffma2_rn(f2b(e.z), W[2], q);
    return __ffma2_rn(f2b(e.w), W[3], q);
}

__device__ __forceinline__ float gnt_tanh(float x) {      // same as graphnet.cu: ex2 + rcp, absolute error <= 2.4e-7
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.f));
    return copysignf(fmaf(-2.f, r, 1.f), x);
}

__global__ void __launch_bounds__(GT_NT, 1) graphnet_train_tc_kernel(const __grid_constant__ GnTcArgs a) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const GnTcSmem S = gnt_smem();
    const int net = blockIdx.y, Gn = gridDim.x, bx = blockIdx.x;
    const int A = a.A, A2 = 2 * A, O = net == 0 ? A2 : 1;
    const GnTcOffsets oa = gnt_offsets(A2), o = gnt_offsets(O);
    const int NPtot = oa.NP + gnt_offsets(1).NP, NPs = (NPtot + 3) & ~3;
    const int net_base = net == 0 ? 0 : oa.NP;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cq = warp >> 2;                 // TMEM lane quadrant / 16-column group of the epilogues
    const bool mine = lane < 16;                            // M = 64: only the first 16 lanes of a quadrant hold rows
    const int erow = 16 * q + lane;                         // row of this thread in the epilogues (valid if mine)
    const int hgrp = warp & 7, pset = warp >> 3;            // encoder: 8 hidden units per warp, two pair sets
    const int hh = lane & 7, fq = lane >> 3, eh = 8 * hgrp + hh, f0 = GT_KC * fq;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(sm + S.bar);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(sm + S.bar + 16);
    float* sgs = reinterpret_cast<float*>(sm + S.bar + 24);
    const uint32_t sbase = umma::smem_u32(sm);
    float* XQ = reinterpret_cast<float*>(sm + S.XQ);
    float* sE = reinterpret_cast<float*>(sm + S.sE);
    float* sLeg = reinterpret_cast<float*>(sm + S.sLeg);
    int* pinfo = reinterpret_cast<int*>(sm + S.pinfo);
    int* rowmeta = reinterpret_cast<int*>(sm + S.rowmeta);
    float* sbout = reinterpret_cast<float*>(sm + S.bout);
    float* sred = reinterpret_cast<float*>(sm + S.red);

    if (warp == 0) umma::tmem_alloc(tslot, GC_COLS);
    if (tid == 0) { umma::mbar_init(mbar, 1); umma::fence_mbar_init(); }
    for (int i = tid; i < (S.Xc[0] - S.WoT[0]) / 4; i += GT_NT) reinterpret_cast<uint32_t*>(sm + S.WoT[0])[i] = 0u;   // pad rows of WoT
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tslot;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const uint32_t tm_u = __shfl_sync(0xffffffffu, tmem, 0), sb_u = __shfl_sync(0xffffffffu, sbase, 0);
    uint32_t ph = 0;
    bool ok = true;
    int ovf = 0;
    const float klc = a.kl_coeff[0];
    const bool has_tail = a.tail.theta != nullptr;
    const int nsteps = (has_tail && a.tail.nsteps > 1) ? a.tail.nsteps : 1;
    const int step0 = a.step_ctr ? *a.step_ctr : 0;
    const int G = 2 * Gn, flat = net * Gn + bx;
    TailStep ts;
    ts.round = 0; ts.last = false; ts.nsteps = nsteps; ts.b1p = 0.f; ts.b2p = 0.f; ts.seq = 0u; ts.epoch = 0u;
    if (has_tail) {
        ts.b1p = __ldcg(a.tail.beta_pow);
        ts.b2p = __ldcg(a.tail.beta_pow + 1);
        ts.seq = (a.tail.world > 1) ? *a.tail.seq : 0u;
        ts.epoch = __ldcg(a.tail.barrier_ws + 4 + 1);
    }
    float* gp = a.grad_part + (int64_t)bx * NPs + net_base;
    const float inv = a.hp.inv_global_mb;

    auto mma_wait = [&]() {      // every thread: wait for the commit of the MMA batch issued after the last barrier
        ok = umma::mbar_wait(mbar, ph) && ok;
        ph ^= 1;
        umma::fence_after_sync();
    };
    auto hand_off = [&]() {      // generic shared-memory writes -> async proxy, TMEM reads retired, block barrier
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    };

    // pull the batch rows [r0, r0 + n) (state, adjacency, node index, loss inputs) into L2 ahead of their tile: the batch is
    // larger than L2, so the staging loads of a tile would otherwise pay the full HBM latency on the critical path
    auto prefetch_rows = [&](int64_t r0, int n) {
        if (n <= 0) return;
        const char* base = reinterpret_cast<const char*>(a.state + r0 * GT_N * GT_S);
        const int lines = (n * GT_N * GT_S * 4 + 127) >> 7;
        for (int i = tid; i < lines; i += GT_NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + 128 * i));
        const char* ab = reinterpret_cast<const char*>(a.adj + r0 * GT_N * GT_N);
        for (int i = tid; i < ((n * GT_N * GT_N * 4 + 127) >> 7); i += GT_NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(ab + 128 * i));
        if (tid < 8) {
            const void* ptrs[8] = {a.node_idx + r0, a.old_logp + r0, a.vf_preds + r0, a.adv + r0, a.vtarg + r0, a.actions + r0 * a.A,
                                   a.old_logits + r0 * 2 * a.A, a.old_logits + r0 * 2 * a.A + 32};
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ptrs[tid]));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(ptrs[tid]) + 128));
        }
    };

#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        if (s > 0) {      // every CTA must have written its Adam slice of the previous step
            if (tid == 0 && !sgd_wait_weights(a.tail, 0, G, s)) { ok = false; if (a.status) atomicOr(a.status, 64); }
            __syncthreads();
        }
        // ---- this step's weights: encoder columns -> registers, MPNN / head matrices -> fp16 hi/lo UMMA images -------------
        const float* th = a.theta + net_base;
        float We[GT_KC][GT_E], be[GT_KC];
#pragma unroll
        for (int k = 0; k < GT_KC; ++k) {
            const int f = f0 + k;
            const bool v = k < GT_KC && f < GT_F && (fq < 3 || k < 4);
#pragma unroll
            for (int e = 0; e < GT_E; ++e) We[k][e] = v ? kTanhIn * __ldcg(th + o.We + e * GT_F * GT_H + f * GT_H + eh) : 0.f;
            be[k] = v ? kTanhIn * __ldcg(th + o.be + f * GT_H + eh) : 0.f;
        }
        // packed column pairs of this lane: (k0, k1), (k2, k3) and k4 duplicated (its two halves serve the two pairs of an iteration)
        float2 W01[GT_E], W23[GT_E], W4[GT_E];
#pragma unroll
        for (int e = 0; e < GT_E; ++e) { W01[e] = f2(We[0][e], We[1][e]); W23[e] = f2(We[2][e], We[3][e]); W4[e] = f2b(We[4][e]); }
        const float2 b01 = f2(be[0], be[1]), b23 = f2(be[2], be[3]), b4 = f2b(be[4]);
        for (int i = tid; i < 2 * GT_H * GT_H; i += GT_NT) {
            const int mtx = i >> 12, j = i & 4095, r = j >> 6, c = j & 63;      // W[r = in][c = out], chunked along c
            const float w = __ldcg(th + (mtx ? o.Wm : o.Wu) + j) * TC_SW;
            __half hi, lo;
            umma::split_f16(w, hi, lo);
            const int off = tc_chunk_off(GT_H, r, c) * 2;
            *reinterpret_cast<__half*>(sm + (mtx ? S.Wm[0] : S.Wu[0]) + off) = hi;
            *reinterpret_cast<__half*>(sm + (mtx ? S.Wm[1] : S.Wu[1]) + off) = lo;
        }
        for (int i = tid; i < GT_H * O; i += GT_NT) {
            const int k = i / O, oo = i - k * O;                                  // WoT[o][k] = W_out[k][o], chunked along k
            const float w = __ldcg(th + o.Wo + i) * TC_SW;
            __half hi, lo;
            umma::split_f16(w, hi, lo);
            const int off = tc_chunk_off(TC_NO, oo, k) * 2;
            *reinterpret_cast<__half*>(sm + S.WoT[0] + off) = hi;
            *reinterpret_cast<__half*>(sm + S.WoT[1] + off) = lo;
        }
        if (tid < 16) sbout[tid] = tid < O ? __ldcg(th + o.bo + tid) : 0.f;

        const int step = step0 + s;
        const int mb = a.mb_perm ? a.mb_perm[step] : step;
        const int64_t mb0 = (int64_t)mb * a.MB, mb1 = min(mb0 + a.MB, a.R);
        const int rpc = (((a.MB + Gn - 1) / Gn) + 7) & ~7;
        const int64_t cr0 = min(mb0 + (int64_t)bx * rpc, mb1), cr1 = min(cr0 + rpc, mb1);

        // gradient accumulators of the packed columns: [e] = dW_e row e, [4] = db_e; g4: halves = the two pairs of an iteration
        float2 g01[GT_E + 1], g23[GT_E + 1], g4[GT_E + 1];
#pragma unroll
        for (int e = 0; e <= GT_E; ++e) { g01[e] = f2b(0.f); g23[e] = f2b(0.f); g4[e] = f2b(0.f); }
        double st[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) st[i] = 0.0;
        float gbo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) gbo[i] = 0.f;
        bool first = true;
        __syncthreads();

#pragma unroll 1
        for (int64_t row0 = cr0; row0 < cr1; row0 += GT_ROWS) {
            const int nrows = (int)min((int64_t)GT_ROWS, cr1 - row0);
            prefetch_rows(row0 + GT_ROWS, (int)min((int64_t)GT_ROWS, cr1 - row0 - GT_ROWS));      // next tile -> L2
            // ---- pairs of this tile ---------------------------------------------------------------------------------------
            if (tid < GT_ROWS) {
                int need = 0, snd = 0, idx = 0;
                if (tid < nrows) {
                    const int64_t r = row0 + tid;
                    idx = min(max(a.node_idx[r], 0), GT_N - 1);
#pragma unroll
                    for (int n = 0; n < GT_N; ++n) {
                        const bool sd = a.adj[r * GT_N * GT_N + n * GT_N + idx] != 0.f;
                        snd |= (sd ? 1 : 0) << n;
                        need |= ((sd || n == idx) ? 1 : 0) << n;
                    }
                }
                rowmeta[tid * 8 + 6] = __popc(need);
                rowmeta[tid * 8 + 7] = need | (snd << 4) | (idx << 8);
            }
            __syncthreads();
            if (tid < GT_ROWS) {
                int off = 0;
                for (int j = 0; j < tid; ++j) off += rowmeta[j * 8 + 6];
                const int code = rowmeta[tid * 8 + 7], need = code & 15, snd = (code >> 4) & 15, idx = code >> 8;
                int cnt = 0, pc = -1;
#pragma unroll
                for (int n = 0; n < GT_N; ++n) {
                    rowmeta[tid * 8 + 2 + n] = -1;
                    if ((need >> n) & 1) {
                        pinfo[off] = tid | (n << 8);
                        if (n == idx) pc = off;
                        if ((snd >> n) & 1) rowmeta[tid * 8 + 2 + cnt++] = off;
                        ++off;
                    }
                }
                rowmeta[tid * 8 + 0] = pc;
                rowmeta[tid * 8 + 1] = cnt;
                if (tid == GT_ROWS - 1) reinterpret_cast<int*>(sred)[0] = off;      // pairs in this tile
            }
            __syncthreads();
            const int npairs = reinterpret_cast<const int*>(sred)[0];
            // stage e [4] and leg [19] of every pair: asynchronous 4-byte copies, all in flight at once (32 slots per pair: slot
            // t < 23 copies float t of the node's state vector)
            for (int i = tid; i < npairs * 32; i += GT_NT) {
                const int pr = i >> 5, j = i & 31;
                if (j < GT_S) {
                    const int code = pinfo[pr];
                    const float* src = a.state + ((row0 + (code & 255)) * GT_N + (code >> 8)) * GT_S + j;
                    float* dst = j < GT_F ? sLeg + pr * GT_LEG + (j / GT_KC) * 8 + (j % GT_KC) : sE + pr * GT_E + (j - GT_F);
                    tc_cp4(dst, src);
                } else if (j == GT_S) {
                    sLeg[pr * GT_LEG + 3 * 8 + 4] = 0.f;      // f = 19 does not exist
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();

            // ---- encoder forward: x of every pair -> XQ (two pairs per iteration, packed FP32 pairs) ------------------------
#pragma unroll 1
            for (int i = pset; i < npairs; i += 4) {
                const int i2 = min(i + 2, npairs - 1);
                const float4 ea = *reinterpret_cast<const float4*>(sE + i * GT_E);
                const float4 eb = *reinterpret_cast<const float4*>(sE + i2 * GT_E);
                const float4 la = *reinterpret_cast<const float4*>(sLeg + i * GT_LEG + 8 * fq);
                const float4 lb = *reinterpret_cast<const float4*>(sLeg + i2 * GT_LEG + 8 * fq);
                const float2 l4 = f2(sLeg[i * GT_LEG + 8 * fq + 4], sLeg[i2 * GT_LEG + 8 * fq + 4]);
                const float2 ra01 = gnt_rexp2(gnt_pre2(ea, W01, b01)), ra23 = gnt_rexp2(gnt_pre2(ea, W23, b23));
                const float2 rb01 = gnt_rexp2(gnt_pre2(eb, W01, b01)), rb23 = gnt_rexp2(gnt_pre2(eb, W23, b23));
                float2 q4 = __ffma2_rn(f2(ea.x, eb.x), W4[0], b4);
                q4 = __ffma2_rn(f2(ea.y, eb.y), W4[1], q4);
                q4 = __ffma2_rn(f2(ea.z, eb.z), W4[2], q4);
                q4 = __ffma2_rn(f2(ea.w, eb.w), W4[3], q4);
                const float2 r4 = gnt_rexp2(q4);
                const float2 m2 = f2b(-2.f), one = f2b(1.f);
                float2 sa = __fmul2_rn(f2(la.x, la.y), __ffma2_rn(m2, ra01, one));
                sa = __ffma2_rn(f2(la.z, la.w), __ffma2_rn(m2, ra23, one), sa);
                float2 sb = __fmul2_rn(f2(lb.x, lb.y), __ffma2_rn(m2, rb01, one));
                sb = __ffma2_rn(f2(lb.z, lb.w), __ffma2_rn(m2, rb23, one), sb);
                const float2 s4 = __fmul2_rn(l4, __ffma2_rn(m2, r4, one));
                float pa = (sa.x + sa.y) + s4.x, pb = (sb.x + sb.y) + s4.y;
                pa += __shfl_xor_sync(0xffffffffu, pa, 8);
                pb += __shfl_xor_sync(0xffffffffu, pb, 8);
                pa += __shfl_xor_sync(0xffffffffu, pa, 16);
                pb += __shfl_xor_sync(0xffffffffu, pb, 16);
                if (fq == 0) XQ[i * GT_XS + eh] = gnt_tanh(pa);
                if (fq == 1 && i + 2 < npairs) XQ[i2 * GT_XS + eh] = gnt_tanh(pb);
            }
            __syncthreads();
            // ---- MPNN operands: Xc (controlled node), Xmean (mean over the senders) -> fp16 hi/lo, chunked [64][64] -------
            {
                const int r = tid & 63, c8 = tid >> 6;
                float xc[8], xm[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { xc[j] = 0.f; xm[j] = 0.f; }
                if (r < nrows) {
                    const int pc = rowmeta[r * 8 + 0], cnt = rowmeta[r * 8 + 1];
                    const float4* src = reinterpret_cast<const float4*>(XQ + pc * GT_XS + 8 * c8);
                    const float4 v0 = src[0], v1 = src[1];
                    xc[0] = v0.x; xc[1] = v0.y; xc[2] = v0.z; xc[3] = v0.w; xc[4] = v1.x; xc[5] = v1.y; xc[6] = v1.z; xc[7] = v1.w;
                    for (int sidx = 0; sidx < cnt; ++sidx) {
                        const float4* sp = reinterpret_cast<const float4*>(XQ + rowmeta[r * 8 + 2 + sidx] * GT_XS + 8 * c8);
                        const float4 u0 = sp[0], u1 = sp[1];
                        xm[0] += u0.x; xm[1] += u0.y; xm[2] += u0.z; xm[3] += u0.w; xm[4] += u1.x; xm[5] += u1.y; xm[6] += u1.z; xm[7] += u1.w;
                    }
                    const float ic = cnt > 0 ? 1.f / (float)cnt : 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) xm[j] *= ic;
                }
                uint4 hi, lo;
                tc_split8(xc, TC_SH, hi, lo);
                *reinterpret_cast<uint4*>(sm + S.Xc[0] + (c8 * GT_ROWS + r) * 16) = hi;
                *reinterpret_cast<uint4*>(sm + S.Xc[1] + (c8 * GT_ROWS + r) * 16) = lo;
                tc_split8(xm, TC_SH, hi, lo);
                *reinterpret_cast<uint4*>(sm + S.Xm[0] + (c8 * GT_ROWS + r) * 16) = hi;
                *reinterpret_cast<uint4*>(sm + S.Xm[1] + (c8 * GT_ROWS + r) * 16) = lo;
            }
            hand_off();
            // ---- Ypre = Xc W_upd + Xmean W_msg ------------------------------------------------------------------------------
            if (warp == 0) {
                tc_gemm_u(tm_u + GC_YP, sb_u + S.Xc[0], sb_u + S.Xc[1], GT_ROWS, false, sb_u + S.Wu[0], sb_u + S.Wu[1], GT_H, true,
                          64, 64, 4, false, 3);
                tc_gemm_u(tm_u + GC_YP, sb_u + S.Xm[0], sb_u + S.Xm[1], GT_ROWS, false, sb_u + S.Wm[0], sb_u + S.Wm[1], GT_H, true,
                          64, 64, 4, true, 3);
                umma::mma_commit_elect(mbar);
            }
            mma_wait();
            {   // y = tanh(Ypre) -> fp16 hi/lo chunked [64][64]
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[8];
                    umma::tmem_ld8(tmem + tlane + GC_YP + 16 * cq + 8 * c, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = gnt_tanh(v[j] * (1.f / (TC_SH * TC_SW)));
                    if (mine) {
                        uint4 hi, lo;
                        tc_split8(v, TC_SH, hi, lo);
                        const int off = ((2 * cq + c) * GT_ROWS + erow) * 16;
                        *reinterpret_cast<uint4*>(sm + S.Y[0] + off) = hi;
                        *reinterpret_cast<uint4*>(sm + S.Y[1] + off) = lo;
                    }
                }
            }
            hand_off();
            // ---- head: Hout[64][16] = Y WoT^T ---------------------------------------------------------------------------------
            if (warp == 0) {
                tc_gemm_u(tm_u + GC_HO, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS, false, sb_u + S.WoT[0], sb_u + S.WoT[1], TC_NO, false,
                          64, TC_NO, 4, false, 3);
                umma::mma_commit_elect(mbar);
            }
            mma_wait();
            // ---- PPO loss (actor CTAs: policy part, critic CTAs: value part) -> DL (fp16 hi/lo x gradient scale) -----------
            if (cq == 0) {
                float out[16], dl[16];
                umma::tmem_ld16(tmem + tlane + GC_HO, out);
#pragma unroll
                for (int i = 0; i < 16; ++i) dl[i] = 0.f;
                if (mine && erow < nrows) {
                    const int64_t gr = row0 + erow;
                    double sv[DDRL_NSTAT];
#pragma unroll
                    for (int i = 0; i < DDRL_NSTAT; ++i) sv[i] = 0.0;
                    if (net == 0) {
                        float lg[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) lg[i] = fmaf(out[i], 1.f / (TC_SH * TC_SW), sbout[i]);
                        ppo_row_policy(lg, A, a.actions + gr * A, a.old_logits + gr * A2, a.old_logp[gr], a.adv[gr], klc,
                                       a.hp.clip_param, a.hp.entropy_coeff, 1.f, dl, sv);
                        st[0] += sv[0]; st[1] += sv[1]; st[2] += sv[3];
                    } else {
                        const float val = fmaf(out[0], 1.f / (TC_SH * TC_SW), sbout[0]);
                        dl[0] = ppo_row_value(val, a.vf_preds[gr], a.vtarg[gr], a.hp.vf_clip_param, a.hp.vf_loss_coeff, 1.f, sv);
                        st[0] += sv[2]; st[1] += sv[4]; st[2] += sv[5]; st[3] += sv[6]; st[4] += sv[7];
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) gbo[i] += dl[i];
                }
                if (first) {   // one power-of-two gradient scale per CTA and step: max|dl| * scale ~ TC_GTARGET
                    float mx = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) mx = fmaxf(mx, fabsf(dl[i]));
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                    if (lane == 0) sred[8 + q] = mx;
                    asm volatile("bar.sync 1, 128;" ::: "memory");      // the four loss warps
                    mx = fmaxf(fmaxf(sred[8], sred[9]), fmaxf(sred[10], sred[11]));
                    int e = 0;
                    if (mx > 0.f && mx < 3.0e38f) e = (int)floorf(log2f(TC_GTARGET / mx));
                    e = max(-20, min(20, e));
                    if (q == 0 && lane == 0) sgs[0] = exp2f((float)e);
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                const float sg = sgs[0];
                if (mine) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint4 hi, lo;
                        ovf |= tc_split8(&dl[8 * c], sg, hi, lo) ? 8 : 0;
                        *reinterpret_cast<uint4*>(sm + S.DL[0] + (c * GT_ROWS + erow) * 16) = hi;
                        *reinterpret_cast<uint4*>(sm + S.DL[1] + (c * GT_ROWS + erow) * 16) = lo;
                    }
                }
            }
            hand_off();
            const float sg = sgs[0];
            // ---- dY = DL WoT;  gW_out (+)= Y^T DL --------------------------------------------------------------------------
            if (warp == 0) {
                tc_gemm_u(tm_u + GC_DY, sb_u + S.DL[0], sb_u + S.DL[1], GT_ROWS, false, sb_u + S.WoT[0], sb_u + S.WoT[1], TC_NO, true,
                          64, 64, 1, false, 3);
                tc_gemm_u(tm_u + GC_GWO, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS, true, sb_u + S.DL[0], sb_u + S.DL[1], GT_ROWS, true,
                          64, TC_NO, 4, !first, 3);
                umma::mma_commit_elect(mbar);
            }
            mma_wait();
            {   // dYpre = dY (1 - y^2), in place over Y (fp16 hi/lo x gradient scale)
                float mxv = 0.f;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    float v[8], y[8];
                    umma::tmem_ld8(tmem + tlane + GC_DY + 16 * cq + 8 * c, v);
                    if (mine) {
                        const int off = ((2 * cq + c) * GT_ROWS + erow) * 16;
                        tc_join8(*reinterpret_cast<const uint4*>(sm + S.Y[0] + off), *reinterpret_cast<const uint4*>(sm + S.Y[1] + off),
                                 1.f / TC_SH, y);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j] = v[j] * (1.f / (sg * TC_SW)) * (1.f - y[j] * y[j]);
                            mxv = fmaxf(mxv, fabsf(v[j]));
                        }
                        uint4 hi, lo;
                        tc_split8(v, sg, hi, lo);
                        *reinterpret_cast<uint4*>(sm + S.Y[0] + off) = hi;
                        *reinterpret_cast<uint4*>(sm + S.Y[1] + off) = lo;
                    }
                }
                if (!(mxv * sg <= 60000.f)) ovf |= 16;
            }
            hand_off();
            // ---- dXc = dYpre W_upd^T, dXmean = dYpre W_msg^T;  gW_upd (+)= Xc^T dYpre, gW_msg (+)= Xmean^T dYpre -----------
            if (warp == 0) {
                // four independent accumulators: issue product by product round robin, so consecutive MMAs never wait for each
                // other's accumulate (a dependent chain runs at ~100 cycles per MMA, independent ones at ~45)
#pragma unroll
                for (int pr = 0; pr < 3; ++pr) {
                    const int pm = 1 << pr;
                    tc_gemm_mask_u(tm_u + GC_DXC, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS, false, sb_u + S.Wu[0], sb_u + S.Wu[1], GT_H,
                                   false, 64, 64, 4, pr > 0, pm);
                    tc_gemm_mask_u(tm_u + GC_DXM, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS, false, sb_u + S.Wm[0], sb_u + S.Wm[1], GT_H,
                                   false, 64, 64, 4, pr > 0, pm);
                    tc_gemm_mask_u(tm_u + GC_GWU, sb_u + S.Xc[0], sb_u + S.Xc[1], GT_ROWS, true, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS,
                                   true, 64, 64, 4, pr > 0 || !first, pm);
                    tc_gemm_mask_u(tm_u + GC_GWM, sb_u + S.Xm[0], sb_u + S.Xm[1], GT_ROWS, true, sb_u + S.Y[0], sb_u + S.Y[1], GT_ROWS,
                                   true, 64, 64, 4, pr > 0 || !first, pm);
                }
                umma::mma_commit_elect(mbar);
            }
            mma_wait();
            {   // d x_pre of every pair of this thread's row: (dXc if controlled) + (dXmean / #senders if sender), x (1 - x^2)
                float dc[16], dm[16];
                umma::tmem_ld16(tmem + tlane + GC_DXC + 16 * cq, dc);
                umma::tmem_ld16(tmem + tlane + GC_DXM + 16 * cq, dm);
                if (mine && erow < nrows) {
                    const int pc = rowmeta[erow * 8 + 0], cnt = rowmeta[erow * 8 + 1];
                    const float ic = cnt > 0 ? 1.f / (float)cnt : 0.f;
                    const float k1 = 1.f / (sg * TC_SW);
                    // the pairs of a row are consecutive: [off, off + n)
                    int off = 0x7fffffff, n = rowmeta[erow * 8 + 6];
                    off = pc;
                    for (int sidx = 0; sidx < cnt; ++sidx) off = min(off, rowmeta[erow * 8 + 2 + sidx]);
                    for (int pi = off; pi < off + n; ++pi) {
                        bool is_s = false;
                        for (int sidx = 0; sidx < cnt; ++sidx) is_s = is_s || rowmeta[erow * 8 + 2 + sidx] == pi;
                        const float wc = pi == pc ? k1 : 0.f, ws = is_s ? k1 * ic : 0.f;
                        float* xp = XQ + pi * GT_XS + 16 * cq;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float x = xp[j];
                            xp[j] = (dc[j] * wc + dm[j] * ws) * (1.f - x * x);
                        }
                    }
                }
            }
            umma::fence_before_sync();
            __syncthreads();
            umma::fence_after_sync();
            // ---- encoder backward: recompute r, accumulate gW_e / gb_e of this lane's columns (two pairs per iteration) ------
#pragma unroll 1
            for (int i = pset; i < npairs; i += 4) {
                const int i2 = min(i + 2, npairs - 1);
                const float4 ea = *reinterpret_cast<const float4*>(sE + i * GT_E);
                const float4 eb = *reinterpret_cast<const float4*>(sE + i2 * GT_E);
                const float4 la = *reinterpret_cast<const float4*>(sLeg + i * GT_LEG + 8 * fq);
                const float4 lb = *reinterpret_cast<const float4*>(sLeg + i2 * GT_LEG + 8 * fq);
                const float2 l4 = f2(sLeg[i * GT_LEG + 8 * fq + 4], sLeg[i2 * GT_LEG + 8 * fq + 4]);
                const float da = 4.f * XQ[i * GT_XS + eh], db = (i + 2 < npairs) ? 4.f * XQ[i2 * GT_XS + eh] : 0.f;
                // dw = dx_pre leg (1 - w^2) = (4 dx_pre leg) (r - r^2)
                float2 r = gnt_rexp2(gnt_pre2(ea, W01, b01));
                float2 dw = __fmul2_rn(__fmul2_rn(f2b(da), f2(la.x, la.y)), __ffma2_rn(f2(-r.x, -r.y), r, r));
                g01[4] = __fadd2_rn(g01[4], dw);
                g01[0] = __ffma2_rn(f2b(ea.x), dw, g01[0]); g01[1] = __ffma2_rn(f2b(ea.y), dw, g01[1]);
                g01[2] = __ffma2_rn(f2b(ea.z), dw, g01[2]); g01[3] = __ffma2_rn(f2b(ea.w), dw, g01[3]);
                r = gnt_rexp2(gnt_pre2(eb, W01, b01));
                dw = __fmul2_rn(__fmul2_rn(f2b(db), f2(lb.x, lb.y)), __ffma2_rn(f2(-r.x, -r.y), r, r));
                g01[4] = __fadd2_rn(g01[4], dw);
                g01[0] = __ffma2_rn(f2b(eb.x), dw, g01[0]); g01[1] = __ffma2_rn(f2b(eb.y), dw, g01[1]);
                g01[2] = __ffma2_rn(f2b(eb.z), dw, g01[2]); g01[3] = __ffma2_rn(f2b(eb.w), dw, g01[3]);
                r = gnt_rexp2(gnt_pre2(ea, W23, b23));
                dw = __fmul2_rn(__fmul2_rn(f2b(da), f2(la.z, la.w)), __ffma2_rn(f2(-r.x, -r.y), r, r));
                g23[4] = __fadd2_rn(g23[4], dw);
                g23[0] = __ffma2_rn(f2b(ea.x), dw, g23[0]); g23[1] = __ffma2_rn(f2b(ea.y), dw, g23[1]);
                g23[2] = __ffma2_rn(f2b(ea.z), dw, g23[2]); g23[3] = __ffma2_rn(f2b(ea.w), dw, g23[3]);
                r = gnt_rexp2(gnt_pre2(eb, W23, b23));
                dw = __fmul2_rn(__fmul2_rn(f2b(db), f2(lb.z, lb.w)), __ffma2_rn(f2(-r.x, -r.y), r, r));
                g23[4] = __fadd2_rn(g23[4], dw);
                g23[0] = __ffma2_rn(f2b(eb.x), dw, g23[0]); g23[1] = __ffma2_rn(f2b(eb.y), dw, g23[1]);
                g23[2] = __ffma2_rn(f2b(eb.z), dw, g23[2]); g23[3] = __ffma2_rn(f2b(eb.w), dw, g23[3]);
                float2 q4 = __ffma2_rn(f2(ea.x, eb.x), W4[0], b4);
                q4 = __ffma2_rn(f2(ea.y, eb.y), W4[1], q4);
                q4 = __ffma2_rn(f2(ea.z, eb.z), W4[2], q4);
                q4 = __ffma2_rn(f2(ea.w, eb.w), W4[3], q4);
                r = gnt_rexp2(q4);
                dw = __fmul2_rn(__fmul2_rn(f2(da, db), l4), __ffma2_rn(f2(-r.x, -r.y), r, r));
                g4[4] = __fadd2_rn(g4[4], dw);
                g4[0] = __ffma2_rn(f2(ea.x, eb.x), dw, g4[0]); g4[1] = __ffma2_rn(f2(ea.y, eb.y), dw, g4[1]);
                g4[2] = __ffma2_rn(f2(ea.z, eb.z), dw, g4[2]); g4[3] = __ffma2_rn(f2(ea.w, eb.w), dw, g4[3]);
            }
            first = false;
            __syncthreads();
        }

        // ---- write-out: this CTA's half of partial row bx ------------------------------------------------------------------
        const bool any = cr1 > cr0;
        {   // encoder gradients: add the two pair sets (fixed order), thread-owned columns -> flat order
            float* scr = XQ;      // [8 hgrp][32 lanes][25]
            float gWe[GT_KC][GT_E], gbe[GT_KC];      // unpack: column k, row e
#pragma unroll
            for (int e = 0; e < GT_E; ++e) {
                gWe[0][e] = g01[e].x; gWe[1][e] = g01[e].y; gWe[2][e] = g23[e].x; gWe[3][e] = g23[e].y; gWe[4][e] = g4[e].x + g4[e].y;
            }
            gbe[0] = g01[4].x; gbe[1] = g01[4].y; gbe[2] = g23[4].x; gbe[3] = g23[4].y; gbe[4] = g4[4].x + g4[4].y;
            if (pset == 1) {
                float* d = scr + (hgrp * 32 + lane) * 25;
#pragma unroll
                for (int k = 0; k < GT_KC; ++k) {
                    d[5 * k + 4] = gbe[k];
#pragma unroll
                    for (int e = 0; e < GT_E; ++e) d[5 * k + e] = gWe[k][e];
                }
            }
            __syncthreads();
            if (pset == 0) {
                const float* d = scr + (hgrp * 32 + lane) * 25;
#pragma unroll
                for (int k = 0; k < GT_KC; ++k) {
                    const int f = f0 + k;
                    if (f < GT_F && (fq < 3 || k < 4)) {
                        gp[o.be + f * GT_H + eh] = (gbe[k] + d[5 * k + 4]) * inv;
#pragma unroll
                        for (int e = 0; e < GT_E; ++e) gp[o.We + e * GT_F * GT_H + f * GT_H + eh] = (gWe[k][e] + d[5 * k + e]) * inv;
                    }
                }
            }
        }
        {   // MPNN / head gradients from TMEM (zero when the CTA had no rows: the accumulators were never written)
            const float sgw = any ? sgs[0] : 1.f;
            const float kw = inv / (TC_SH * sgw);
            float u[16], m[16];
            if (any) {
                umma::tmem_ld16(tmem + tlane + GC_GWU + 16 * cq, u);
                umma::tmem_ld16(tmem + tlane + GC_GWM + 16 * cq, m);
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) { u[j] = 0.f; m[j] = 0.f; }
            }
            if (mine) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    gp[o.Wu + erow * GT_H + 16 * cq + j] = u[j] * kw;
                    gp[o.Wm + erow * GT_H + 16 * cq + j] = m[j] * kw;
                }
            }
            if (cq == 0) {
                float w[16];
                if (any) umma::tmem_ld16(tmem + tlane + GC_GWO, w);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = 0.f;
                }
                if (mine) {
                    for (int oo = 0; oo < O; ++oo) gp[o.Wo + erow * O + oo] = w[oo] * kw;
                }
                // head bias gradient and loss statistics: the four loss warps -> per-warp sums
                double* redd = reinterpret_cast<double*>(sred) + 8;      // [4 warps][24]
                for (int i = 0; i < 16; ++i) {
                    const float sx = warp_sum(gbo[i]);
                    if (lane == 0) redd[q * 24 + 8 + i] = (double)sx;
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const double sx = warp_sum(st[i]);
                    if (lane == 0) redd[q * 24 + i] = sx;
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        {
            const double* redd = reinterpret_cast<const double*>(sred) + 8;
            auto sum4 = [&](int i) { return (redd[i] + redd[24 + i]) + (redd[48 + i] + redd[72 + i]); };
            if (tid < O) gp[o.bo + tid] = (float)sum4(8 + tid) * inv;
            if (net == 1 && tid >= 32 && tid < 32 + (NPs - NPtot)) a.grad_part[(int64_t)bx * NPs + NPtot + (tid - 32)] = 0.f;   // padding floats
            if (tid < DDRL_NSTAT && a.stat_part) {
                // stat slots: 0 -surr, 1 KL, 2 vf, 3 entropy, 4 R, 5 R^2, 6 R-v, 7 (R-v)^2; actor fills 0 1 3, critic the rest
                const bool act_slot = tid == 0 || tid == 1 || tid == 3;
                const int idx = tid == 0 ? 0 : tid == 1 ? 1 : tid == 3 ? 2 : tid == 2 ? 0 : tid - 3;
                if (act_slot == (net == 0)) a.stat_part[(int64_t)bx * DDRL_NSTAT + tid] = sum4(idx);
            }
        }
        if (s + 1 < nsteps) {      // the next step's first tile does not depend on the weights: pull it into L2 behind the tail
            const int mbn = a.mb_perm ? a.mb_perm[step + 1] : step + 1;
            const int64_t n0 = (int64_t)mbn * a.MB, n1 = min(n0 + a.MB, a.R);
            const int64_t c0n = min(n0 + (int64_t)bx * rpc, n1), c1n = min(c0n + rpc, n1);
            prefetch_rows(c0n, (int)min((int64_t)GT_ROWS, c1n - c0n));
        }
        if (has_tail) {
            ts.round = s + 1;
            ts.last = s == nsteps - 1;
            const bool tok = sgd_step_tail(a.tail, ts, a.grad_part, a.stat_part, 0, 1, flat, G, NPtot, step, 0, 0, XQ, nullptr, Gn);
            ok = ok && tok;
            ts.b1p *= a.tail.beta1;
            ts.b2p *= a.tail.beta2;
            ts.seq += 1u;
        }
    }
    if (a.status) {
        if (tid == 0 && !ok) atomicOr(a.status, 1);
        if (ovf) atomicOr(a.status, ovf);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, GC_COLS);
}

}  // namespace ddrl

using namespace ddrl;

extern "C" int ddrl_graphnet_train_step_tc(const float* theta, const int32_t* node_idx, const float* state, const float* adj,
                                           const float* actions, const float* old_logits, const float* old_logp,
                                           const float* vf_preds, const float* adv, const float* vtarg, int64_t R, int A, int MB,
                                           const int32_t* mb_perm, const int32_t* step_ctr, const float* kl_coeff,
                                           const ddrl_ppo_hyper* hyper, int ctas_per_net, float* grad_part, double* stat_part,
                                           int* status, const ddrl_sgd_tail* tail, void* stream) {
    DDRL_REQUIRE(theta && node_idx && state && adj && actions && old_logits && old_logp && vf_preds && adv && vtarg && kl_coeff &&
                     hyper && grad_part && stat_part,
                 DDRL_E_BADARG, "graphnet_train_step_tc: null pointer");
    DDRL_REQUIRE(R >= 1 && MB >= 1 && ctas_per_net >= 1, DDRL_E_BADARG, "graphnet_train_step_tc: bad R/MB/ctas");
    DDRL_REQUIRE(A >= 1 && A <= DDRL_MAX_ACT, DDRL_E_UNSUPPORTED_SHAPE, "graphnet_train_step_tc: unsupported A=%d", A);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    DDRL_REQUIRE(2 * ctas_per_net <= sms, DDRL_E_BADARG, "graphnet_train_step_tc: 2 x %d CTAs must be co-resident (%d SMs)",
                 ctas_per_net, sms);
    GnTcArgs a;
    a.theta = theta; a.node_idx = node_idx; a.state = state; a.adj = adj; a.actions = actions; a.old_logits = old_logits;
    a.old_logp = old_logp; a.vf_preds = vf_preds; a.adv = adv; a.vtarg = vtarg; a.R = R; a.A = A; a.MB = MB;
    a.mb_perm = mb_perm; a.step_ctr = step_ctr; a.kl_coeff = kl_coeff; a.hp = *hyper; a.grad_part = grad_part;
    a.stat_part = stat_part; a.status = status;
    a.tail = SgdTail{};
    if (tail) {
        const int rc = sgd_tail_check(tail, 2 * ctas_per_net, "graphnet_train_step_tc");
        if (rc != DDRL_OK) return rc;
        DDRL_REQUIRE(!tail->fcnet_img && !tail->fcnet_tc_img && !tail->ll_ws, DDRL_E_BADARG,
                     "graphnet_train_step_tc: the tail must not carry FCNet weight images / an LL workspace");
        a.tail = *tail;
        a.tail.grad_acc = nullptr;      // (accumulation vector: FCNet ping-pong kernel only; this kernel writes per-CTA partials)
    }
    const size_t smem = (size_t)gnt_smem().total;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(graphnet_train_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        DDRL_REQUIRE(e == cudaSuccess, DDRL_E_CUDA, "graphnet_train_step_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr = true;
    }
    graphnet_train_tc_kernel<<<dim3(ctas_per_net, 2), GT_NT, smem, (cudaStream_t)stream>>>(a);
    DDRL_CHECK_LAUNCH("graphnet_train_step_tc");
    return DDRL_OK;
}
