// tcgen05 tensor-core path (sm_100a): UMMA self test + (later) the tensor-core FCNet step.
#include <algorithm>

#include "common.cuh"
#include "umma.cuh"

namespace ddrl {

// D[128][N] = sum_k A(m,k) * B(n,k) on the 5th-gen tensor cores through TMEM, with the operands staged in the
// chunked shared-memory layout of umma.cuh.  A(m,k) = Abuf[m][k] (K-major) or Abuf[k][m] (MN-major view of the same
// kind of buffer); B alike.  split != 0: every operand is split into fp16 (hi, lo) and three products
// hi*hi + hi*lo + lo*hi are accumulated in FP32 (the precision scheme of the tensor-core training step).
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, int ra, int ca, const float* __restrict__ B, int rb, int cb, int N,
                     int K, int a_mn, int b_mn, int split, float* __restrict__ D, int* __restrict__ status) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mbar;
    __half* sAh = reinterpret_cast<__half*>(smraw);
    __half* sAl = sAh + ra * ca;
    __half* sBh = sAl + ra * ca;
    __half* sBl = sBh + rb * cb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < ra * ca; i += blockDim.x) {
        const int r = i / ca, c = i - r * ca;
        __half hi, lo;
        umma::split_f16(A[i], hi, lo);
        const int o = ((c >> 3) * ra + r) * 8 + (c & 7);
        sAh[o] = hi; sAl[o] = lo;
    }
    for (int i = tid; i < rb * cb; i += blockDim.x) {
        const int r = i / cb, c = i - r * cb;
        __half hi, lo;
        umma::split_f16(B[i], hi, lo);
        const int o = ((c >> 3) * rb + r) * 8 + (c & 7);
        sBh[o] = hi; sBl[o] = lo;
    }
    const uint32_t ncols = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) { umma::mbar_init(&mbar, 1); umma::fence_mbar_init(); }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t taddr = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = umma::idesc_f16(128, N, a_mn != 0, b_mn != 0);
        const int nprod = split ? 3 : 1;
        bool acc = false;
        for (int pr = 0; pr < nprod; ++pr) {
            const __half* pa = pr == 2 ? sAl : sAh;
            const __half* pb = pr == 1 ? sBl : sBh;
            for (int ks = 0; ks < K / 16; ++ks) {
                // K-major: 2 chunks of 8 per MMA, chunk stride rows*16 B.  MN-major: 16 K rows, 16 B each.
                const uint32_t aoff = a_mn ? ks * 16 * 16 : ks * 2 * ra * 16;
                const uint32_t boff = b_mn ? ks * 16 * 16 : ks * 2 * rb * 16;
                const uint64_t ad = a_mn ? umma::desc_mnmajor(umma::smem_u32(pa) + aoff, ra) : umma::desc_kmajor(umma::smem_u32(pa) + aoff, ra);
                const uint64_t bd = b_mn ? umma::desc_mnmajor(umma::smem_u32(pb) + boff, rb) : umma::desc_kmajor(umma::smem_u32(pb) + boff, rb);
                umma::mma_f16(taddr, ad, bd, idesc, acc);
                acc = true;
            }
        }
        umma::mma_commit(&mbar);
    }
    const bool ok = umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (tid == 0) *status = 1;   // timed out: report, do not touch TMEM results
    } else {
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            umma::tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j < N) D[(warp * 32 + lane) * N + c0 + j] = v[j];
        }
        if (tid == 0) *status = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, ncols);
}

}  // namespace ddrl

using namespace ddrl;

extern "C" int ddrl_umma_selftest(const float* A, int ra, int ca, const float* B, int rb, int cb, int N, int K, int a_mn,
                                  int b_mn, int split, float* D, int* status, void* stream) {
    DDRL_REQUIRE(A && B && D && status, DDRL_E_BADARG, "umma_selftest: null pointer");
    DDRL_REQUIRE(ra % 8 == 0 && rb % 8 == 0 && ca % 8 == 0 && cb % 8 == 0 && K % 16 == 0 && N % 16 == 0 && N >= 16 && N <= 256,
                 DDRL_E_UNSUPPORTED_SHAPE, "umma_selftest: shapes must be multiples of 8 (rows/cols), 16 (K, N)");
    DDRL_REQUIRE((a_mn ? (ca == 128 && ra >= K) : (ra == 128 && ca >= K)) && (b_mn ? (cb >= N && rb >= K) : (rb >= N && cb >= K)),
                 DDRL_E_UNSUPPORTED_SHAPE, "umma_selftest: buffer shapes do not cover M=128 x N x K");
    const size_t smem = (size_t)(ra * ca + rb * cb) * 2 * sizeof(__half) + 128;
    DDRL_REQUIRE(smem <= 200 * 1024, DDRL_E_UNSUPPORTED_SHAPE, "umma_selftest: %zu bytes of shared memory", smem);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            set_error("umma_selftest: cannot raise dynamic shared memory");
            return DDRL_E_CUDA;
        }
        attr = true;
    }
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, ra, ca, B, rb, cb, N, K, a_mn, b_mn, split, D, status);
    DDRL_CHECK_LAUNCH("umma_selftest");
    return DDRL_OK;
}
