// Tensor-core (tcgen05) FCNet training step: packed weight image and shared-memory layout (byte offsets).
#pragma once
#include "common.cuh"

namespace ddrl {

constexpr int TC_ROWS = 128;   // rows per tile = UMMA M
constexpr int TC_NT = 256;     // threads per CTA: thread = (row, 32-column half of the 64-wide branch tile)
constexpr int TC_NO = 16;      // head-gradient operand width (2A <= 16 outputs, value head uses column 0)

__host__ __device__ inline int tc_kx(int D) { return (D + 1 + 15) & ~15; }   // >= 1 pad column (holds the constant 1)

// chunked fp16 matrix [rows][cols]: element offset of (r, c), see umma.cuh
__host__ __device__ inline int tc_chunk_off(int rows, int r, int c) { return ((c >> 3) * rows + r) * 8 + (c & 7); }

// Image (== first part of shared memory), per policy.  Branch 0 = policy, 1 = value.
//   W1[b][hi|lo] : fp16 [64 out c][KX in d] chunked along d   (F1: B operand, K-major)
//   W2[b][hi|lo] : fp16 [64 in j][64 out c] chunked along c   (F2: B MN-major;  B4: B K-major)
//   f32 block    : b1c[128] b2c[128] Wo[64][2A] bo[pad4] Wvo[64] bvo[4]
struct TcImg {
    int W1[2][2], W2[2][2], b1c, b2c, Wo, bo, Wvo, bvo, bytes;
};
__host__ __device__ inline TcImg tc_img(int D, int A) {
    const int KX = tc_kx(D);
    TcImg L;
    int p = 0;
    for (int b = 0; b < 2; ++b)
        for (int h = 0; h < 2; ++h) { L.W1[b][h] = p; p += 64 * KX * 2; }
    for (int b = 0; b < 2; ++b)
        for (int h = 0; h < 2; ++h) { L.W2[b][h] = p; p += 64 * 64 * 2; }
    L.b1c = p; p += 128 * 4;
    L.b2c = p; p += 128 * 4;
    L.Wo = p;  p += 64 * 2 * A * 4;
    L.bo = p;  p += ((2 * A + 3) & ~3) * 4;
    L.Wvo = p; p += 64 * 4;
    L.bvo = p; p += 16;
    L.bytes = (p + 15) & ~15;
    return L;
}

struct TcSmem {
    int X[2], H1[2], H2[2], DL[2], xraw, pf, hpart, dlf, red, bar, total;
};
__host__ __device__ inline TcSmem tc_smem(int D, int A) {
    const int KX = tc_kx(D);
    const int LDo = 2 * A + 1;
    TcSmem s;
    int p = tc_img(D, A).bytes;
    for (int h = 0; h < 2; ++h) { s.X[h] = p; p += TC_ROWS * KX * 2; }
    for (int h = 0; h < 2; ++h) { s.H1[h] = p; p += TC_ROWS * 64 * 2; }
    for (int h = 0; h < 2; ++h) { s.H2[h] = p; p += TC_ROWS * 64 * 2; }
    for (int h = 0; h < 2; ++h) { s.DL[h] = p; p += TC_ROWS * TC_NO * 2; }
    s.xraw = p;  p += ((TC_ROWS * D * 4) + 15) & ~15;
    s.pf = p;    p += ((TC_ROWS * (3 * A + 4) * 4) + 15) & ~15;
    s.hpart = p; p += ((TC_ROWS * LDo * 4) + 15) & ~15;
    s.dlf = p;   p += ((TC_ROWS * LDo * 4) + 15) & ~15;
    s.red = p;   p += 8 * 8 * 8;     // [8 warps][8 stats] doubles
    s.bar = p;   p += 32;            // mbarrier (8 B) + tmem slot (4 B)
    s.total = p;
    return s;
}

// Where flat parameter j lives in the image: fp16 pair (byte offsets of hi and lo, f16 = true) or one float.
__host__ __device__ inline void tc_img_pos(const TcImg& L, const FcOffsets& o, int D, int A, int j, bool& f16, int& p0, int& p1) {
    const int KX = tc_kx(D), A2 = 2 * A;
    f16 = false; p1 = -1;
    auto w1 = [&](int b, int i) { const int d = i >> 6, c = i & 63; f16 = true;
                                  const int e = tc_chunk_off(64, c, d) * 2; p0 = L.W1[b][0] + e; p1 = L.W1[b][1] + e; };
    auto w2 = [&](int b, int i) { const int jn = i >> 6, c = i & 63; f16 = true;
                                  const int e = tc_chunk_off(64, jn, c) * 2; p0 = L.W2[b][0] + e; p1 = L.W2[b][1] + e; };
    if (j < o.b1)       w1(0, j - o.W1);
    else if (j < o.Wv1) p0 = L.b1c + (j - o.b1) * 4;
    else if (j < o.bv1) w1(1, j - o.Wv1);
    else if (j < o.W2)  p0 = L.b1c + (64 + j - o.bv1) * 4;
    else if (j < o.b2)  w2(0, j - o.W2);
    else if (j < o.Wv2) p0 = L.b2c + (j - o.b2) * 4;
    else if (j < o.bv2) w2(1, j - o.Wv2);
    else if (j < o.Wo)  p0 = L.b2c + (64 + j - o.bv2) * 4;
    else if (j < o.bo)  p0 = L.Wo + (j - o.Wo) * 4;
    else if (j < o.Wvo) p0 = L.bo + (j - o.bo) * 4;
    else if (j < o.bvo) p0 = L.Wvo + (j - o.Wvo) * 4;
    else                p0 = L.bvo;
    (void)KX; (void)A2;
}

}  // namespace ddrl
