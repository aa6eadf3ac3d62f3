// Shared-memory layout of the FCNet kernels and the mapping flat-parameter -> packed image.
#pragma once
#include "common.cuh"

namespace ddrl {

constexpr int H = DDRL_HIDDEN;   // 64
constexpr int HC = 2 * H;        // 128: policy | value concatenated
constexpr int TM = 64;           // rows per tile
constexpr int NT = 256;          // threads per CTA
constexpr int LDH = HC + 4;      // activation row stride (pad keeps 128-bit row reads conflict free)
constexpr int LDT = HC + 4;      // stride of the transposed layer-2 weights
constexpr int LDD = 20;          // stride of the per-row head outputs / head gradients (2A+1 <= 17)
constexpr int MAXHEAD = (H * (2 * DDRL_MAX_ACT + 1) + NT - 1) / NT;  // 5 head-gradient items / thread

// Offsets in floats.  [0, x) is the weight region == the packed image (multiple of 4 floats):
//   W1c [Dp][128]  = [W1 | Wv1] per input row, rows D..Dp zero      b1c [128] = [b1 | bv1]
//   W2c [64][128]  = [W2 | Wv2] per input row                        b2c [128]
//   W2Tc[64][132]  = [W2^T | Wv2^T] (row = output index)             (backward: dz2 . W2^T)
//   WhT [20][128]  : row q < 2A = [Wo[:, q] | 0], row 2A = [0 | Wvo], rest zero   (backward: dl . Wh^T)
//   Wo [64][2A], bo, Wvo [64], bvo                                    (heads, forward)
struct FcSmem {
    int W1c, b1c, W2c, b2c, W2Tc, WhT, Wo, bo, Wvo, bvo, x, x2, pf, h1, h2, out, dl, red, norm, total;
};

__host__ __device__ inline FcSmem fc_smem(int D, int A, bool has_norm, bool train = false) {
    const int Dp = (D + 3) & ~3;
    FcSmem s;
    int p = 0;
    s.W1c = p;  p += Dp * HC;
    s.b1c = p;  p += HC;
    s.W2c = p;  p += H * HC;
    s.b2c = p;  p += HC;
    s.W2Tc = p; p += H * LDT;
    s.WhT = p;  p += LDD * HC;
    s.Wo = p;   p += H * 2 * A;
    s.bo = p;   p += ((2 * A + 3) & ~3);
    s.Wvo = p;  p += H;
    s.bvo = p;  p += 4;
    s.x = p;    p += TM * Dp;
    s.x2 = p;   p += train ? TM * Dp : 0;                 // second x buffer (cp.async double buffering)
    s.pf = p;   p += train ? TM * (3 * A + 4) : 0;        // staged loss inputs: act[64][A] ol[64][2A] 4 x [64]
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

// Position(s) of flat parameter j (checkpoint order) inside the weight region; p1 = -1 if stored once.
__host__ __device__ inline void fc_img_pos(const FcSmem& L, const FcOffsets& o, int D, int A, int j, int& p0, int& p1) {
    const int A2 = 2 * A;
    p1 = -1;
    if (j < o.b1)       { const int i = j - o.W1;  p0 = L.W1c + (i >> 6) * HC + (i & 63); }
    else if (j < o.Wv1) { p0 = L.b1c + (j - o.b1); }
    else if (j < o.bv1) { const int i = j - o.Wv1; p0 = L.W1c + (i >> 6) * HC + H + (i & 63); }
    else if (j < o.W2)  { p0 = L.b1c + H + (j - o.bv1); }
    else if (j < o.b2)  { const int i = j - o.W2;  const int k = i >> 6, c = i & 63;
                          p0 = L.W2c + k * HC + c;      p1 = L.W2Tc + c * LDT + k; }
    else if (j < o.Wv2) { p0 = L.b2c + (j - o.b2); }
    else if (j < o.bv2) { const int i = j - o.Wv2; const int k = i >> 6, c = i & 63;
                          p0 = L.W2c + k * HC + H + c;  p1 = L.W2Tc + c * LDT + H + k; }
    else if (j < o.Wo)  { p0 = L.b2c + H + (j - o.bv2); }
    else if (j < o.bo)  { const int i = j - o.Wo;  const int k = i / A2, q = i - k * A2;
                          p0 = L.Wo + i;                p1 = L.WhT + q * HC + k; }
    else if (j < o.Wvo) { p0 = L.bo + (j - o.bo); }
    else if (j < o.bvo) { const int k = j - o.Wvo; p0 = L.Wvo + k; p1 = L.WhT + A2 * HC + H + k; }
    else                { p0 = L.bvo; }
    (void)D;
}

}  // namespace ddrl
