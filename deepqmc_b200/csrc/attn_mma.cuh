// Softmax attention of a plain forward (S = 1) on the tensor cores: warp-level mma.sync.m16n8k16 with "3xFP16" operands.
//
//   O[b, i, h, :] = sum_j softmax_j(q_i . k_j / sqrt(dh)) v_j        (reference: gnn/update_features.py:273-280
//                                                                      hk.MultiHeadAttention; algebra hkext.py:215-253)
// One warp task = 16 queries of one (walker, head) pair against all of its keys (electrons, plus the TransPsiformer's constant
// nuclear tokens).  Scores S = Q K^T and outputs O = P V are m16n8k16 products of half operands split into hi + lo
// (x 2^e = hi + lo, 22 significant bits; three products  hi.lo + lo.hi + hi.hi  accumulated in fp32: fp32-class accuracy, the
// power-of-two scales are undone exactly).  No shared memory: every fragment is loaded from global memory in the layout the
// instruction wants -- a quad of lanes reads 32 contiguous bytes of a row per instruction (full sectors) -- converted in
// registers, and the probabilities move from the accumulator layout of S to the A-operand layout of P V without leaving the
// register file (the C fragment of two adjacent 8-key tiles IS the A fragment of one 16-key tile).
// dh = 64 only (4 k-tiles of 16); keys <= 8 NK8.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace dq {

#ifdef DQMC_EMU
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  float c[4] = {d[0], d[1], d[2], d[3]};
  emu::mma_m16n8k16_f16(d, a, b, c);
}
__device__ __forceinline__ uint32_t am_pack_half2(float lo, float hi);
__device__ __forceinline__ float am_half_to_float(uint32_t h16) { return emu::h2f((uint16_t)h16); }
#else
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t am_pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float am_half_to_float(uint32_t h16) {
  float f;
  asm("{\n\t.reg .b16 h;\n\tcvt.u16.u32 h, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(f) : "r"(h16));
  return f;
}
#endif

#ifdef DQMC_EMU
namespace attn_emu {
inline uint16_t f2h(float f) {  // round to nearest even
  uint32_t x;
  std::memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const int32_t e = (int32_t)((x >> 23) & 255u) - 127;
  uint32_t m = x & 0x7FFFFFu;
  if (((x >> 23) & 255u) == 255u) return (uint16_t)(sign | 0x7C00u | (m ? 0x200u : 0));
  if (e > 15) return (uint16_t)(sign | 0x7C00u);
  if (e >= -14) {
    uint32_t h = ((uint32_t)(e + 15) << 10) | (m >> 13);
    const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  if (e < -25) return (uint16_t)sign;
  m |= 0x800000u;
  const int shift = -e - 14 + 13;
  uint32_t h = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}
}  // namespace attn_emu
__device__ __forceinline__ uint32_t am_pack_half2(float lo, float hi) {
  return (uint32_t)attn_emu::f2h(lo) | ((uint32_t)attn_emu::f2h(hi) << 16);
}
#endif

// (x0, x1) -> packed hi halves and packed lo halves (x - hi)
__device__ __forceinline__ void am_split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = am_pack_half2(x0, x1);
  lo = am_pack_half2(x0 - am_half_to_float(hi & 0xFFFFu), x1 - am_half_to_float(hi >> 16));
}

// NK8 = number of 8-key tiles (keys padded to 8 NK8 <= 48); block = 4 warps, each warp walks over (pair, query tile) tasks.
template <int NK8>
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(const float* __restrict__ QKV, int ldq, float* __restrict__ O, int ldo, int N, int H, int dmodel, float scale,
                    int n_pairs, const float* __restrict__ Kn, const float* __restrict__ Vn, int Mn) {
  constexpr int DH = 64, NK16 = (NK8 + 1) / 2;
  constexpr float kQS = 16.f, kPS = 1024.f;  // operand scales: q, k, v by 2^4, probabilities by 2^10
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int NKEY = N + Mn;
  const int MT = (N + 15) / 16;  // query tiles per pair
  const int n_tasks = n_pairs * MT;
  const int wglobal = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), wtotal = gridDim.x * (blockDim.x >> 5);
  for (int task = wglobal; task < n_tasks; task += wtotal) {
    const int pair = task / MT, mt = task - pair * MT;
    const int b = pair / H, h = pair - b * H;
    const float* base = QKV + (size_t)b * N * ldq + h * DH;
    // row pointers of key j (K part); the V part sits dmodel further (nuclear tokens: separate arrays)
    auto krow = [&](int j) -> const float* {
      j = j < NKEY ? j : NKEY - 1;  // padded keys read a valid row, their scores are masked
      return j < N ? base + (size_t)j * ldq + dmodel : Kn + (size_t)(j - N) * dmodel + h * DH;
    };
    auto vrow = [&](int j) -> const float* {
      j = j < NKEY ? j : NKEY - 1;
      return j < N ? base + (size_t)j * ldq + 2 * dmodel : Vn + (size_t)(j - N) * dmodel + h * DH;
    };
    // ---- Q fragments of this query tile (rows g, g + 8), hi / lo, 4 k-tiles of 16
    const int q0 = mt * 16 + g, q1 = q0 + 8;
    const float* qp0 = base + (size_t)(q0 < N ? q0 : N - 1) * ldq;
    const float* qp1 = base + (size_t)(q1 < N ? q1 : N - 1) * ldq;
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const float2 a0 = __ldg((const float2*)(qp0 + kt * 16 + 2 * t)), a1 = __ldg((const float2*)(qp1 + kt * 16 + 2 * t));
      const float2 a2 = __ldg((const float2*)(qp0 + kt * 16 + 8 + 2 * t)), a3 = __ldg((const float2*)(qp1 + kt * 16 + 8 + 2 * t));
      am_split2(a0.x * kQS, a0.y * kQS, qh[kt][0], ql[kt][0]);
      am_split2(a1.x * kQS, a1.y * kQS, qh[kt][1], ql[kt][1]);
      am_split2(a2.x * kQS, a2.y * kQS, qh[kt][2], ql[kt][2]);
      am_split2(a3.x * kQS, a3.y * kQS, qh[kt][3], ql[kt][3]);
    }
    // ---- scores: S[16 x 8 NK8] = Q K^T; B fragment of key tile nt: (k = dh index, n = key nt * 8 + g)
    float s[NK8][4];
#pragma unroll
    for (int nt = 0; nt < NK8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const float* kp = krow(nt * 8 + g);
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const float2 b0 = __ldg((const float2*)(kp + kt * 16 + 2 * t)), b1 = __ldg((const float2*)(kp + kt * 16 + 8 + 2 * t));
        uint32_t kh[2], kl[2];
        am_split2(b0.x * kQS, b0.y * kQS, kh[0], kl[0]);
        am_split2(b1.x * kQS, b1.y * kQS, kh[1], kl[1]);
        mma16816(s[nt], qh[kt], kl);
        mma16816(s[nt], ql[kt], kh);
        mma16816(s[nt], qh[kt], kh);
      }
    }
    // ---- softmax over the keys of rows g (c0, c1) and g + 8 (c2, c3); a lane holds keys nt * 8 + 2 t, + 1
    const float us = scale / (kQS * kQS);
    float m0 = -3.0e38f, m1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < NK8; ++nt) {
      const int j0 = nt * 8 + 2 * t;
      s[nt][0] = j0 < NKEY ? s[nt][0] * us : -3.0e38f;
      s[nt][1] = j0 + 1 < NKEY ? s[nt][1] * us : -3.0e38f;
      s[nt][2] = j0 < NKEY ? s[nt][2] * us : -3.0e38f;
      s[nt][3] = j0 + 1 < NKEY ? s[nt][3] * us : -3.0e38f;
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NK8; ++nt) {
      s[nt][0] = m_exp(s[nt][0] - m0); s[nt][1] = m_exp(s[nt][1] - m0);
      s[nt][2] = m_exp(s[nt][2] - m1); s[nt][3] = m_exp(s[nt][3] - m1);
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- probabilities as A fragments of P V: key tile pair (2 kk, 2 kk + 1) = one 16-key k-tile
    uint32_t ph[NK16][4], pl[NK16][4];
#pragma unroll
    for (int kk = 0; kk < NK16; ++kk) {
      const int n0 = 2 * kk, n1 = 2 * kk + 1;
      am_split2(s[n0][0] * kPS, s[n0][1] * kPS, ph[kk][0], pl[kk][0]);
      am_split2(s[n0][2] * kPS, s[n0][3] * kPS, ph[kk][1], pl[kk][1]);
      if (n1 < NK8) {
        am_split2(s[n1 < NK8 ? n1 : n0][0] * kPS, s[n1 < NK8 ? n1 : n0][1] * kPS, ph[kk][2], pl[kk][2]);
        am_split2(s[n1 < NK8 ? n1 : n0][2] * kPS, s[n1 < NK8 ? n1 : n0][3] * kPS, ph[kk][3], pl[kk][3]);
      } else {
        ph[kk][2] = ph[kk][3] = pl[kk][2] = pl[kk][3] = 0u;
      }
    }
    // ---- O[16 x 64] = P V: B fragment of dh tile nt: (k = key 16 kk + 2 t (+1, +8, +9), n = dh nt * 8 + g)
    const float uo0 = 1.f / (l0 * kPS * kQS), uo1 = 1.f / (l1 * kPS * kQS);
    float* op0 = O + ((size_t)b * N + q0) * ldo + h * DH;
    float* op1 = O + ((size_t)b * N + q1) * ldo + h * DH;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < NK16; ++kk) {
        const int j = kk * 16 + 2 * t;
        const float v00 = __ldg(vrow(j) + nt * 8 + g), v01 = __ldg(vrow(j + 1) + nt * 8 + g);
        const float v10 = __ldg(vrow(j + 8) + nt * 8 + g), v11 = __ldg(vrow(j + 9) + nt * 8 + g);
        uint32_t vh[2], vl[2];
        am_split2(v00 * kQS, v01 * kQS, vh[0], vl[0]);
        am_split2(v10 * kQS, v11 * kQS, vh[1], vl[1]);
        mma16816(o, ph[kk], vl);
        mma16816(o, pl[kk], vh);
        mma16816(o, ph[kk], vh);
      }
      if (q0 < N) *(float2*)(op0 + nt * 8 + 2 * t) = make_float2(o[0] * uo0, o[1] * uo0);
      if (q1 < N) *(float2*)(op1 + nt * 8 + 2 * t) = make_float2(o[2] * uo1, o[3] * uo1);
    }
  }
}

}  // namespace dq
