// Functional CPU model of the sm_100a primitives the tensor-core kernels use -- mbarrier, TMA tiled loads (2-D, 128-byte
// swizzle), tcgen05.{alloc,mma,commit,ld,st}, TMEM -- DEVELOPMENT TOOL ONLY (never linked into libdqmc_b200.so).
//
// The kernels in deepqmc_b200/csrc/gemm_tcgen05.cuh / fused_tc.cuh are written against the small wrapper API of tc_ptx.cuh;
// with -DDQMC_EMU that header includes this file instead of the inline-PTX versions, so the SAME kernel source (warp roles,
// barrier protocol, descriptor arithmetic, swizzled shared-memory addressing, epilogue indexing) runs on the fiber emulator
// of cuda_emu.h.  The model is the documented behaviour the hardware-verified 3xTF32 GEMM already relies on:
//   * mbarrier: phase completes when the pending-arrival count and the transaction-byte count are both zero;
//     try_wait.parity(P) succeeds once the phase of parity P has completed (a fresh barrier: parity 1 succeeds at once).
//   * TMA 2-D tiled load, SWIZZLE_128B, box inner extent = 128 bytes: box row i lands at dst + 128 i with its 16-byte chunks
//     permuted by XOR with (i mod 8) -- i.e. shared address bits [4,7) ^= bits [7,10); out-of-bounds elements are zero.
//   * shared-memory matrix descriptor (K-major, SWIZZLE_128B, SBO = 1024 B): element (row r, byte b of the 32-byte k-step)
//     is read from  start + (r / 8) SBO + (r % 8) 128 + b  with the same XOR applied to the final address.
//   * tcgen05.mma cta_group::1, M = 128: D[lane m][column c0 + n] (+)= sum_k A[m][k] B[n][k], fp32 accumulation;
//     kind::tf32 reads fp32 words and drops the low 13 mantissa bits, kind::f16 reads IEEE halves (8 / 16 k per instruction).
//   * tcgen05.ld/st .32x32b.x32: warp w touches TMEM lanes 32 (w mod 4) + lane, 32 consecutive columns.
// MMAs execute synchronously at issue, so tcgen05.commit arrives immediately (asynchrony of the real pipe is not modelled;
// the mbarrier protocol still has to be consistent or the fibers dead-lock, which the bounded wait reports).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "cuda_emu.h"

// ---- minimal stand-ins for the driver types --------------------------------------------------------------------------
struct CUtensorMap {
  const void* base;
  uint64_t dim0, dim1;      // elements (inner, outer)
  uint64_t stride1_bytes;   // outer stride
  uint32_t box0, box1;      // elements
  uint32_t elem_bytes;
  uint32_t swizzle;         // 0 none, 3 = 128B
};
#define __grid_constant__
#define DQMC_TC_SMEM(name) unsigned char* name = dq::tc::emu_tc::smem_base()

namespace dq {
namespace tc {

namespace emu_tc {
struct MBar { int init = 0; int pending = 0; long tx = 0; int phase = 0; };
struct St {
  std::map<uintptr_t, MBar> bars;                  // keyed by host address of the 8-byte barrier object
  std::vector<uint32_t> tmem = std::vector<uint32_t>(128 * 512, 0xFFFFFFFFu);
  std::map<int, std::pair<long, int>> named;       // named barrier id -> (generation, count)
};
inline St& st() { static thread_local St s; return s; }
inline unsigned char* smem_base() { return (unsigned char*)(((uintptr_t)emu::S().dyn_smem.data() + 1023) & ~(uintptr_t)1023); }
inline unsigned char* sptr(uint32_t a) { return smem_base() + a; }
inline uint32_t swz128(uint32_t a) { return a ^ (((a >> 7) & 7u) << 4); }
inline float half_to_float(uint16_t h) {
  const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
  float v;
  if (e == 0) v = std::ldexp((float)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = std::ldexp((float)(m | 1024u), (int)e - 25);
  return s ? -v : v;
}
inline uint16_t float_to_half_rn(float f) {  // round to nearest even, IEEE binary16 (subnormals, overflow -> inf)
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
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;  // may carry into the exponent: still correct
    return (uint16_t)(sign | h);
  }
  if (e < -25) return (uint16_t)sign;
  m |= 0x800000u;
  const int shift = -e - 14 + 13;  // 14 .. 24
  uint32_t h = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem > half || (rem == half && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}
}  // namespace emu_tc

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu_tc::smem_base()); }

// ---- mbarrier --------------------------------------------------------------------------------------------------------
inline emu_tc::MBar& mb(uint64_t* bar) { return emu_tc::st().bars[(uintptr_t)bar]; }
inline void mb_check(emu_tc::MBar& b) {
  if (b.pending < 0) { std::fprintf(stderr, "tcgen05_emu: mbarrier over-arrived\n"); std::abort(); }
  if (b.pending == 0 && b.tx == 0) { b.phase ^= 1; b.pending = b.init; }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  emu_tc::MBar& b = mb(bar);
  b.init = b.pending = (int)count; b.tx = 0; b.phase = 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { emu_tc::MBar& b = mb(bar); --b.pending; mb_check(b); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  emu_tc::MBar& b = mb(bar);
  b.tx += bytes; --b.pending; mb_check(b);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) { emu_tc::MBar& b = mb(bar); b.tx -= bytes; mb_check(b); }
inline std::vector<long>& wait_trace() { static thread_local std::vector<long> v(2048, -1); return v; }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  long spins = 0;
  wait_trace()[threadIdx.x] = (long)smem_u32(bar) * 2 + (parity & 1);
  while ((uint32_t)mb(bar).phase == (parity & 1u)) {
    emu::yield();
    if (++spins > 2000000) {
      if (err_flag) *err_flag = 1;
      std::fprintf(stderr, "tcgen05_emu: mbarrier wait timed out (thread %d, barrier at shared offset %u, parity %u, phase %d pending %d tx %ld): protocol dead-lock\n",
                   (int)threadIdx.x, smem_u32(bar), parity, mb(bar).phase, mb(bar).pending, mb(bar).tx);
      for (int i = 0; i < (int)blockDim.x; i += 32)  // where every warp's lane 0 waited last (shared offset, parity)
        std::fprintf(stderr, "  warp %d lane0: last mbarrier wait at offset %ld parity %ld\n", i / 32, wait_trace()[i] / 2, wait_trace()[i] & 1);
      std::abort();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap*) {}

// named barrier among `count` threads (bar.sync id, count)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  auto& nb = emu_tc::st().named[id];
  const long gen = nb.first;
  if (++nb.second == count) { nb.second = 0; ++nb.first; return; }
  while (emu_tc::st().named[id].first == gen) emu::yield();
}

// ---- TMA ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  const CUtensorMap& m = *map;
  const uint32_t row_bytes = m.box0 * m.elem_bytes;
  if (m.swizzle == 3 && row_bytes != 128) { std::fprintf(stderr, "tcgen05_emu: SWIZZLE_128B box must be 128 bytes wide\n"); std::abort(); }
  const uint32_t d0 = smem_u32(dst);
  if (m.swizzle == 3 && (d0 & 1023u)) { std::fprintf(stderr, "tcgen05_emu: swizzled TMA destination must be 1024-byte aligned\n"); std::abort(); }
  for (uint32_t i = 0; i < m.box1; ++i) {
    const int64_t gy = (int64_t)y + i;
    for (uint32_t b = 0; b < row_bytes; ++b) {
      const int64_t gx = (int64_t)x + b / m.elem_bytes;
      unsigned char v = 0;
      if (gy >= 0 && gy < (int64_t)m.dim1 && gx >= 0 && gx < (int64_t)m.dim0)
        v = ((const unsigned char*)m.base)[gy * m.stride1_bytes + gx * m.elem_bytes + b % m.elem_bytes];
      uint32_t a = d0 + i * row_bytes + b;
      if (m.swizzle == 3) a = emu_tc::swz128(a);
      *emu_tc::sptr(a) = v;
    }
  }
  mbar_complete_tx(bar, m.box1 * row_bytes);
}

// ---- descriptors ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
inline void umma_any(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, bool f16) {
  const int N = (int)((idesc >> 17) & 63u) << 3, M = (int)((idesc >> 24) & 31u) << 4;
  const uint32_t afmt = (idesc >> 7) & 7u, bfmt = (idesc >> 10) & 7u;
  if (M != 128 || N < 16 || N > 256 || (N & 15)) { std::fprintf(stderr, "tcgen05_emu: unsupported MMA shape %dx%d\n", M, N); std::abort(); }
  if ((f16 && (afmt != 0 || bfmt != 0)) || (!f16 && (afmt != 2 || bfmt != 2))) {
    std::fprintf(stderr, "tcgen05_emu: instruction descriptor formats do not match the MMA kind\n"); std::abort();
  }
  const uint32_t lta = (uint32_t)((adesc >> 61) & 7u), ltb = (uint32_t)((bdesc >> 61) & 7u);
  if ((lta != 2 && lta != 0) || (ltb != 2 && ltb != 0)) { std::fprintf(stderr, "tcgen05_emu: only SWIZZLE_128B / unswizzled descriptors\n"); std::abort(); }
  const uint32_t a0 = (uint32_t)(adesc & 0x3FFF) << 4, b0 = (uint32_t)(bdesc & 0x3FFF) << 4;
  const uint32_t sboa = (uint32_t)((adesc >> 32) & 0x3FFF) << 4, sbob = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  const uint32_t lboa = (uint32_t)((adesc >> 16) & 0x3FFF) << 4, lbob = (uint32_t)((bdesc >> 16) & 0x3FFF) << 4;
  const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xFFFFu;
  if (lane0 != 0 || col0 + N > 512) { std::fprintf(stderr, "tcgen05_emu: accumulator outside TMEM\n"); std::abort(); }
  const int KE = f16 ? 16 : 8;  // 32 bytes of K per instruction
  auto rd = [&](uint32_t base, uint32_t sbo, uint32_t lbo, uint32_t lt, int r, int k) -> float {
    const uint32_t eb = f16 ? 2u : 4u;
    // unswizzled: 16-byte chunk c of row r at (r / 8) SBO + c LBO + (r % 8) 16
    const uint32_t kbyte = (uint32_t)k * eb;
    const uint32_t a = lt == 2 ? emu_tc::swz128(base + (uint32_t)(r >> 3) * sbo + (uint32_t)(r & 7) * 128u + kbyte)
                               : base + (uint32_t)(r >> 3) * sbo + (kbyte >> 4) * lbo + (uint32_t)(r & 7) * 16u + (kbyte & 15u);
    if (f16) { uint16_t h; std::memcpy(&h, emu_tc::sptr(a), 2); return emu_tc::half_to_float(h); }
    uint32_t w; std::memcpy(&w, emu_tc::sptr(a), 4); w &= 0xFFFFE000u;
    float v; std::memcpy(&v, &w, 4); return v;
  };
  std::vector<float> Bt((size_t)N * KE);
  for (int n = 0; n < N; ++n) for (int k = 0; k < KE; ++k) Bt[(size_t)n * KE + k] = rd(b0, sbob, lbob, ltb, n, k);
  auto& T = emu_tc::st().tmem;
  for (int m = 0; m < M; ++m) {
    float ar[16];
    for (int k = 0; k < KE; ++k) ar[k] = rd(a0, sboa, lboa, lta, m, k);
    for (int n = 0; n < N; ++n) {
      double s = 0.0;
      for (int k = 0; k < KE; ++k) s += (double)ar[k] * (double)Bt[(size_t)n * KE + k];
      uint32_t& cell = T[(size_t)m * 512 + col0 + n];
      float d;
      std::memcpy(&d, &cell, 4);
      d = acc ? (float)((double)d + s) : (float)s;
      std::memcpy(&cell, &d, 4);
    }
  }
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  umma_any(tmem_d, adesc, bdesc, idesc, acc, false);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  umma_any(tmem_d, adesc, bdesc, idesc, acc, true);
}
// A operand in tensor memory: A[m][k] = half (k % 2) of cell (lane m, column a_col + k / 2)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  const int N = (int)((idesc >> 17) & 63u) << 3, M = (int)((idesc >> 24) & 31u) << 4;
  if (M != 128 || N < 16 || N > 256 || (N & 15)) { std::fprintf(stderr, "tcgen05_emu: unsupported TS MMA shape %dx%d\n", M, N); std::abort(); }
  if (((idesc >> 7) & 7u) != 0 || ((idesc >> 10) & 7u) != 0) { std::fprintf(stderr, "tcgen05_emu: TS MMA expects f16 formats\n"); std::abort(); }
  if (((bdesc >> 61) & 7u) != 2) { std::fprintf(stderr, "tcgen05_emu: only SWIZZLE_128B descriptors\n"); std::abort(); }
  const uint32_t b0 = (uint32_t)(bdesc & 0x3FFF) << 4, sbob = (uint32_t)((bdesc >> 32) & 0x3FFF) << 4;
  const uint32_t col0 = tmem_d & 0xFFFFu, acol = tmem_a & 0xFFFFu;
  if ((tmem_d >> 16) != 0 || (tmem_a >> 16) != 0 || col0 + N > 512 || acol + 8 > 512) { std::fprintf(stderr, "tcgen05_emu: TS MMA operand outside TMEM\n"); std::abort(); }
  if (!(acol + 8 <= col0 || col0 + (uint32_t)N <= acol)) { std::fprintf(stderr, "tcgen05_emu: TS MMA A and D overlap\n"); std::abort(); }
  auto& T = emu_tc::st().tmem;
  std::vector<float> Bt((size_t)N * 16);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) {
      const uint32_t a = emu_tc::swz128(b0 + (uint32_t)(n >> 3) * sbob + (uint32_t)(n & 7) * 128u + (uint32_t)k * 2u);
      uint16_t h; std::memcpy(&h, emu_tc::sptr(a), 2);
      Bt[(size_t)n * 16 + k] = emu_tc::half_to_float(h);
    }
  for (int m = 0; m < M; ++m) {
    float ar[16];
    for (int k = 0; k < 16; ++k) {
      const uint32_t cell = T[(size_t)m * 512 + acol + k / 2];
      ar[k] = emu_tc::half_to_float((uint16_t)((k & 1) ? cell >> 16 : cell & 0xFFFFu));
    }
    for (int n = 0; n < N; ++n) {
      double s = 0.0;
      for (int k = 0; k < 16; ++k) s += (double)ar[k] * (double)Bt[(size_t)n * 16 + k];
      uint32_t& cell = T[(size_t)m * 512 + col0 + n];
      float d; std::memcpy(&d, &cell, 4);
      d = acc ? (float)((double)d + s) : (float)s;
      std::memcpy(&cell, &d, 4);
    }
  }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15u) || ((uintptr_t)src & 15u) || (smem_u32(dst) & 15u)) { std::fprintf(stderr, "tcgen05_emu: bulk copy alignment\n"); std::abort(); }
  std::memcpy(dst, src, bytes);
  mbar_complete_tx(bar, bytes);
}
__device__ __forceinline__ void fence_proxy_async_all() {}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { mbar_arrive(bar); }

// ---- TMEM ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {  // whole warp calls; lane 0 writes the base (0)
  if (cols != 32 && cols != 64 && cols != 128 && cols != 256 && cols != 512) { std::fprintf(stderr, "tcgen05_emu: bad TMEM allocation\n"); std::abort(); }
  if ((threadIdx.x & 31) == 0) {
    std::fill(emu_tc::st().tmem.begin(), emu_tc::st().tmem.end(), 0xFFFFFFFFu);  // fresh TMEM holds garbage (NaN)
    *slot = 0;
  }
}
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
inline uint32_t* tmem_cell(uint32_t taddr, int i) {
  const uint32_t lane_base = taddr >> 16, col = (taddr & 0xFFFFu) + (uint32_t)i;
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
  if (lane_base != (uint32_t)(32 * (warp & 3))) { std::fprintf(stderr, "tcgen05_emu: warp %d may only touch TMEM lanes %d..%d (asked %u)\n", warp, 32 * (warp & 3), 32 * (warp & 3) + 31, lane_base); std::abort(); }
  if (col >= 512) { std::fprintf(stderr, "tcgen05_emu: TMEM column out of range\n"); std::abort(); }
  return &emu_tc::st().tmem[(size_t)(lane_base + lane) * 512 + col];
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) { for (int i = 0; i < 32; ++i) v[i] = *tmem_cell(taddr, i); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) { for (int i = 0; i < 32; ++i) *tmem_cell(taddr, i) = v[i]; }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) { for (int i = 0; i < 16; ++i) *tmem_cell(taddr, i) = v[i]; }
__device__ __forceinline__ void tmem_ld_wait() {}
__device__ __forceinline__ void tmem_st_wait() {}

// ---- scalar helpers with a PTX counterpart ---------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) { return std::exp2(x); }
__device__ __forceinline__ bool elect_one() { return (threadIdx.x & 31) == 0; }
__device__ __forceinline__ float fast_div(float a, float b) { return a / b; }
__device__ __forceinline__ uint32_t pack_half2_rn(float lo, float hi) {  // cvt.rn.f16x2.f32: low half <- lo
  return (uint32_t)emu_tc::float_to_half_rn(lo) | ((uint32_t)emu_tc::float_to_half_rn(hi) << 16);
}
__device__ __forceinline__ float half_bits_to_float(uint32_t h16) { return emu_tc::half_to_float((uint16_t)h16); }
__device__ __forceinline__ void tc_trap() { std::fprintf(stderr, "tcgen05_emu: trap\n"); std::abort(); }
__device__ __forceinline__ long long tc_clock() { return 0; }
__device__ __forceinline__ float tf32_rna(float x) {  // cvt.rna.tf32.f32: nearest, ties away from zero
  uint32_t b;
  std::memcpy(&b, &x, 4);
  if (((b >> 23) & 255u) != 255u) b += 0x1000u;
  b &= 0xFFFFE000u;
  float y;
  std::memcpy(&y, &b, 4);
  return y;
}

// 2-CTA (cta_group::2) forms are not modelled: the CTA-pair kernel variant is never launched in emulator builds
inline void no_pair() { std::fprintf(stderr, "tcgen05_emu: cta_group::2 is not modelled\n"); std::abort(); }
__device__ __forceinline__ uint32_t cluster_ctarank() { return 0; }
__device__ __forceinline__ void cluster_sync_all() { no_pair(); }
__device__ __forceinline__ uint32_t mapa_u32(const void*, uint32_t) { no_pair(); return 0; }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t) { no_pair(); }
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap*, uint32_t, void*, int, int) { no_pair(); }
__device__ __forceinline__ void umma_tf32_2sm(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { no_pair(); }
__device__ __forceinline__ void umma_commit_2sm(uint64_t*) { no_pair(); }
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t*, uint32_t) { no_pair(); }
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t, uint32_t) { no_pair(); }

inline int make_kmajor_map(CUtensorMap* map, const void* base, int elem_bytes, int rows, int K, int box_k, int box_rows) {
  if (box_k * elem_bytes != 128 || (elem_bytes != 4 && elem_bytes != 2)) return 3;
  if (box_rows > 256 || box_rows < 1) return 2;  // boxDim <= 256 per dimension
  map->base = base; map->dim0 = (uint64_t)K; map->dim1 = (uint64_t)rows; map->stride1_bytes = (uint64_t)K * elem_bytes;
  map->box0 = (uint32_t)box_k; map->box1 = (uint32_t)box_rows; map->elem_bytes = (uint32_t)elem_bytes; map->swizzle = 3;
  return 0;
}

}  // namespace tc
}  // namespace dq
