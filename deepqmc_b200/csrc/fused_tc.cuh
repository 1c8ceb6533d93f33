// Fused Psiformer MLP block of a plain forward (S = 1) on the 5th-gen tensor cores, "3xFP16" operands (see gemm_tcgen05.cuh):
//
//     A  = X + O Wo                      (attention output projection + residual)
//     M1 = tanh(A W1 + b1)
//     X' = A + tanh(M1 W2 + b2)          (reference: gnn/update_features.py:241-286 attention layer with
//                                          conf/ansatz/psiformer.yaml:84-101 MLP; hkext.py:22-113, residual rule :116-137)
//
// for a tile of 128 rows (row = walker x electron) per CTA, persistent over tiles.  The three GEMMs of a tile run back to back
// with every intermediate on chip: A stays in TMEM columns [0, d) as fp32 (it is the residual of the last stage), the operand
// of the next GEMM is written by the epilogue of the previous one straight into the 128-byte-swizzled K-major operand buffer
// in shared memory (split into hi / lo halves), and only O and X are read from / X' written to HBM: 3 x 4 d bytes per row
// instead of 8 x 4 d with one launch per layer.
//
// warp roles (320 threads):
//   warps 0-7  workers : stage the O tile (fp32 -> hi/lo halves), then the three epilogues.  Warp w owns TMEM lanes / tile rows
//                        32 (w % 4) .. +31 and the column half w / 4 (k-blocks {0,1} or {2,3} of the next operand).
//   warp  8    TMA     : streams the pre-split weight planes [d x 64 halves] (hi, lo per k-block) of Wo, W1, W2 through a
//                        3-stage ring; runs ahead across GEMM / tile boundaries.
//   warp  9    MMA     : tcgen05.mma kind::f16, M = 128, N = d: per k-block  a_lo w_hi + a_hi w_hi  (hi plane),  a_hi w_lo  (lo plane).
// shared memory: operand buffer 4 k-blocks x {hi, lo} x 16 KB = 128 KB, weight ring 3 x 32 KB, biases, barriers.
// TMEM (512 columns): [0, d) GEMM 1 accumulator -> A;  [256, 256 + d) accumulator of GEMM 2, then of GEMM 3.
#pragma once
#include <cstdint>

#include "tc_ptx.cuh"

namespace dq {
namespace tc {

constexpr int kMlpThreads = 320;
constexpr int kMlpStages = 3;

struct MlpParams {
  const float* O; int ldo;     // attention output rows [M][d]
  const float* X; int ldx;     // residual stream rows [M][d]
  float* Out; int ldout;       // X' rows [M][d]; may alias O (a tile reads its O rows before it writes them)
  const float* b1; const float* b2;
  int M, d;
  float a_scale;               // power of two applied to every activation operand before the hi / lo split
  float us0, us1, us2;         // accumulator unscale of the three GEMMs: 1 / (a_scale * weight scale)
  int* err_flag;
};

struct MlpSmem {
  static __host__ __device__ int abuf(int kb, int plane) { return (kb * 2 + plane) * 16384; }   // [128 rows][128 B]
  static __host__ __device__ int wring(int s) { return 131072 + s * 32768; }                      // [<= 256 rows][128 B]
  static __host__ __device__ int bias() { return 131072 + kMlpStages * 32768; }                   // b1[256], b2[256]
  static __host__ __device__ int bars() { return bias() + 2048; }
  static __host__ __device__ int total() { return bars() + 256; }
};

// tanh of the plain-forward epilogues (same as gemm_tcgen05.cuh tanh_fwd): absolute error <= ~3e-7
__device__ __forceinline__ float mlp_tanh(float x) {
  const float x2 = x * x;
  const float poly = x + x * x2 * (-0.33333333333f + x2 * (0.13333333333f + x2 * (-0.05396825397f)));
  const float e = ex2_approx(x * 2.8853900817779268f);
  const float big = 1.f - fast_div(2.f, 1.f + e);
  return fabsf(x) < 0.15f ? poly : big;
}

// (x0, x1) -> packed hi halves and packed lo halves, x = hi + lo to 22 significant bits.  hi is formed in fp32 by Veltkamp's
// splitting (c = 8193 x, hi = c - (c - x): x rounded to 11 bits, three full-rate instructions) instead of converting the packed
// half back (F2F conversions issue at a fraction of the FMA rate and made up a fifth of the whole-trunk kernel's stall samples);
// x - hi is exact, both packs are then plain cvt.rn.f16x2.  Values below the normal half range (2^-14 after scaling) keep
// the absolute floor of 2^-25 that the split has anyway.
__device__ __forceinline__ void split_half2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const float c0 = __fmul_rn(x0, 8193.f), c1 = __fmul_rn(x1, 8193.f);
  const float h0 = __fsub_rn(c0, __fsub_rn(c0, x0)), h1 = __fsub_rn(c1, __fsub_rn(c1, x1));
  hi = pack_half2_rn(h0, h1);
  lo = pack_half2_rn(__fsub_rn(x0, h0), __fsub_rn(x1, h1));
}

// 32 consecutive columns [c0, c0 + 32) of tile row `row` (values v, already scaled by a_scale) -> hi / lo halves in the
// K-major operand buffer: k-block c0 / 64, 16-byte chunks 4 (c0 / 32 % 2) .. +3 of the row's 128-byte line.
__device__ __forceinline__ void store_operand_chunk(unsigned char* smem, int row, int c0, const float* v) {
  const int kb = c0 >> 6, cbase = ((c0 >> 5) & 1) * 4;
  unsigned char* ph = smem + MlpSmem::abuf(kb, 0) + (row >> 3) * 1024 + (row & 7) * 128;
  unsigned char* pl = smem + MlpSmem::abuf(kb, 1) + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      split_half2(v[8 * q + 2 * e], v[8 * q + 2 * e + 1], h[e], l[e]);
    }
    const int off = ((cbase + q) ^ (row & 7)) << 4;
    *(uint4*)(ph + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *(uint4*)(pl + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_block_f16_kernel(const __grid_constant__ CUtensorMap wo_hi, const __grid_constant__ CUtensorMap wo_lo,
                     const __grid_constant__ CUtensorMap w1_hi, const __grid_constant__ CUtensorMap w1_lo,
                     const __grid_constant__ CUtensorMap w2_hi, const __grid_constant__ CUtensorMap w2_lo, MlpParams p) {
  DQMC_TC_SMEM(smem);
  if ((smem_u32(smem) & 1023u) != 0u) tc_trap();
  uint64_t* bars = (uint64_t*)(smem + MlpSmem::bars());
  uint64_t* afull = bars;                 // [4] operand k-block kb written (128 worker threads each)
  uint64_t* wfull = bars + 4;             // [kMlpStages] weight plane landed (TMA tx)
  uint64_t* wempty = bars + 4 + kMlpStages;      // [kMlpStages] weight plane consumed (tcgen05.commit)
  uint64_t* accfull = bars + 4 + 2 * kMlpStages;  // accumulator of the current GEMM complete (tcgen05.commit)
  uint64_t* tmemfree = accfull + 1;       // all 256 workers are done with the tile's TMEM contents
  uint32_t* tmem_slot = (uint32_t*)(tmemfree + 1);
  float* sb1 = (float*)(smem + MlpSmem::bias());
  float* sb2 = sb1 + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = p.d, KB = d / 64;
  const int MT = (p.M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) mbar_init(&afull[k], 128);
    for (int s = 0; s < kMlpStages; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
    mbar_init(accfull, 1);
    mbar_init(tmemfree, 256);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    sb1[i] = i < d ? p.b1[i] : 0.f;
    sb2[i] = i < d ? p.b2[i] : 0.f;
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&wo_hi); tma_prefetch_desc(&wo_lo); tma_prefetch_desc(&w1_hi);
    tma_prefetch_desc(&w1_lo); tma_prefetch_desc(&w2_hi); tma_prefetch_desc(&w2_lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ===================== weight planes: Wo, W1, W2 of every tile, k-block by k-block, hi then lo ==========
    if (elect_one()) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < MT; tile += gridDim.x) {
        for (int g = 0; g < 3; ++g) {
          const CUtensorMap* mh = g == 0 ? &wo_hi : (g == 1 ? &w1_hi : &w2_hi);
          const CUtensorMap* ml = g == 0 ? &wo_lo : (g == 1 ? &w1_lo : &w2_lo);
          for (int kb = 0; kb < KB; ++kb)
            for (int plane = 0; plane < 2; ++plane, ++it) {
              const int s = it % kMlpStages;
              const uint32_t ph = (it / kMlpStages) & 1;
              mbar_wait(&wempty[s], ph ^ 1, p.err_flag);
              mbar_expect_tx(&wfull[s], (uint32_t)d * 128u);
              tma_load_2d(plane == 0 ? mh : ml, &wfull[s], smem + MlpSmem::wring(s), kb * 64, 0);
            }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =========================================================================
    const uint32_t idesc = make_idesc_f16(128, d);
    uint32_t it = 0, tcount = 0, gcount = 0;
    for (int tile = blockIdx.x; tile < MT; tile += gridDim.x, ++tcount) {
      mbar_wait(tmemfree, (tcount & 1) ^ 1, p.err_flag);  // previous tile's epilogues have drained TMEM
      tc_fence_after();
      for (int g = 0; g < 3; ++g, ++gcount) {
        const uint32_t d_tmem = tmem_base + (g == 0 ? 0u : 256u);
        const uint32_t aph = gcount & 1;  // afull[kb] completes once per GEMM
        if (g > 0)  // the operand comes from the previous epilogue, which also reads the accumulator this GEMM overwrites
          for (int kb = 0; kb < KB; ++kb) mbar_wait(&afull[kb], aph, p.err_flag);
        for (int kb = 0; kb < KB; ++kb) {
          if (g == 0) mbar_wait(&afull[kb], aph, p.err_flag);
          const uint32_t ah = smem_u32(smem + MlpSmem::abuf(kb, 0)), al = smem_u32(smem + MlpSmem::abuf(kb, 1));
          for (int plane = 0; plane < 2; ++plane, ++it) {
            const int s = it % kMlpStages;
            const uint32_t ph = (it / kMlpStages) & 1;
            mbar_wait(&wfull[s], ph, p.err_flag);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t w = smem_u32(smem + MlpSmem::wring(s));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t ko = k * 32;  // 16 halves per instruction
                if (plane == 0) {
                  umma_f16(d_tmem, make_desc(al + ko), make_desc(w + ko), idesc, (kb | k) ? 1u : 0u);
                  umma_f16(d_tmem, make_desc(ah + ko), make_desc(w + ko), idesc, 1u);
                } else {
                  umma_f16(d_tmem, make_desc(ah + ko), make_desc(w + ko), idesc, 1u);
                }
              }
              umma_commit(&wempty[s]);
              if (kb == KB - 1 && plane == 1) umma_commit(accfull);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ===================== workers: warps 0-7 ===================================================================
    const int q4 = warp & 3, half = warp >> 2;       // TMEM lane quarter / tile rows 32 q4 .. +31; column half
    const int trow = 32 * q4 + lane;                  // this thread's tile row (= TMEM lane)
    const uint32_t tlane = (uint32_t)(32 * q4) << 16;
    const int nchunk = d / 32;                        // 32-column chunks per row
    const int c_lo = half * (nchunk / 2), c_hi = (half + 1) * (nchunk / 2);  // this warp's chunks = its KB / 2 k-blocks
    uint32_t gcount = 0;
    for (int tile = blockIdx.x; tile < MT; tile += gridDim.x) {
      const int row = tile * 128 + trow;
      const bool valid = row < p.M;
      // ---- stage the O tile: coalesced float4 loads (8 lanes x 16 B per row), hi / lo split, k-blocks of this warp's half.
      // Rows of one instruction differ in bit 2 within each half-warp: the 8-byte stores are then conflict-free.
      {
        const int cq = lane & 7, rs = lane >> 3;
        for (int kb = half * (KB / 2); kb < (half + 1) * (KB / 2); ++kb) {  // KB is 2 or 4 (d = 128 | 256)
          unsigned char* ph = smem + MlpSmem::abuf(kb, 0);
          unsigned char* pl = smem + MlpSmem::abuf(kb, 1);
#pragma unroll 4
          for (int j = 0; j < 16; ++j) {                 // 8 row groups x 2 halves of the 64-float k-block
            const int seg = j & 1, jj = j >> 1;
            const int r = 32 * q4 + (jj >> 1) * 8 + 4 * (rs & 1) + (rs >> 1) + 2 * (jj & 1);
            const int grow = tile * 128 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (grow < p.M) v = __ldg((const float4*)(p.O + (size_t)grow * p.ldo + kb * 64 + seg * 32 + cq * 4));
            const float sc = p.a_scale;
            const float x0 = v.x * sc, x1 = v.y * sc, x2 = v.z * sc, x3 = v.w * sc;
            const uint32_t h01 = pack_half2_rn(x0, x1), h23 = pack_half2_rn(x2, x3);
            const uint32_t l01 = pack_half2_rn(x0 - half_bits_to_float(h01 & 0xFFFFu), x1 - half_bits_to_float(h01 >> 16));
            const uint32_t l23 = pack_half2_rn(x2 - half_bits_to_float(h23 & 0xFFFFu), x3 - half_bits_to_float(h23 >> 16));
            const int c16 = seg * 4 + (cq >> 1);
            const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4) + (cq & 1) * 8;
            *(uint2*)(ph + off) = make_uint2(h01, h23);
            *(uint2*)(pl + off) = make_uint2(l01, l23);
          }
          fence_proxy_async();
          mbar_arrive(&afull[kb]);
        }
      }
      // ---- epilogue 1: A = X + O Wo -> TMEM [0, d) (fp32) and the operand buffer
      mbar_wait(accfull, gcount & 1, p.err_flag);
      ++gcount;
      tc_fence_after();
      for (int c = c_lo; c < c_hi; ++c) {
        float4 xr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          xr[i] = valid ? __ldg((const float4*)(p.X + (size_t)row * p.ldx + c * 32 + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t v[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(c * 32), v);
        tmem_ld_wait();
        float a[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[4 * i] = xr[i].x + __uint_as_float(v[4 * i]) * p.us0;
          a[4 * i + 1] = xr[i].y + __uint_as_float(v[4 * i + 1]) * p.us0;
          a[4 * i + 2] = xr[i].z + __uint_as_float(v[4 * i + 2]) * p.us0;
          a[4 * i + 3] = xr[i].w + __uint_as_float(v[4 * i + 3]) * p.us0;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(a[i]);
        tmem_st32(tmem_base + tlane + (uint32_t)(c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] *= p.a_scale;
        store_operand_chunk(smem, trow, c * 32, a);
        if (c & 1) {  // second chunk of a k-block: hand it to the MMA warp
          tmem_st_wait();
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&afull[c >> 1]);
        }
      }
      // ---- epilogue 2: M1 = tanh(A W1 + b1) -> operand buffer
      mbar_wait(accfull, gcount & 1, p.err_flag);
      ++gcount;
      tc_fence_after();
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(256 + c * 32), v);
        tmem_ld_wait();
        float a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = mlp_tanh(__uint_as_float(v[i]) * p.us1 + sb1[c * 32 + i]) * p.a_scale;
        store_operand_chunk(smem, trow, c * 32, a);
        if (c & 1) {
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&afull[c >> 1]);
        }
      }
      // ---- epilogue 3: X' = A + tanh(M1 W2 + b2) -> HBM
      mbar_wait(accfull, gcount & 1, p.err_flag);
      ++gcount;
      tc_fence_after();
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[32], r[32];
        tmem_ld32(tmem_base + tlane + (uint32_t)(256 + c * 32), v);
        tmem_ld32(tmem_base + tlane + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (valid) {
          float* op = p.Out + (size_t)row * p.ldout + c * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 o;
            o.x = __uint_as_float(r[4 * i]) + mlp_tanh(__uint_as_float(v[4 * i]) * p.us2 + sb2[c * 32 + 4 * i]);
            o.y = __uint_as_float(r[4 * i + 1]) + mlp_tanh(__uint_as_float(v[4 * i + 1]) * p.us2 + sb2[c * 32 + 4 * i + 1]);
            o.z = __uint_as_float(r[4 * i + 2]) + mlp_tanh(__uint_as_float(v[4 * i + 2]) * p.us2 + sb2[c * 32 + 4 * i + 2]);
            o.w = __uint_as_float(r[4 * i + 3]) + mlp_tanh(__uint_as_float(v[4 * i + 3]) * p.us2 + sb2[c * 32 + 4 * i + 3]);
            *(float4*)(op + 4 * i) = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmemfree);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tc
}  // namespace dq
