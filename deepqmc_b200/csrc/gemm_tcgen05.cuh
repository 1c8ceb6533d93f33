// fp32-accurate dense-layer GEMM on the 5th-gen tensor cores (sm_100a): 3xTF32 split
//   C[row(m), :] = (Res) + A[row(m), :] @ W + (bias on value rows),  A, W, C fp32 in HBM.
//
//   a = a_hi + a_lo, w = w_hi + w_lo with *_hi = rna_tf32(.), *_lo = rna_tf32(. - *_hi) (both exactly representable in TF32);
//   acc(fp32, TMEM) = a_hi w_lo + a_lo w_hi + a_hi w_hi        -> ~2^-21 relative per product,
//   i.e. the accuracy class the reference demands (jax_default_matmul_precision='highest',
//   NVIDIA_TF32_OVERRIDE=0: src/deepqmc/__init__.py:9-34) at 1/3 of the TF32 tensor peak.
//
// F16 = true (plain forwards, S = 1): the same kernel with IEEE-half operands -- "3xFP16":
//   a 2^ea = a_hi + a_lo, w 2^ew = w_hi + w_lo with *_hi = rn_half(.), *_lo = rn_half(. - *_hi): 22 significant bits,
//   products exact in the fp32 accumulator, power-of-two scales undone in the epilogue (exact).  kind::f16 runs at twice
//   the kind::tf32 rate and a 128-byte swizzle row holds 64 k-values instead of 32, so a shared-memory stage covers twice
//   the K extent.  Halves below 2^-14 are subnormal (absolute spacing 2^-24): the scales keep O(1) activations and the
//   weights of a layer far above that, the absolute floor is ~2^-25 / 2^ea per activation.  Only plain forwards use it
//   (values bounded by construction: residual stream, tanh outputs, attention averages); the forward-Laplacian rows
//   (derivative slots of unbounded dynamic range) stay on 3xTF32.
//
// Persistent warp-specialised kernel, one CTA per SM, tile = 128 rows x BN columns x K:
//   warps 0-3  epilogue      : TMEM -> registers (tcgen05.ld) -> +bias/+residual -> global
//   warps 4-7  A producers   : global fp32 rows -> split hi/lo -> 128B-swizzled K-major smem
//   warp  8    W producer    : TMA (cp.async.bulk.tensor) of the pre-split W^T hi/lo tiles
//   warp  9    MMA issuer    : tcgen05.mma.kind::tf32 (one elected lane), commits to mbarriers
// smem ring of kStages k-blocks (32 fp32 = one 128B swizzle row), 2 accumulators in TMEM so the
// epilogue of tile i overlaps the main loop of tile i+1.  W^T (N x K, K contiguous) is split into
// hi/lo once per parameter upload (engine.cu).
#pragma once
#include <cstdint>

#include "tc_ptx.cuh"

namespace dq {
namespace tc {

constexpr int kBM = 128;          // rows per tile (UMMA_M)
constexpr int kBK = 32;           // fp32 per k-block = 128 bytes = one swizzle row
constexpr int kStages = 2;
constexpr int kUmmaK = 8;         // tf32: 32 bytes per MMA k-step
constexpr int kThreads = 320;
constexpr uint32_t kTmemCols = 512;

struct Params {
  const float* A; int lda;
  const float* bias;
  const float* Res; int ldr;
  float* C; int ldc;
  int M, N, K;
  int S;
  int sliced, Nel, z_split;
  int BN;            // 64 | 128 | 256
  int act;           // 0: none, 1: tanh with forward-Laplacian propagation fused into the epilogue
  int rpt;           // rows per tile (<= 128): G*S for act = 1 so that slot groups never straddle tiles
  int* err_flag;     // device int: set to non-zero if a barrier wait times out
  float a_scale;     // F16: activations are multiplied by this power of two before the split ...
  float unscale;     // ... and the accumulator by 2^-(ea + ew) in the epilogue
};

// tanh for the plain-forward epilogue: odd polynomial below |x| = 0.15, 1 - 2 / (1 + e^{2x})
// (ex2.approx + fast division) above; absolute error <= ~3e-7.  The forward-Laplacian epilogue
// keeps tanhf (derivative slots amplify the error); plain forwards only feed log|psi| ratios.
__device__ __forceinline__ float tanh_fwd(float x) {
  const float x2 = x * x;
  const float poly = x + x * x2 * (-0.33333333333f + x2 * (0.13333333333f + x2 * (-0.05396825397f)));
  const float e = ex2_approx(x * 2.8853900817779268f);
  const float big = 1.f - fast_div(2.f, 1.f + e);
  return fabsf(x) < 0.15f ? poly : big;
}

__device__ __forceinline__ size_t phys_row(const Params& p, int m, int z) {
  if (!p.sliced) return (size_t)m;
  int b = m / p.S, s = m % p.S;
  return ((size_t)b * p.Nel + z) * p.S + s;
}

// TWO = CTA pair (cta_group::2): every CTA stages its own 128 activation rows and HALF of the weight tile, so a
// stage is 64 KB instead of 96 KB and the ring gets a third stage.
template <bool TWO>
struct SmemLayoutT {
  static constexpr int kSt = TWO ? 3 : 2;
  // byte offsets from the 1024B-aligned base
  static __host__ __device__ int wrows(int BN) { return TWO ? BN / 2 : BN; }
  static __host__ __device__ int a_hi(int s, int BN) { return s * stage_bytes(BN); }
  static __host__ __device__ int a_lo(int s, int BN) { return s * stage_bytes(BN) + kBM * 128; }
  static __host__ __device__ int w_hi(int s, int BN) { return s * stage_bytes(BN) + 2 * kBM * 128; }
  static __host__ __device__ int w_lo(int s, int BN) { return s * stage_bytes(BN) + 2 * kBM * 128 + wrows(BN) * 128; }
  static __host__ __device__ int stage_bytes(int BN) { return 2 * kBM * 128 + 2 * wrows(BN) * 128; }
  static __host__ __device__ int bars(int BN) { return kSt * stage_bytes(BN); }
  static __host__ __device__ int epi(int BN) { return bars(BN) + 256; }              // 4 warps x 32 x 36 floats
  static __host__ __device__ int side(int BN) { return epi(BN) + 4 * 32 * 36 * 4 + 4 * 32 * 8; }  // y', y'', sum z_t^2: [3][32 groups][32]
  static __host__ __device__ int sbias(int BN) { return side(BN) + 3 * 32 * 32 * 4; }  // bias of the tile's BN columns
  static __host__ __device__ int total(int BN) { return sbias(BN) + 256 * 4 + 64; }
};
using SmemLayout = SmemLayoutT<false>;

template <bool TWO, bool F16 = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm3xtf32_kernel(const __grid_constant__ CUtensorMap map_hi0, const __grid_constant__ CUtensorMap map_lo0,
                  const __grid_constant__ CUtensorMap map_hi1, const __grid_constant__ CUtensorMap map_lo1, Params p) {
  // 1024-byte aligned dynamic shared memory (SWIZZLE_128B atoms).  No integer round-up of the
  // pointer: that would drop the shared address space and turn every access into a generic LD/ST.
  DQMC_TC_SMEM(smem);
  if ((smem_u32(smem) & 1023u) != 0u) tc_trap();
  using SmemLayout = SmemLayoutT<TWO>;
  constexpr int kStages = SmemLayout::kSt;
  const int BN = p.BN;
  uint64_t* bars = (uint64_t*)(smem + SmemLayout::bars(BN));
  uint64_t* full_a = bars;                 // [kStages] count 128 producer threads per CTA (pair: both CTAs -> the leader's)
  uint64_t* full_w = bars + kStages;       // [kStages] count 1 + tx bytes (pair: both halves land on the leader's)
  uint64_t* empty = bars + 2 * kStages;    // [kStages] count 1 (tcgen05.commit; pair: multicast to both CTAs)
  uint64_t* tmem_full = bars + 3 * kStages;       // [2] count 1 (commit)
  uint64_t* tmem_empty = bars + 3 * kStages + 2;  // [2] count 128 epilogue threads per CTA (pair: leader's collects both)
  uint32_t* tmem_base_slot = (uint32_t*)(bars + 3 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RPT = p.rpt;
  const int MT = (p.M + RPT - 1) / RPT, NT = (p.N + BN - 1) / BN;
  const int Z = p.sliced ? p.Nel : 1;
  static_assert(!(TWO && F16), "the half-precision variant is single-CTA");
  // k-blocks of 32 floats as the producers see them; F16: two of them (64 halves = one 128-byte swizzle row) per stage
  const int KB = p.K / kBK;
  constexpr int kSub = F16 ? 2 : 1;
  // work distribution: single CTA: tile = (z, mt, nt); pair: the two CTAs of a cluster take the M tiles 2 mt2 + rank
  // of a pair tile (z, mt2, nt) -- an M tile index >= MT simply has no valid rows
  const uint32_t crank = TWO ? cluster_ctarank() : 0u;
  const bool leader = crank == 0u;
  const int MTX = TWO ? (MT + 1) / 2 : MT;
  const int n_tiles = Z * MTX * NT;
  const int tile0 = TWO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstep = TWO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr uint32_t kPer = TWO ? 256u : 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_a[s], kPer * kSub);
      mbar_init(&full_w[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kPer);
    }
    fence_barrier_init();
  }
  if (warp == 9) {  // TMEM allocation (whole warp), address lands in shared memory
    if constexpr (TWO) tmem_alloc_2sm(tmem_base_slot, kTmemCols);
    else tmem_alloc(tmem_base_slot, kTmemCols);
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_hi0); tma_prefetch_desc(&map_lo0);
    if (p.sliced) { tma_prefetch_desc(&map_hi1); tma_prefetch_desc(&map_lo1); }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp >= 4 && warp < 8) {
    // ===================== A producers ======================================================
    // warp pw owns tile rows [32 pw, 32 pw + 32); instruction j of a k-block loads rows
    // 32 pw + 4 j + (lane >> 3), 16-byte chunk (lane & 7): every LDG covers 4 full 128-byte lines.
    const int pw = warp - 4;
    const int chunk = lane & 7, rsub = lane >> 3;
    // F16: rows of one instruction differ in bit 2 within each half-warp, which keeps the 8-byte stores below conflict-free
    // (the 128B swizzle XORs the 16-byte chunk index with row & 7; a half row of 64 bytes is written per row and instruction)
    auto trow_of = [&](int j) {
      return F16 ? pw * 32 + (j >> 1) * 8 + 4 * (rsub & 1) + (rsub >> 1) + 2 * (j & 1) : pw * 32 + 4 * j + rsub;
    };
    uint32_t it = 0;  // running counter of 32-float k-blocks (ring position = it / kSub)
    const uint32_t full_a_leader = TWO ? mapa_u32(full_a, 0) : 0u;  // cluster address of the leader's full_a[0]
    for (int tile = tile0; tile < n_tiles; tile += tstep) {
      const int mt = TWO ? 2 * ((tile / NT) % MTX) + (int)crank : (tile / NT) % MT, z = tile / (NT * MTX);
      const float* rowp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int lrow = trow_of(j);
        const int m = mt * RPT + lrow;
        rowp[j] = (lrow < RPT && m < p.M) ? p.A + phys_row(p, m, z) * p.lda + chunk * 4 : nullptr;
      }
      // register ring of 3 k-blocks: loads run two k-blocks ahead of the conversion.  (A single load stream
      // across tiles with a 4-deep ring was measured 20-30 % SLOWER: round-1 profile notes.)
      float4 b0[8], b1[8], b2[8];
      auto load_kb = [&](float4(&buf)[8], int kbi) {
        if (kbi < KB) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            buf[j] = rowp[j] ? __ldg((const float4*)(rowp[j] + kbi * kBK)) : make_float4(0, 0, 0, 0);
        }
      };
      auto process = [&](const float4(&buf)[8]) {
        const int s = (it / kSub) % kStages;
        const uint32_t ph = ((it / kSub) / kStages) & 1;
        const int half = F16 ? (int)(it & 1u) : 0;  // F16: which 64-byte half of the stage's 128-byte rows
        if (half == 0) mbar_wait(&empty[s], ph ^ 1, p.err_flag);
        unsigned char* ah = smem + SmemLayout::a_hi(s, BN);
        unsigned char* al = smem + SmemLayout::a_lo(s, BN);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int trow = trow_of(j);
          if constexpr (F16) {
            const float sc = p.a_scale;
            const float x0 = buf[j].x * sc, x1 = buf[j].y * sc, x2 = buf[j].z * sc, x3 = buf[j].w * sc;
            const uint32_t h01 = pack_half2_rn(x0, x1), h23 = pack_half2_rn(x2, x3);
            const uint32_t l01 = pack_half2_rn(x0 - half_bits_to_float(h01 & 0xFFFFu), x1 - half_bits_to_float(h01 >> 16));
            const uint32_t l23 = pack_half2_rn(x2 - half_bits_to_float(h23 & 0xFFFFu), x3 - half_bits_to_float(h23 >> 16));
            const int c16 = half * 4 + (chunk >> 1);
            const int off = (trow >> 3) * 1024 + (trow & 7) * 128 + ((c16 ^ (trow & 7)) << 4) + (chunk & 1) * 8;
            *(uint2*)(ah + off) = make_uint2(h01, h23);
            *(uint2*)(al + off) = make_uint2(l01, l23);
          } else {
            float4 v = buf[j], h, l;
            h.x = tf32_rna(v.x); l.x = tf32_rna(v.x - h.x);
            h.y = tf32_rna(v.y); l.y = tf32_rna(v.y - h.y);
            h.z = tf32_rna(v.z); l.z = tf32_rna(v.z - h.z);
            h.w = tf32_rna(v.w); l.w = tf32_rna(v.w - h.w);
            const int off = (trow >> 3) * 1024 + (trow & 7) * 128 + ((chunk ^ (trow & 7)) << 4);
            *(float4*)(ah + off) = h;
            *(float4*)(al + off) = l;
          }
        }
        fence_proxy_async();  // generic-proxy writes -> visible to the tensor-core (async) proxy
        if (TWO && !leader) mbar_arrive_cluster(full_a_leader + 8u * (uint32_t)s);
        else mbar_arrive(&full_a[s]);
        ++it;
      };
      load_kb(b0, 0);
      load_kb(b1, 1);
      for (int kb = 0; kb < KB; kb += 3) {
        load_kb(b2, kb + 2);
        process(b0);
        if (kb + 1 < KB) { load_kb(b0, kb + 3); process(b1); }
        if (kb + 2 < KB) { load_kb(b1, kb + 4); process(b2); }
      }
    }
  } else if (warp == 8) {
    // ===================== W producer: TMA of the pre-split weight tiles ====================
    if (elect_one()) {
      uint32_t it = 0;
      const uint32_t full_w_leader = TWO ? mapa_u32(full_w, 0) : 0u;
      for (int tile = tile0; tile < n_tiles; tile += tstep) {
        const int nt = tile % NT, z = tile / (NT * MTX);
        const bool second = p.sliced && z >= p.z_split;
        const CUtensorMap* mh = second ? &map_hi1 : &map_hi0;
        const CUtensorMap* ml = second ? &map_lo1 : &map_lo0;
        for (int kb = 0; kb < KB / kSub; ++kb, ++it) {  // one stage = kSub k-blocks = 128 bytes of K per operand row
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1, p.err_flag);
          if constexpr (TWO) {
            // each CTA fetches its half of the weight tile (BN / 2 rows of W^T); both halves complete on the leader's barrier
            if (leader) mbar_expect_tx(&full_w[s], 2u * BN * 128u);
            const int y = nt * BN + (int)crank * (BN / 2);
            tma_load_2d_2sm(mh, full_w_leader + 8u * (uint32_t)s, smem + SmemLayout::w_hi(s, BN), kb * kBK, y);
            tma_load_2d_2sm(ml, full_w_leader + 8u * (uint32_t)s, smem + SmemLayout::w_lo(s, BN), kb * kBK, y);
          } else {
            mbar_expect_tx(&full_w[s], 2u * BN * 128u);
            tma_load_2d(mh, &full_w[s], smem + SmemLayout::w_hi(s, BN), kb * kBK * kSub, nt * BN);
            tma_load_2d(ml, &full_w[s], smem + SmemLayout::w_lo(s, BN), kb * kBK * kSub, nt * BN);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer ========================================================
    const uint32_t idesc = F16 ? make_idesc_f16(kBM, BN) : make_idesc(TWO ? 2 * kBM : kBM, BN);
    uint32_t it = 0, tcount = 0;
    for (int tile = tile0; tile < n_tiles && leader; tile += tstep, ++tcount) {  // pair: only the leader issues
      const int acc = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(&tmem_empty[acc], aph ^ 1, p.err_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      const int KS = KB / kSub;
      for (int kb = 0; kb < KS; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_a[s], ph, p.err_flag);
        mbar_wait(&full_w[s], ph, p.err_flag);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ah = smem_u32(smem + SmemLayout::a_hi(s, BN)), al = smem_u32(smem + SmemLayout::a_lo(s, BN));
          const uint32_t wh = smem_u32(smem + SmemLayout::w_hi(s, BN)), wl = smem_u32(smem + SmemLayout::w_lo(s, BN));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t ko = k * kUmmaK * 4;  // byte offset inside the 128B swizzle row
            if constexpr (F16) {  // 32 bytes = 16 halves per instruction, same descriptor stepping
              umma_f16(d_tmem, make_desc(ah + ko), make_desc(wl + ko), idesc, (kb | k) ? 1u : 0u);
              umma_f16(d_tmem, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
              umma_f16(d_tmem, make_desc(ah + ko), make_desc(wh + ko), idesc, 1u);
            } else if constexpr (TWO) {
              umma_tf32_2sm(d_tmem, make_desc(ah + ko), make_desc(wl + ko), idesc, (kb | k) ? 1u : 0u);
              umma_tf32_2sm(d_tmem, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
              umma_tf32_2sm(d_tmem, make_desc(ah + ko), make_desc(wh + ko), idesc, 1u);
            } else {
              umma_tf32(d_tmem, make_desc(ah + ko), make_desc(wl + ko), idesc, (kb | k) ? 1u : 0u);
              umma_tf32(d_tmem, make_desc(al + ko), make_desc(wh + ko), idesc, 1u);
              umma_tf32(d_tmem, make_desc(ah + ko), make_desc(wh + ko), idesc, 1u);
            }
          }
          if constexpr (TWO) {
            umma_commit_2sm(&empty[s]);                         // frees the stage in BOTH CTAs when the MMAs retire
            if (kb == KS - 1) umma_commit_2sm(&tmem_full[acc]);  // accumulator complete (both epilogues)
          } else {
            umma_commit(&empty[s]);                         // frees the smem stage when the MMAs retire
            if (kb == KS - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: warps 0-3 <-> TMEM lanes 32*warp .. +31 ================
    // tcgen05.ld hands each lane one accumulator ROW (32 columns per chunk).  A 32x36-float
    // shared-memory transpose per warp re-maps that to "8 lanes x float4 per row", so one warp
    // instruction stores (and loads the residual of) 4 complete 128-byte row segments.  Bias and
    // residual loads are issued before the data they are added to is needed.
    constexpr int kPitch = 36;
    float* stage_all = (float*)(smem + SmemLayout::epi(BN));          // [128][kPitch]
    float* stage = stage_all + warp * 32 * kPitch;
    long long* rowinfo_all = (long long*)(smem + SmemLayout::epi(BN) + 4 * 32 * kPitch * 4);
    long long* rowinfo = rowinfo_all + warp * 32;
    float* side1 = (float*)(smem + SmemLayout::side(BN));              // y'  [group][32]
    float* side2 = side1 + 32 * 32;                                    // y'' [group][32]
    float* side3 = side2 + 32 * 32;                                    // sum_t z_t^2 [group][32]
    float* sbias = (float*)(smem + SmemLayout::sbias(BN));             // [BN]
    const int etid = threadIdx.x;                                      // 0..127 (epilogue warps are warps 0-3)
    const float* __restrict__ Resp = p.Res;
    float* __restrict__ Cp = p.C;
    const int rsub = lane >> 3, cq = lane & 7;
    const int S = p.S;
    uint32_t tcount = 0;
    const uint32_t tmem_empty_leader = TWO ? mapa_u32(tmem_empty, 0) : 0u;
    for (int tile = tile0; tile < n_tiles; tile += tstep, ++tcount) {
      const int nt = tile % NT, mt = TWO ? 2 * ((tile / NT) % MTX) + (int)crank : (tile / NT) % MT, z = tile / (NT * MTX);
      const int acc = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      {
        const int lrow = warp * 32 + lane;
        const int m = mt * RPT + lrow;
        long long info = -1;  // invalid row
        if (lrow < RPT && m < p.M) {
          const long long pr = (long long)phys_row(p, m, z);
          // physical row | group index inside the tile | slot  (one division per row and tile)
          info = (pr << 16) | ((long long)(lrow / S) << 8) | (long long)(pr % S);
        }
        __syncwarp();
        rowinfo[lane] = info;
        __syncwarp();
      }
      long long inf[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) inf[jj] = rowinfo[4 * jj + rsub];
      const int rows_here = (p.M - mt * RPT) < RPT ? (p.M - mt * RPT) : RPT;
      const int ngrp = S > 1 ? rows_here / S : rows_here;  // whole slot groups in this tile (S == 1: rows)
      if (p.act) {  // bias of this tile's columns -> shared (latency overlaps the wait for the accumulator)
        named_bar_sync(1, 128);  // every epilogue warp is done with the previous tile's bias
        for (int i = etid; i < BN; i += 128) {
          const int cc = nt * BN + i;
          sbias[i] = (p.bias && cc < p.N) ? __ldg(p.bias + cc) : 0.f;
        }
        named_bar_sync(1, 128);
      }
      const int nchunk = BN / 32;
      const bool vec_ok = (p.N % 4) == 0;
      // residual rows are fetched ONE CHUNK AHEAD (the first chunk before the accumulator wait): their HBM
      // latency overlaps the TMEM read / transpose / store of the previous chunk instead of stalling every chunk
      float4 res[8], resn[8];
      auto load_res = [&](float4(&dst)[8], int cc) {
        const int colc = nt * BN + cc * 32 + 4 * cq;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) dst[jj] = make_float4(0, 0, 0, 0);
        if (Resp && vec_ok && colc < p.N) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (inf[jj] >= 0) dst[jj] = __ldg((const float4*)(Resp + (size_t)(inf[jj] >> 16) * p.ldr + colc));
        }
      };
      load_res(res, 0);
      mbar_wait(&tmem_full[acc], aph, p.err_flag);
      tc_fence_after();
      for (int c = 0; c < nchunk; ++c) {
        const int col = nt * BN + c * 32 + 4 * cq;  // first of this lane's 4 columns
        const bool chunk_on = nt * BN + c * 32 < p.N;  // warp-uniform
        float4 bq = make_float4(0, 0, 0, 0);
        if (c > 0) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) res[jj] = resn[jj];
        }
        if (c + 1 < nchunk) load_res(resn, c + 1);
        if (chunk_on && vec_ok && col < p.N) {
          if (p.bias && !p.act) bq = __ldg((const float4*)(p.bias + col));
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 256 + c * 32), v);
        tmem_ld_wait();
        if (c == nchunk - 1) {  // accumulator fully read: hand it back to the MMA warp early
          tc_fence_before();
          if (TWO && !leader) mbar_arrive_cluster(tmem_empty_leader + 8u * (uint32_t)acc);
          else mbar_arrive(&tmem_empty[acc]);
        }
        if (!chunk_on) continue;
        if constexpr (F16) {  // undo the power-of-two operand scales (exact)
          const float us = p.unscale;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * us);
        }
        const bool act_rows = p.act && S == 1;  // plain forward: every row is a value row -> tanh in registers
        if (act_rows) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(tanh_fwd(__uint_as_float(v[i]) + sbias[c * 32 + i]));
        }
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4)
          *(float4*)(stage + lane * kPitch + 4 * q4) =
              make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]),
                          __uint_as_float(v[4 * q4 + 3]));
        if (p.act && !act_rows) {
          // ---- tanh + forward-Laplacian propagation (reference: hkext.py:104-113 MLP activation; rule
          // y_t = y' z_t, y_L = y' z_L + y'' sum_t z_t^2).  Tiles hold whole slot groups (rpt = G*S).
          named_bar_sync(1, 128);
          // phase A (work spread evenly over the 128 epilogue threads, no divergence):
          // per (group, column): y = tanh(z0 + b) written over the value row, y', y'', sum_t z_t^2
          {
            for (int idx = etid; idx < ngrp * 32; idx += 128) {
              const int g = idx >> 5, cc = idx & 31;
              const int r0 = (S > 1 ? g * S : g);
              float* zp = stage_all + r0 * kPitch + cc;
              const float y = tanhf(*zp + sbias[c * 32 + cc]);
              *zp = y;
              if (S > 1) {
                const float y1 = 1.f - y * y;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                int t = 1;
                for (; t + 3 <= S - 2; t += 4) {
                  const float a0 = zp[t * kPitch], a1 = zp[(t + 1) * kPitch], a2 = zp[(t + 2) * kPitch],
                              a3 = zp[(t + 3) * kPitch];
                  s0 += a0 * a0; s1 += a1 * a1; s2 += a2 * a2; s3 += a3 * a3;
                }
                for (; t <= S - 2; ++t) {
                  const float a = zp[t * kPitch];
                  s0 += a * a;
                }
                const float ss = (s0 + s1) + (s2 + s3);
                side1[g * 32 + cc] = y1;
                side2[g * 32 + cc] = -2.f * y * y1;
                side3[g * 32 + cc] = ss;
              }
            }
          }
          named_bar_sync(1, 128);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (inf[jj] < 0) continue;
            const int slot = (int)(inf[jj] & 255);
            const int lrow = warp * 32 + 4 * jj + rsub;
            float4 o = *(const float4*)(stage_all + lrow * kPitch + 4 * cq);
            if (slot > 0) {
              const int g = (int)((inf[jj] >> 8) & 255);
              const float4 y1 = *(const float4*)(side1 + g * 32 + 4 * cq);
              o.x *= y1.x; o.y *= y1.y; o.z *= y1.z; o.w *= y1.w;
              if (slot == S - 1) {
                const float4 y2 = *(const float4*)(side2 + g * 32 + 4 * cq);
                const float4 ss = *(const float4*)(side3 + g * 32 + 4 * cq);
                o.x += y2.x * ss.x; o.y += y2.y * ss.y; o.z += y2.z * ss.z; o.w += y2.w * ss.w;
              }
            }
            o.x += res[jj].x; o.y += res[jj].y; o.z += res[jj].z; o.w += res[jj].w;
            if (col < p.N) *(float4*)(Cp + (size_t)(inf[jj] >> 16) * p.ldc + col) = o;
          }
          named_bar_sync(1, 128);  // stage / side buffers reused by the next chunk
          continue;
        }
        __syncwarp();
        if (vec_ok) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float4 o = *(const float4*)(stage + (4 * jj + rsub) * kPitch + 4 * cq);
            o.x += res[jj].x; o.y += res[jj].y; o.z += res[jj].z; o.w += res[jj].w;
            if ((inf[jj] & 255) == 0) { o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w; }
            if (inf[jj] >= 0 && col < p.N) *(float4*)(Cp + (size_t)(inf[jj] >> 16) * p.ldc + col) = o;
          }
        } else {  // ragged N: scalar tail path
#pragma unroll 1
          for (int jj = 0; jj < 8; ++jj) {
            const long long inj = rowinfo[4 * jj + rsub];
            if (inj < 0) continue;
            const size_t pr = (size_t)(inj >> 16);
            for (int e = 0; e < 4; ++e) {
              if (col + e >= p.N) break;
              float o = stage[(4 * jj + rsub) * kPitch + 4 * cq + e];
              if (p.bias && !p.act && (inj & 255) == 0) o += p.bias[col + e];
              if (Resp) o += Resp[pr * p.ldr + col + e];
              Cp[pr * p.ldc + col + e] = o;
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer can still reach into it
  if (warp == 9) {
    tc_fence_after();
    if constexpr (TWO) tmem_dealloc_2sm(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---- host side -----------------------------------------------------------------------------
// W^T split tensors: [Nrows][K] fp32, K contiguous.  Box = 32 fp32 (128 B) x BN rows, 128B swizzle.
inline int make_weight_map(CUtensorMap* map, const float* wt, int Nrows, int K, int BN) {
  return make_kmajor_map(map, wt, 4, Nrows, K, kBK, BN);
}

inline int pick_bn(int N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

}  // namespace tc
}  // namespace dq
