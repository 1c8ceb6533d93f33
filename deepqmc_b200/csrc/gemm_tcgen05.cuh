// fp32-accurate dense-layer GEMM on the 5th-gen tensor cores (sm_100a): 3xTF32 split
//   C[row(m), :] = (Res) + A[row(m), :] @ W + (bias on value rows),  A, W, C fp32 in HBM.
//
//   a = a_hi + a_lo, w = w_hi + w_lo with *_hi = top 19 bits (exactly representable in TF32);
//   acc(fp32, TMEM) = a_hi w_lo + a_lo w_hi + a_hi w_hi        -> ~2^-21 relative per product,
//   i.e. the accuracy class the reference demands (jax_default_matmul_precision='highest',
//   NVIDIA_TF32_OVERRIDE=0: src/deepqmc/__init__.py:9-34) at 1/3 of the TF32 tensor peak.
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
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

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
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must not hang the GPU box -> flag + trap after ~2 s.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) atomicExch(err_flag, 1);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// ---- CTA-pair (cta_group::2) helpers: instruction forms as in CUTLASS' cute/arch/copy_sm100_tma.hpp
// (SM100_TMA_2SM_LOAD_2D), cutlass/arch/barrier.h (ClusterBarrier::arrive, umma_arrive_multicast_2x1SM) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared-memory object of THIS CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of one CTA's half of a CTA-pair operand; completes on the mbarrier at cluster address `bar_cluster`
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// commit the MMAs issued so far; the arrival lands on the barrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 1024B/16
// | version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c_format=F32 (1) [4,6),
// a_format=b_format=TF32 (2) [7,10),[10,13), a/b K-major (0), n>>3 [17,23), m>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tanh for the plain-forward epilogue: odd polynomial below |x| = 0.15, 1 - 2 / (1 + e^{2x})
// (ex2.approx + fast division) above; absolute error <= ~3e-7.  The forward-Laplacian epilogue
// keeps tanhf (derivative slots amplify the error); plain forwards only feed log|psi| ratios.
__device__ __forceinline__ float tanh_fwd(float x) {
  const float x2 = x * x;
  const float poly = x + x * x2 * (-0.33333333333f + x2 * (0.13333333333f + x2 * (-0.05396825397f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  const float big = 1.f - __fdividef(2.f, 1.f + e);
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

template <bool TWO>
__global__ void __launch_bounds__(kThreads, 1)
gemm3xtf32_kernel(const __grid_constant__ CUtensorMap map_hi0, const __grid_constant__ CUtensorMap map_lo0,
                  const __grid_constant__ CUtensorMap map_hi1, const __grid_constant__ CUtensorMap map_lo1, Params p) {
  // 1024-byte aligned dynamic shared memory (SWIZZLE_128B atoms).  No integer round-up of the
  // pointer: that would drop the shared address space and turn every access into a generic LD/ST.
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
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
  const int KB = p.K / kBK;
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
      mbar_init(&full_a[s], kPer);
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
    if constexpr (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
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
    uint32_t it = 0;  // running k-block counter (ring position)
    const uint32_t full_a_leader = TWO ? mapa_u32(full_a, 0) : 0u;  // cluster address of the leader's full_a[0]
    for (int tile = tile0; tile < n_tiles; tile += tstep) {
      const int mt = TWO ? 2 * ((tile / NT) % MTX) + (int)crank : (tile / NT) % MT, z = tile / (NT * MTX);
      const float* rowp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int lrow = pw * 32 + 4 * j + rsub;
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
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1, p.err_flag);
        unsigned char* ah = smem + SmemLayout::a_hi(s, BN);
        unsigned char* al = smem + SmemLayout::a_lo(s, BN);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int trow = pw * 32 + 4 * j + rsub;
          float4 v = buf[j], h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
          const int off = (trow >> 3) * 1024 + (trow & 7) * 128 + ((chunk ^ (trow & 7)) << 4);
          *(float4*)(ah + off) = h;
          *(float4*)(al + off) = l;
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
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t full_w_leader = TWO ? mapa_u32(full_w, 0) : 0u;
      for (int tile = tile0; tile < n_tiles; tile += tstep) {
        const int nt = tile % NT, z = tile / (NT * MTX);
        const bool second = p.sliced && z >= p.z_split;
        const CUtensorMap* mh = second ? &map_hi1 : &map_hi0;
        const CUtensorMap* ml = second ? &map_lo1 : &map_lo0;
        for (int kb = 0; kb < KB; ++kb, ++it) {
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
            tma_load_2d(mh, &full_w[s], smem + SmemLayout::w_hi(s, BN), kb * kBK, nt * BN);
            tma_load_2d(ml, &full_w[s], smem + SmemLayout::w_lo(s, BN), kb * kBK, nt * BN);
          }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer ========================================================
    const uint32_t idesc = make_idesc(TWO ? 2 * kBM : kBM, BN);
    uint32_t it = 0, tcount = 0;
    for (int tile = tile0; tile < n_tiles && leader; tile += tstep, ++tcount) {  // pair: only the leader issues
      const int acc = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(&tmem_empty[acc], aph ^ 1, p.err_flag);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_a[s], ph, p.err_flag);
        mbar_wait(&full_w[s], ph, p.err_flag);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t ah = smem_u32(smem + SmemLayout::a_hi(s, BN)), al = smem_u32(smem + SmemLayout::a_lo(s, BN));
          const uint32_t wh = smem_u32(smem + SmemLayout::w_hi(s, BN)), wl = smem_u32(smem + SmemLayout::w_lo(s, BN));
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint32_t ko = k * kUmmaK * 4;  // byte offset inside the 128B swizzle row
            if constexpr (TWO) {
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
            if (kb == KB - 1) umma_commit_2sm(&tmem_full[acc]);  // accumulator complete (both epilogues)
          } else {
            umma_commit(&empty[s]);                         // frees the smem stage when the MMAs retire
            if (kb == KB - 1) umma_commit(&tmem_full[acc]);  // accumulator complete
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
        asm volatile("bar.sync 1, 128;" ::: "memory");  // every epilogue warp is done with the previous tile's bias
        for (int i = etid; i < BN; i += 128) {
          const int cc = nt * BN + i;
          sbias[i] = (p.bias && cc < p.N) ? __ldg(p.bias + cc) : 0.f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
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
          asm volatile("bar.sync 1, 128;" ::: "memory");
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
          asm volatile("bar.sync 1, 128;" ::: "memory");
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
          asm volatile("bar.sync 1, 128;" ::: "memory");  // stage / side buffers reused by the next chunk
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
    if constexpr (TWO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---- host side -----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && ptr)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// W^T split tensors: [Nrows][K] fp32, K contiguous.  Box = 32 fp32 (128 B) x BN rows, 128B swizzle.
inline int make_weight_map(CUtensorMap* map, const float* wt, int Nrows, int K, int BN) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)Nrows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)wt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

inline int pick_bn(int N) { return N > 128 ? 256 : (N > 64 ? 128 : 64); }

}  // namespace tc
}  // namespace dq
