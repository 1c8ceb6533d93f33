// Microbenchmark (development aid): issue rate of tcgen05.mma kind::f16, M = 128, one CTA per SM, operands = whatever the
// shared / tensor memory holds.  Variants: SS / TS (A from tensor memory), N = 64 / 128 / 256, one accumulator vs two
// alternating accumulators.  Prints clocks per instruction (chip median over the CTAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepqmc_b200/csrc -o /tmp/umma_rate tools/microbench/umma_rate.cu
#include <cstdio>
#include <vector>
#include <algorithm>
#include "tc_ptx.cuh"
using namespace dq::tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int ts, int N, int nacc, int n_mma, int commit_every, int n_commit, long long* out) {
  __shared__ uint64_t cbar[8];
  DQMC_TC_SMEM(smem);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&cbar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    long long t0 = 0, t1 = 0;
    const uint32_t idesc = make_idesc_f16(128, N);
    if (elect_one()) {
      const uint64_t ad = make_desc(smem_u32(smem)), bd = make_desc(smem_u32(smem) + 65536u);
      t0 = clock64();
      const int group = commit_every ? commit_every : n_mma;  // no division inside the issue loop
      int i = 0, g = 0;
      while (i < n_mma) {
#pragma unroll 4
        for (int e = 0; e < group; ++e, ++i) {
          const uint32_t d = tb + 256u + (nacc == 2 ? (uint32_t)(i & 1) * 128u : 0u);
          const uint32_t k = (uint32_t)(i & 3);
          const uint64_t b = bd + 2u * k + (uint64_t)(((i >> 2) & 3) * 1024u);  // 16 KB slots
          if (ts) umma_f16_ts(d, tb + 8u * k + (uint32_t)((i >> 2) & 7) * 32u, b, idesc, 1u);
          else umma_f16(d, ad + 2u * k + (uint64_t)(((i >> 2) & 3) * 1024u), b, idesc, 1u);
        }
        if (commit_every)
          for (int c = 0; c < n_commit; ++c) umma_commit(&cbar[(g + c) & 7]);
        ++g;
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, nullptr);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 480;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128, 256})
      for (int nacc = 1; nacc <= 2; ++nacc) {
        if (N == 256 && nacc == 2) continue;
        for (int grid : {1, 148}) {
          rate_kernel<<<grid, 128, smem>>>(ts, N, nacc, n, 0, 0, d_out);
          rate_kernel<<<grid, 128, smem>>>(ts, N, nacc, n, 0, 0, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<long long> h(grid);
          cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          std::sort(h.begin(), h.end());
          printf("%s N=%3d accumulators=%d grid=%3d: %7.1f clk / MMA (median CTA; min %.1f max %.1f)\n", ts ? "TS" : "SS", N, nacc, grid,
                 (double)h[grid / 2] / n, (double)h[0] / n, (double)h[grid - 1] / n);
        }
      }
  // cost of tcgen05.commit between the instructions (TS, N = 128, one accumulator, whole chip)
  for (int every : {0, 24, 12, 6, 4})
    for (int nc : {1, 2, 3}) {
      if (!every && nc > 1) continue;
      rate_kernel<<<148, 128, smem>>>(1, 128, 1, n, every, nc, d_out);
      rate_kernel<<<148, 128, smem>>>(1, 128, 1, n, every, nc, d_out);
      cudaDeviceSynchronize();
      std::vector<long long> h(148);
      cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("TS N=128 commit x%d every %2d MMAs: %7.1f clk / MMA\n", nc, every, (double)h[74] / n);
    }
  return 0;
}
