// Minimal CPU emulation of the CUDA execution model, DEVELOPMENT/TEST TOOL ONLY.
//
// The build container has no GPU; this header lets the *same kernel sources* under
// deepqmc_b200/csrc be compiled with g++ (-DDQMC_EMU) and executed block by block on the
// CPU so that indexing / reduction / barrier logic can be checked against the oracle before
// spending GPU minutes.  Threads of a block are ucontext fibers scheduled round-robin;
// __syncthreads / warp shuffles are cooperative yields.  It is never linked into the shipped
// library (libdqmc_b200.so is built by nvcc only) and the product path never falls back to it.
#pragma once
#include <ucontext.h>

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };

namespace emu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  uint3_emu tidx{0, 0, 0};
  int lin = 0;
  bool done = false;
};
struct State {
  std::vector<Fiber> fibers;
  ucontext_t main_ctx;
  int cur = 0;
  int nthreads = 0;
  uint3_emu bidx{0, 0, 0};
  dim3 bdim, gdim;
  // block barrier
  long bar_gen = 0;
  int bar_count = 0;
  // warp exchange
  std::vector<uint64_t> shfl_buf;      // [nthreads]
  std::vector<int> warp_count;         // arrivals per warp
  std::vector<long> warp_gen;          // generation per warp
  std::vector<unsigned char> dyn_smem;
  std::function<void()> body;
};
inline State& S() { static thread_local State s; return s; }

inline void yield() {
  State& s = S();
  int me = s.cur;
  // next runnable fiber (round robin); if none other, continue
  for (int k = 1; k <= s.nthreads; ++k) {
    int nxt = (me + k) % s.nthreads;
    if (!s.fibers[nxt].done) {
      if (nxt == me) return;
      s.cur = nxt;
      swapcontext(&s.fibers[me].ctx, &s.fibers[nxt].ctx);
      return;
    }
  }
}
inline void fiber_entry() {
  State& s = S();
  s.body();
  int me = s.cur;
  s.fibers[me].done = true;
  // switch to next unfinished fiber or back to main
  for (int k = 1; k <= s.nthreads; ++k) {
    int nxt = (me + k) % s.nthreads;
    if (!s.fibers[nxt].done) {
      s.cur = nxt;
      setcontext(&s.fibers[nxt].ctx);
    }
  }
  setcontext(&s.main_ctx);
}
inline void syncthreads() {
  State& s = S();
  long gen = s.bar_gen;
  if (++s.bar_count == s.nthreads) {
    s.bar_count = 0;
    s.bar_gen++;
    return;
  }
  while (s.bar_gen == gen) yield();
}
inline void syncwarp() {
  State& s = S();
  int w = s.fibers[s.cur].lin / 32;
  int wsize = std::min(32, s.nthreads - w * 32);
  long gen = s.warp_gen[w];
  if (++s.warp_count[w] == wsize) {
    s.warp_count[w] = 0;
    s.warp_gen[w]++;
    return;
  }
  while (s.warp_gen[w] == gen) yield();
}
template <class T>
inline T shfl_idx(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shfl of <=8 byte types");
  State& s = S();
  int lin = s.fibers[s.cur].lin;
  int w = lin / 32;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  s.shfl_buf[lin] = raw;
  syncwarp();
  int src = w * 32 + (src_lane & 31);
  if (src >= s.nthreads) src = lin;
  uint64_t got = s.shfl_buf[src];
  syncwarp();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
  State& s = S();
  assert(block.y == 1 && block.z == 1 && "emulator supports 1-D blocks");
  s.nthreads = block.x;
  s.bdim = block;
  s.gdim = grid;
  s.body = body;
  s.dyn_smem.assign(smem + 64, 0);
  if ((int)s.fibers.size() < s.nthreads) s.fibers.resize(s.nthreads);
  s.shfl_buf.assign(s.nthreads, 0);
  int nw = (s.nthreads + 31) / 32;
  const size_t kStack = 256 * 1024;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bidx = {bx, by, bz};
        s.bar_gen = 0;
        s.bar_count = 0;
        s.warp_count.assign(nw, 0);
        s.warp_gen.assign(nw, 0);
        for (int t = 0; t < s.nthreads; ++t) {
          Fiber& f = s.fibers[t];
          if (f.stack.size() != kStack) f.stack.resize(kStack);
          f.tidx = {(unsigned)t, 0, 0};
          f.lin = t;
          f.done = false;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack.data();
          f.ctx.uc_stack.ss_size = kStack;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        s.cur = 0;
        swapcontext(&s.main_ctx, &s.fibers[0].ctx);
      }
}
}  // namespace emu

struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <class T> inline T __ldg(const T* p) { return *p; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define threadIdx (emu::S().fibers[emu::S().cur].tidx)
#define blockIdx (emu::S().bidx)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)
inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu::shfl_idx(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) {
  return emu::shfl_idx(v, (emu::S().fibers[emu::S().cur].lin & 31) ^ m);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, int d) {
  int lane = emu::S().fibers[emu::S().cur].lin & 31;
  return emu::shfl_idx(v, lane + d < 32 ? lane + d : lane);
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }

// ---- tiny runtime shim ---------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
#define DQMC_DYN_SMEM(name) unsigned char* name = (unsigned char*)(((uintptr_t)emu::S().dyn_smem.data() + 15) & ~(uintptr_t)15)
#define DQMC_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), smem, [=]() { kern(__VA_ARGS__); })
