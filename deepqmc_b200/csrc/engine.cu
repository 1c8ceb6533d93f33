// Host side of libdqmc_b200.so: engine object, parameter table, workspace planning, kernel
// sequencing, C ABI (include/dqmc_b200.h).  Built by nvcc for sm_100a; with -DDQMC_EMU the same
// file builds against tools/cuda_emu for CPU-side logic checks during development (never shipped).
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "dqmc_b200.h"
#include "kernels_mcmc.cuh"
#include "kernels_slater.cuh"
#include "kernels_trunk.cuh"
#include "kernels_gnn.cuh"
#include "kernels_bwd.cuh"
#include "attn_mma.cuh"
#if !defined(DQMC_NO_TCGEN05)
#include "gemm_tcgen05.cuh"
#include "fused_tc.cuh"
#include "trunk_tc.cuh"
#endif

namespace dq {

struct ParamEntry {
  std::string name;
  int64_t offset;
  int rows, cols;
};

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct EngineBase {
  dqmc_config cfg;
  int device = 0;
  std::string err;
  int64_t launches = 0;
  std::vector<ParamEntry> entries;
  int64_t total = 0;
  // Planning pass: the host code of an entry point is walked with every CUDA call skipped, so that the number of workspace
  // bytes a call carves is computed by the SAME code that carves them (dqmc_workspace_bytes, dqmc_debug_plan).  An engine
  // created with device < 0 is plan-only (dry for its whole life: no CUDA context, no allocation).
  bool dry = false;
  bool plan_only = false;
  mutable const char* dry_hwm = nullptr;  // highest workspace address carved during a dry pass
  void note_hwm(const void* q) const { if (dry && (const char*)q > dry_hwm) dry_hwm = (const char*)q; }
  // optional per-launch timing of the dominant (GEMM) kernels with CUDA events on the caller's stream
  bool prof = false;
  double prof_flops = 0;
  int64_t prof_n = 0;
  // per kernel class: 0 row GEMM (one dense layer per launch), 1 fused MLP block, 2 whole trunk (all layers in one launch)
  double prof_cls_flops[3] = {0, 0, 0};
  int64_t prof_cls_n[3] = {0, 0, 0};
  std::vector<int> prof_cls;
#ifndef DQMC_EMU
  std::vector<cudaEvent_t> prof_ev;
  void prof_note(cudaEvent_t e0, cudaEvent_t e1, double flops, int cls) {
    prof_ev.push_back(e0); prof_ev.push_back(e1);
    prof_cls.push_back(cls);
    prof_flops += flops; prof_cls_flops[cls] += flops;
    ++prof_n; ++prof_cls_n[cls];
  }
#endif
  virtual ~EngineBase() {}
  virtual int set_params(const double* host, int64_t n, cudaStream_t st) = 0;
  virtual int64_t ws_bytes(int B, int mode) = 0;
  virtual int64_t ws_bytes_min(int B, int mode) = 0;
  virtual int debug_plan(int B, int mode, int64_t wsb, int64_t* planned, int64_t* carved) = 0;
  virtual int stats_pack(const void* E, const void* stats, int B, double* out, cudaStream_t st) = 0;
  virtual int debug_mlp_block(int layer, const void* O, const void* X, void* Out, int rows, cudaStream_t st) = 0;
  virtual int debug_trunk(const void* X0, void* Out, int rows, cudaStream_t st) = 0;
  virtual int forward(const void* r, const void* R, int Rb, int B, void* sign, void* logp, void* ws, int64_t wsb,
                      cudaStream_t st) = 0;
  virtual int local_energy(const void* r, const void* R, int Rb, int B, uint64_t seed, const void* twist, void* E,
                           void* stats, void* sign, void* logp, void* grad, void* ws, int64_t wsb,
                           cudaStream_t st) = 0;
  virtual int langevin(void* r, void* sign, void* logp, void* force, int32_t* age, void* tau, const void* R, int Rb, int B,
                       int n_sub, double target, int max_age, uint64_t seed, uint64_t step0, uint64_t woff, const void* nn,
                       const void* nu, void* stats, void* ws, int64_t wsb, cudaStream_t st) = 0;
  virtual int vjp_params(const void* r, const void* R, int Rb, int B, const void* weights, void* sign, void* logp,
                         void* grad_params, void* ws, int64_t wsb, cudaStream_t st) = 0;
  virtual int orbitals(const void* r, const void* R, int Rb, int B, void* out, void* ws, int64_t wsb, cudaStream_t st) = 0;
  virtual int set_ph(int n_tab, int n_grid, double r_max, const double* tables, const int32_t* tab_of_nuc) = 0;
  virtual int debug_gemm(const char* wname, const char* bname, const void* A, const void* Res, void* C, int Mr, int S,
                         int sliced, int backend, cudaStream_t st) = 0;
  virtual int mcmc(void* r, void* sign, void* logp, int32_t* age, void* tau, const void* R, int Rb, int B, int n_sub,
                   double target, int max_age, uint64_t seed, uint64_t step0, uint64_t woff, const void* nn,
                   const void* nu, void* stats, void* ws, int64_t wsb, cudaStream_t st, double p_exchange = 0.0,
                   const int32_t* ex_flags = nullptr, const int32_t* ex_idx = nullptr) = 0;

  void add(const std::string& n, int rows, int cols) {
    entries.push_back({n, total, rows, cols});
    total += (int64_t)rows * cols;
  }
  void build_layout() {
    const int N = cfg.n_up + cfg.n_down, M = cfg.n_nuc, d = cfg.embedding_dim, K = cfg.n_determinants;
    const int rep = cfg.n_env_per_nuc > 1 ? cfg.n_env_per_nuc : 1;
    if (cfg.kind == DQMC_PSIFORMER || cfg.kind == DQMC_TRANSPSIFORMER) {
      add("emb.w", 4 * M + 1, d);
      for (int l = 0; l < cfg.n_layers; ++l) {
        std::string p = "L" + std::to_string(l) + ".";
        add(p + "wqkv", d, 3 * d);
        add(p + "wo", d, d);
        add(p + "w1", d, d);
        add(p + "b1", 1, d);
        add(p + "w2", d, d);
        add(p + "b2", 1, d);
        if (cfg.kind == DQMC_TRANSPSIFORMER && cfg.n_nuc_tokens > 0) {
          add(p + "kn", cfg.n_nuc_tokens, d);  // key / value rows of the nuclear tokens (host-evaluated stream)
          add(p + "vn", cfg.n_nuc_tokens, d);
        }
      }
    }
    int bf_in = d;  // input width of the final backflow layer
    if (cfg.kind == DQMC_PAULINET) {
      // reference tests/conf/ansatz.yaml and conf/ansatz/default.yaml; entry names mirror the Haiku modules
      const int e = cfg.edge_dim, nl = cfg.gnn_sub_n > 0 ? cfg.gnn_sub_n : 1;
      if (!cfg.gnn_features) add("emb.table", cfg.n_elec_types > 0 ? cfg.n_elec_types : 1, d);
      int dcur = cfg.gnn_features ? 4 * M : d, ecur = 4;
      const int nt = cfg.gnn_conv_ne ? 3 : 2;
      const char* tn[3] = {"same", "anti", "ne"};
      for (int l = 0; l < cfg.n_layers; ++l) {
        std::string p = "G" + std::to_string(l) + ".";
        for (int t = 0; t < nt; ++t) {
          int din = ecur;
          for (int i = 0; i < nl; ++i) {
            const std::string q = p + "w_" + tn[t] + "." + std::to_string(i);
            add(q + ".w", din, cfg.gnn_w_dims[l][i]);
            if (cfg.gnn_w_bias) add(q + ".b", 1, cfg.gnn_w_dims[l][i]);
            din = cfg.gnn_w_dims[l][i];
          }
          if (t < 2) {
            din = dcur;
            for (int i = 0; i < nl; ++i) {
              const std::string q = p + "h_" + tn[t] + "." + std::to_string(i);
              add(q + ".w", din, cfg.gnn_h_dims[l][i]); add(q + ".b", 1, cfg.gnn_h_dims[l][i]);
              din = cfg.gnn_h_dims[l][i];
            }
          } else {
            add(p + "hne", M, e);  // h_ne(nuclear embedding table): walker-independent, evaluated on the host
          }
          if (!cfg.gnn_concat) { add(p + "g_" + tn[t] + ".w", e, d); add(p + "g_" + tn[t] + ".b", 1, d); }
        }
        if (cfg.gnn_concat) {
          add(p + "g.w", 3 * dcur + nt * e, d);
          if (cfg.gnn_g_bias) add(p + "g.b", 1, d);
        }
        if (cfg.gnn_deep_edges && l < cfg.n_layers - 1) {
          int din = ecur;
          for (int i = 0; i < nl; ++i) {
            const std::string q = p + "u." + std::to_string(i);
            add(q + ".w", din, cfg.gnn_u_dims[l][i]); add(q + ".b", 1, cfg.gnn_u_dims[l][i]);
            din = cfg.gnn_u_dims[l][i];
          }
          ecur = e;
        }
        dcur = d;
      }
      int din = d;
      for (int i = 0; i < cfg.jastrow_n; ++i) {
        add("J" + std::to_string(i) + ".w", din, cfg.jastrow_dims[i]);
        if (i < cfg.jastrow_n - 1) add("J" + std::to_string(i) + ".b", 1, cfg.jastrow_dims[i]);
        din = cfg.jastrow_dims[i];
      }
      din = d;
      for (int i = 0; i < cfg.backflow_n; ++i) {
        const std::string q = std::to_string(i);
        add("bfh" + q + ".up", din, cfg.backflow_dims[i]); add("bfh" + q + ".dn", din, cfg.backflow_dims[i]);
        add("bfb" + q + ".up", 1, cfg.backflow_dims[i]); add("bfb" + q + ".dn", 1, cfg.backflow_dims[i]);
        din = cfg.backflow_dims[i];
      }
      bf_in = din;
      add("bfb.up", 1, K * N); add("bfb.dn", 1, K * N);
      if (cfg.conf_linear) add("conf.w", 1, K);
    }
    if (cfg.kind == DQMC_FERMINET) {
      const int de = cfg.edge_dim;
      int din = 4 * M, ein = 4;
      for (int l = 0; l < cfg.n_layers; ++l) {
        std::string p = "F" + std::to_string(l) + ".";
        add(p + "wg", 3 * din + 2 * ein, d);
        add(p + "bg", 1, d);
        if (l < cfg.n_layers - 1) {
          add(p + "wu", ein, de);
          add(p + "bu", 1, de);
        }
        din = d; ein = de;
      }
    }
    add("bf.up", bf_in, K * N * (cfg.backflow_add == 2 ? 2 : 1));
    add("bf.dn", bf_in, K * N * (cfg.backflow_add == 2 ? 2 : 1));
    add("env.pi_up", K * N, M * rep);
    add("env.pi_dn", K * N, M * rep);
    add("env.zeta_up", K * N, M * rep);
    add("env.zeta_dn", K * N, M * rep);
    add("cusp.alpha", 1, 2);
    if (cfg.nuc_cusp_kind) add("cusp.nuc", 1, 1 + M);  // alpha_nuc, nuclear charges
  }
  int64_t off(const std::string& n) const {
    for (auto& e : entries)
      if (e.name == n) return e.offset;
    return -1;
  }
};

template <class T>
__global__ void convert_kernel(const double* __restrict__ src, T* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (T)src[i];
}

// W[K][N] (row-major) -> W^T hi/lo [N][K]: hi = rna_tf32(w), lo = rna_tf32(w - hi)
__global__ void split_transpose_kernel(const float* __restrict__ W, int K, int N, float* __restrict__ hi,
                                       float* __restrict__ lo) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  int n = idx / K, k = idx % K;
  float w = W[(size_t)k * N + n];
#if !defined(DQMC_NO_TCGEN05)
  float h = tc::tf32_rna(w);
  hi[idx] = h;
  lo[idx] = tc::tf32_rna(w - h);
#else
  float h = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
  hi[idx] = h;
  lo[idx] = w - h;
#endif
}

#if !defined(DQMC_NO_TCGEN05)
// W[K][N] (row-major) -> (W 2^e)^T as IEEE halves, hi / lo planes [N][K]: hi = rn(w'), lo = rn(w' - hi)
__global__ void split_transpose_f16_kernel(const float* __restrict__ W, int K, int N, float scale, uint16_t* __restrict__ hi,
                                           uint16_t* __restrict__ lo) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  int n = idx / K, k = idx % K;
  const float w = W[(size_t)k * N + n] * scale;
  const uint32_t h = tc::pack_half2_rn(w, 0.f) & 0xFFFFu;
  const uint32_t l = tc::pack_half2_rn(w - tc::half_bits_to_float(h), 0.f) & 0xFFFFu;
  hi[idx] = (uint16_t)h;
  lo[idx] = (uint16_t)l;
}
#endif

#define DQ_CHECK(call)                                                             \
  do {                                                                             \
    if (dry) break; /* planning pass: no CUDA calls */                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the kernel, not to an engine, and engines for molecules of different
// size share one process: the opt-in therefore only ever grows (a later, smaller engine must not lower the cap under an
// earlier, larger one).
template <class F>
inline cudaError_t raise_dyn_smem(F fn, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, int> high;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  int& h = high[std::make_pair(dev, (const void*)fn)];
  if (bytes > h) h = bytes;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, h);
}

#define DQ_LAUNCH(kern, grid, block, smem, stream, ...)          \
  do {                                                           \
    if (dry) break; /* planning pass */                          \
    auto kfn_ = kern;                                            \
    DQMC_LAUNCH(kfn_, grid, block, smem, stream, __VA_ARGS__);   \
    ++launches;                                                  \
  } while (0)

// Emulator builds (tools/cuda_emu) put a 256-byte guard behind every buffer carved from the workspace and verify the
// guards after each chunk: a buffer that is individually too small (with a consistent total) cannot hide.
// The guard bytes are part of what carve() / the take lambdas advance by, so the PLAN (a dry pass of the same code) contains
// them too: emulator and hardware builds plan by the same rule, there is no build-dependent slack.
#ifdef DQMC_EMU
#define DQ_TAKE_GUARD() do { if (!dry) { std::memset(p, 0xC3, 256); emu_guards.push_back((unsigned char*)p); } p += 256; } while (0)
#else
#define DQ_TAKE_GUARD() do {} while (0)
#endif

template <class T>
struct Engine : EngineBase {
  T* d_params = nullptr;
  T* d_params_t = nullptr;  // every entry transposed ([cols][rows]): operands of the dA = dY W^T products of the reverse pass
  double* d_stage = nullptr;
  T* d_zval = nullptr;
  T* d_znuc = nullptr;  // full nuclear charges (Langevin clean_force)
  int* d_ecp_mask = nullptr;
  T* d_ecp_loc = nullptr;
  T* d_nl_params = nullptr;
  int* d_nl_nuc = nullptr;
  int J = 0;  // nuclei with a non-local channel
  // pseudo-Hamiltonian (reference ecp/pseudo_hamiltonian.py): tables r V_loc / r V_L2 per element on a uniform grid
  T* d_ph_tabs = nullptr;
  int* d_ph_nuc = nullptr;
  int ph_G = 0;
  double ph_rmax = 0;
  mutable std::vector<unsigned char*> emu_guards;  // emulator builds only: guard zones behind the carved workspace buffers
  int check_guards() {
#ifdef DQMC_EMU
    for (unsigned char* g : emu_guards)
      for (int i = 0; i < 256; ++i)
        if (g[i] != 0xC3) { emu_guards.clear(); err = "emulator: a workspace buffer was overrun (guard zone modified)"; return 4; }
    emu_guards.clear();
#endif
    return 0;
  }
  int64_t vjp_ws_cap = 0;  // bytes of the caller's workspace during a reverse pass (extent check of the chunk buffers)
  bool ph_on = false;      // tables uploaded
  bool ph_active = false;  // set around the forward-Laplacian pass of local_energy only
  T* mos_out = nullptr;    // set by orbitals(): the tail writes the orbital matrices of the chunk here and stops
  int attn_tb = 1, attn_tb1 = 1;
  bool attn_gen_mma = false;  // ... same for the generic kernel (TransPsiformer: extra key / value tokens)
  bool attn_fl_mma = false;   // fp32 forward-Laplacian attention: tangent chunks as warp-level 3xTF32 mma.sync products
  int attn_fl_threads = 128;  // block size of the fp32 forward-Laplacian attention (large molecules: one block per SM fits -> more warps)
  bool attn_f32 = false;
  bool embed_fwd_ok = false;
  bool attn_fwd_ok = false;
  bool attn_fwd_pipelined = false;
  bool attn_mma_ok = false;  // plain-forward attention on mma.sync (attn_mma.cuh)
  bool slater_fwd2_ok = false;
  int N, M, d, K, KN, H, dh, T3;
  int BFW = 0;      // row width of the backflow buffer: K N, or 2 K N with multiplicative + additive heads
  int add_off = 0;  // column offset of the additive head in it
  bool gnn = false;    // conv-GNN ("PauliNet" test ansatz)
  int bf_in = 0;       // input width of the final backflow layer
  bool trans = false;  // TransPsiformer: nuclear attention tokens + nucleus-dependent envelopes
  int Mn = 0, env_rep = 1;
  size_t max_smem = 0;
  int n_sms = 148;
#if !defined(DQMC_NO_TCGEN05)
  struct TcWeight {
    float* hi = nullptr; float* lo = nullptr; CUtensorMap mh, ml, mh2, ml2; int N = 0, K = 0, BN = 0;
    // "3xFP16" operands of the plain-forward kernels: (W 2^e)^T as halves, hi / lo planes [N][K]; maps with BN-row boxes
    // (row GEMM) and with all-N-row boxes (fused MLP block, N <= 256)
    uint16_t* h16 = nullptr; CUtensorMap m16h, m16l, m16h_all, m16l_all; float wscale = 1.f; bool f16 = false, f16_all = false;
    CUtensorMap m16h_128, m16l_128; bool f16_128 = false;  // 128-row boxes: weight slots of the whole-trunk kernel (trunk_tc.cuh)
  };
  // non-local ECP group in flight: envelope table of its base walkers [nb][N][K N] (null outside the quadrature forwards),
  // index of the current chunk's first virtual walker, virtual walkers per base walker (J N 12)
  const T* ecp_env = nullptr;
  const T* ecp_emb = nullptr;  // ... and their embedding rows [nb][N][d] (whole-trunk kernel only)
  int64_t ecp_v0 = 0;
  int ecp_vper = 0;
  bool fuse_trunk = true;  // all layers of a plain forward in one persistent launch (trunk_tc.cuh); DQMC_TC_TRUNK=0 disables
  bool trunk_ts = true;    // ... with the A operand of its dense GEMMs in tensor memory (DQMC_TC_TRUNK_TS=0: shared memory)
  CUtensorMap* d_trunk_maps = nullptr;       // [L][4][2]
  unsigned char* d_trunk_scratch = nullptr;  // n_sms x 384 KB Q / K / V planes
  long long* d_trunk_trace = nullptr;        // DQMC_TRUNK_TRACE: 64 clock stamps of one tile (development aid)
  bool gemm_2cta = false;  // CTA-pair (cta_group::2) variant of the dense-layer GEMM
  bool f16_on = true;      // plain forwards (S = 1) on kind::f16 with hi / lo half operands; DQMC_TC_F16=0: stay on 3xTF32
  bool fuse_mlp = true;    // W_o + residual -> W1 + tanh -> W2 + tanh + residual in one launch (S = 1); DQMC_TC_FUSE_MLP=0 disables
  static constexpr float kActScale = 16.f;  // 2^4: |activation| < 4094 representable, absolute floor 2^-29
  std::map<std::string, TcWeight> tcw;
  bool use_tc() const { return std::is_same<T, float>::value && cfg.gemm_backend == DQMC_GEMM_TCGEN05; }
  int prepare_tc_weight(const std::string& name, const float* W, int Kc, int Nc, cudaStream_t st, double wmax) {
    TcWeight& w = tcw[name];
    if (!w.hi) {
      DQ_CHECK(cudaMalloc((void**)&w.hi, sizeof(float) * (size_t)Kc * Nc));
      DQ_CHECK(cudaMalloc((void**)&w.lo, sizeof(float) * (size_t)Kc * Nc));
      w.N = Nc; w.K = Kc; w.BN = tc::pick_bn(Nc);
      if (tc::make_weight_map(&w.mh, w.hi, Nc, Kc, w.BN) || tc::make_weight_map(&w.ml, w.lo, Nc, Kc, w.BN) ||
          tc::make_weight_map(&w.mh2, w.hi, Nc, Kc, w.BN / 2) || tc::make_weight_map(&w.ml2, w.lo, Nc, Kc, w.BN / 2)) {
        err = "cuTensorMapEncodeTiled failed for " + name;
        return 4;
      }
      if (Kc % 64 == 0) {
        DQ_CHECK(cudaMalloc((void**)&w.h16, sizeof(uint16_t) * 2 * (size_t)Kc * Nc));
        const uint16_t* lo16 = w.h16 + (size_t)Kc * Nc;
        if (tc::make_kmajor_map(&w.m16h, w.h16, 2, Nc, Kc, 64, w.BN) || tc::make_kmajor_map(&w.m16l, lo16, 2, Nc, Kc, 64, w.BN)) {
          err = "cuTensorMapEncodeTiled (half planes) failed for " + name;
          return 4;
        }
        w.f16 = true;
        if (Nc <= 256 && Nc % 16 == 0) {
          if (tc::make_kmajor_map(&w.m16h_all, w.h16, 2, Nc, Kc, 64, Nc) || tc::make_kmajor_map(&w.m16l_all, lo16, 2, Nc, Kc, 64, Nc)) {
            err = "cuTensorMapEncodeTiled (whole-N half planes) failed for " + name;
            return 4;
          }
          w.f16_all = true;
        }
        if (Kc == 256 && Nc % 128 == 0 &&
            !tc::make_kmajor_map(&w.m16h_128, w.h16, 2, Nc, Kc, 64, 128) && !tc::make_kmajor_map(&w.m16l_128, lo16, 2, Nc, Kc, 64, 128))
          w.f16_128 = true;
      }
    }
    DQ_LAUNCH(split_transpose_kernel, dim3((Kc * Nc + 255) / 256), dim3(256), 0, st, W, Kc, Nc, w.hi, w.lo);
    if (w.f16) {
      // power-of-two weight scale: the largest |w| of the matrix (of the spin pair for the per-spin heads) lands in [512, 1024)
      int e = 0;
      if (wmax > 0) { std::frexp(wmax, &e); e = 10 - e; }
      e = e > 24 ? 24 : (e < -24 ? -24 : e);
      w.wscale = std::ldexp(1.f, e);
      DQ_LAUNCH(split_transpose_f16_kernel, dim3((Kc * Nc + 255) / 256), dim3(256), 0, st, W, Kc, Nc, w.wscale, w.h16,
                w.h16 + (size_t)Kc * Nc);
    }
    return 0;
  }
#else
  bool use_tc() const { return false; }
#endif

  int init() {
    N = cfg.n_up + cfg.n_down; M = cfg.n_nuc; d = cfg.embedding_dim; K = cfg.n_determinants;
    KN = K * N; H = cfg.n_heads; dh = d / H; T3 = 3 * N;
    BFW = KN * (cfg.backflow_add == 2 ? 2 : 1);
    add_off = cfg.backflow_add == 2 ? KN : 0;
    if (cfg.backflow_add && cfg.kind == DQMC_PAULINET) { err = "additive backflow: linear-head ansatz kinds only"; return 2; }
    if (cfg.kind != DQMC_PSIFORMER && cfg.kind != DQMC_FERMINET && cfg.kind != DQMC_TRANSPSIFORMER &&
        cfg.kind != DQMC_PAULINET) {
      err = "unknown ansatz kind"; return 2;
    }
    trans = cfg.kind == DQMC_TRANSPSIFORMER;
    Mn = trans ? cfg.n_nuc_tokens : 0;
    env_rep = cfg.n_env_per_nuc > 1 ? cfg.n_env_per_nuc : 1;
    if (M * env_rep > 4 * DQMC_MAX_NUC || Mn < 0 || Mn > DQMC_MAX_NUC) { err = "bad config (envelope terms / nuclear tokens)"; return 2; }
    if (H < 1) H = 1;
    if (cfg.kind == DQMC_FERMINET || cfg.kind == DQMC_PAULINET) { H = 1; dh = d; }
    gnn = cfg.kind == DQMC_PAULINET;
    if (gnn && (cfg.jastrow_n > 8 || cfg.backflow_n > 8 || cfg.jastrow_n < 0 || cfg.backflow_n < 0)) { err = "bad MLP depth"; return 2; }
    bf_in = d;
    if (gnn && cfg.backflow_n > 0) bf_in = cfg.backflow_dims[cfg.backflow_n - 1];
    if (M > DQMC_MAX_NUC || d % H != 0 || N < 2) { err = "bad config"; return 2; }
    build_layout();
    DQ_CHECK(cudaSetDevice(device));
    DQ_CHECK(cudaMalloc((void**)&d_params, sizeof(T) * total));
    DQ_CHECK(cudaMalloc((void**)&d_params_t, sizeof(T) * total));
    DQ_CHECK(cudaMalloc((void**)&d_stage, sizeof(double) * total));
    std::vector<T> z(M);
    for (int m = 0; m < M; ++m) z[m] = (T)cfg.z_valence[m];
    DQ_CHECK(cudaMalloc((void**)&d_zval, sizeof(T) * M));
    DQ_CHECK(cudaMemcpy(d_zval, z.data(), sizeof(T) * M, cudaMemcpyHostToDevice));
    {
      std::vector<T> zn(M);
      for (int m = 0; m < M; ++m) zn[m] = (T)cfg.z_nuclear[m];
      DQ_CHECK(cudaMalloc((void**)&d_znuc, sizeof(T) * M));
      DQ_CHECK(cudaMemcpy(d_znuc, zn.data(), sizeof(T) * M, cudaMemcpyHostToDevice));
    }
    DQ_CHECK(cudaMalloc((void**)&d_ecp_mask, sizeof(int) * M));
    DQ_CHECK(cudaMemcpy(d_ecp_mask, cfg.ecp_mask, sizeof(int) * M, cudaMemcpyHostToDevice));
    const int Tm = cfg.ecp_loc_terms;
    if (Tm > 0) {
      std::vector<T> lp((size_t)M * 6 * Tm);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < 3; ++n)
          for (int ab = 0; ab < 2; ++ab)
            for (int t = 0; t < Tm; ++t) lp[(((size_t)m * 3 + n) * 2 + ab) * Tm + t] = (T)cfg.ecp_loc[m][n][ab][t];
      DQ_CHECK(cudaMalloc((void**)&d_ecp_loc, sizeof(T) * lp.size()));
      DQ_CHECK(cudaMemcpy(d_ecp_loc, lp.data(), sizeof(T) * lp.size(), cudaMemcpyHostToDevice));
    }
    const int L = cfg.ecp_nl_lmax_p1, Tn = cfg.ecp_nl_terms;
    if (L > 0 && Tn > 0) {
      std::vector<T> np((size_t)M * L * 2 * Tn);
      std::vector<int> nuc;
      for (int m = 0; m < M; ++m) {
        bool any = false;
        for (int l = 0; l < L; ++l)
          for (int ab = 0; ab < 2; ++ab)
            for (int t = 0; t < Tn; ++t) {
              double v = cfg.ecp_nl[m][l][ab][t];
              np[(((size_t)m * L + l) * 2 + ab) * Tn + t] = (T)v;
              any = any || v != 0.0;
            }
        if (any) nuc.push_back(m);  // nuc_with_nl_pot, gaussian_type_ecp.py:120
      }
      J = (int)nuc.size();
      if (J > 0) {
        DQ_CHECK(cudaMalloc((void**)&d_nl_params, sizeof(T) * np.size()));
        DQ_CHECK(cudaMemcpy(d_nl_params, np.data(), sizeof(T) * np.size(), cudaMemcpyHostToDevice));
        DQ_CHECK(cudaMalloc((void**)&d_nl_nuc, sizeof(int) * J));
        DQ_CHECK(cudaMemcpy(d_nl_nuc, nuc.data(), sizeof(int) * J, cudaMemcpyHostToDevice));
      }
    }
#ifndef DQMC_EMU
    {
      int v = 0;
      DQ_CHECK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
      if (v > 0) n_sms = v;
    }
#endif
    if (const char* ev = std::getenv("DQMC_NSMS")) { int x = std::atoi(ev); if (x >= 1) n_sms = x; }  // test hook
    // opt in to large dynamic shared memory
    const bool psif = cfg.kind == DQMC_PSIFORMER || trans;
    attn_tb = attn_pick_tb<T>(N, dh, T3, 100 * 1024, Mn);
    // the specialised fp32 attention kernels do not take extra tokens yet: TransPsiformer runs the generic one
    attn_f32 = psif && !trans && std::is_same<T, float>::value && dh % 16 == 0 && !std::getenv("DQMC_ATTN_GENERIC");
    if (attn_f32) attn_tb = attn_f32_pick_tb(N, dh, T3, 32 * 1024);  // ~7 blocks/SM for small molecules
    if (const char* ev = std::getenv("DQMC_ATTN_TB")) { int x = std::atoi(ev); if (x >= 1 && x <= T3) attn_tb = x; }
    if (const char* ev = std::getenv("DQMC_ATTN_NT")) { int x = std::atoi(ev); if (x >= 32 && x <= 1024 && x % 32 == 0) attn_fl_threads = x; }
    size_t s_attn = !psif ? 0 : attn_f32 ? attn_f32_smem_bytes(N, dh, attn_tb) : attn_smem_bytes<T>(N, dh, attn_tb, Mn);
    // few resident blocks per SM (large molecules: the tangent chunk fills the shared memory) -> more warps per block
    // (benzene, one 170 KB block per SM: 128 threads 707 ms per 512-walker step, 512 threads 684 ms)
    if (!std::getenv("DQMC_ATTN_NT")) attn_fl_threads = s_attn > 110 * 1024 ? 512 : (s_attn > 56 * 1024 ? 256 : 128);
    // 16 x 8 tiles: worthwhile from about 20 electrons (benzene, N = 30: 686 -> 648 ms per 512-walker step; LiH, N = 4, where
    // 7/8 of every tile is padding: 6.9 -> 15.7 ms, so small molecules keep the SIMT variant)
    attn_fl_mma = attn_f32 && N >= 20 && N <= 32 && dh % 16 == 0;
    if (const char* ev = std::getenv("DQMC_ATTN_FL_MMA")) attn_fl_mma = attn_f32 && N <= 32 && dh % 16 == 0 && std::atoi(ev) != 0;
    if (attn_fl_mma && !std::getenv("DQMC_ATTN_TB")) {  // two resident blocks of 8 warps per SM
      attn_tb = attn_f32_pick_tb(N, dh, T3, 110 * 1024);
      s_attn = attn_f32_smem_bytes(N, dh, attn_tb);
    }
    // generic kernel (extra key / value tokens: TransPsiformer) in fp32: same tensor-core products, row pitch dh + 4
    attn_gen_mma = psif && !attn_f32 && std::is_same<T, float>::value && N >= 20 && dh % 16 == 0 && !std::getenv("DQMC_ATTN_GENERIC");
    if (const char* ev = std::getenv("DQMC_ATTN_FL_MMA"))
      attn_gen_mma = psif && !attn_f32 && std::is_same<T, float>::value && dh % 16 == 0 && std::atoi(ev) != 0;
    if (attn_gen_mma) {
      if (!std::getenv("DQMC_ATTN_TB")) attn_tb = attn_pick_tb<T>(N, dh, T3, 110 * 1024, Mn, 4);
      s_attn = attn_smem_bytes<T>(N, dh, attn_tb, Mn, 4);
    }
    size_t s_sl = slater_smem_bytes<T>(N);
    max_smem = s_attn > s_sl ? s_attn : s_sl;
    if (max_smem > 227 * 1024) { err = "system too large for the shared-memory tiling (N)"; return 2; }
    if (attn_f32) {
      if (launch_attn_f32(nullptr, nullptr, 0, 0, 0, 0.f, 0, (int)s_attn, nullptr, true)) return 1;
    } else if (psif) {
      DQ_CHECK(raise_dyn_smem(attn_fl_kernel<T>, (int)s_attn));
      if constexpr (std::is_same<T, float>::value) DQ_CHECK(raise_dyn_smem((attn_fl_kernel<T, true>), (int)s_attn));
    }
    DQ_CHECK(raise_dyn_smem(slater_kernel<T>, (int)s_sl));
    slater_fwd2_ok = N <= 32 && slater_fwd2_smem_bytes<T>(N, M, K) <= 110 * 1024 && !std::getenv("DQMC_SLATER_GENERIC") &&
                     !std::getenv("DQMC_SLATER_FWD1");
    if (slater_fwd2_ok) {
      DQ_CHECK(raise_dyn_smem(slater_fwd2_kernel<T, 14>, (int)slater_fwd2_smem_bytes<T>(N, M, K)));
      DQ_CHECK(raise_dyn_smem(slater_fwd2_kernel<T, 16>, (int)slater_fwd2_smem_bytes<T>(N, M, K)));
      DQ_CHECK(raise_dyn_smem(slater_fwd2_kernel<T, 28>, (int)slater_fwd2_smem_bytes<T>(N, M, K)));
      DQ_CHECK(raise_dyn_smem(slater_fwd2_kernel<T, 30>, (int)slater_fwd2_smem_bytes<T>(N, M, K)));
      DQ_CHECK(raise_dyn_smem(slater_fwd2_kernel<T, 32>, (int)slater_fwd2_smem_bytes<T>(N, M, K)));
    }
    attn_fwd_ok = psif && std::is_same<T, float>::value && dh == 64 && N <= 32 && N + Mn <= 48 && d % 4 == 0 &&
                  !std::getenv("DQMC_ATTN_GENERIC") && !std::getenv("DQMC_ATTN_FWD_OLD");
    attn_fwd_pipelined = attn_fwd_ok && !trans && std::getenv("DQMC_ATTN_FWD2");  // measured slower than the block-per-walker kernel
    if (attn_fwd_pipelined) {
      const int smem2 = 6 * 4 * N * 64 * (int)sizeof(float);
      DQ_CHECK(raise_dyn_smem(attn_fwd2_f32_kernel<8>, smem2));
      DQ_CHECK(raise_dyn_smem(attn_fwd2_f32_kernel<16>, smem2));
      DQ_CHECK(raise_dyn_smem(attn_fwd2_f32_kernel<32>, smem2));
    }
    if (attn_fwd_ok) {
      const int smem = 4 * 2 * (N + Mn) * 64 * (int)sizeof(float);
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<48, false>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<8, false>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<16, false>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<32, false>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<4, true>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<10, true>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<14, true>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<28, true>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fwd_f32_kernel<30, true>), smem));
    }
    attn_mma_ok = psif && std::is_same<T, float>::value && dh == 64 && N + Mn <= 48 && d % 4 == 0 && !std::getenv("DQMC_ATTN_GENERIC") &&
                  !(std::getenv("DQMC_ATTN_MMA") && std::atoi(std::getenv("DQMC_ATTN_MMA")) == 0);
    embed_fwd_ok = psif && d % 4 == 0 && embed_fwd_smem_bytes<T>(M, d) <= 200 * 1024 && !std::getenv("DQMC_EMBED_GENERIC");
    if (embed_fwd_ok)
      DQ_CHECK(raise_dyn_smem(embed_fwd_kernel<T>, (int)embed_fwd_smem_bytes<T>(M, d)));
    if (cfg.gemm_backend == DQMC_GEMM_TCGEN05) {
#if !defined(DQMC_NO_TCGEN05)
      if (!std::is_same<T, float>::value) { err = "DQMC_GEMM_TCGEN05 needs dtype DQMC_F32"; return 2; }
      if (d % 32 != 0) { err = "DQMC_GEMM_TCGEN05 needs embedding_dim % 32 == 0"; return 2; }
      DQ_CHECK(raise_dyn_smem((tc::gemm3xtf32_kernel<false, false>), tc::SmemLayout::total(256)));
      DQ_CHECK(raise_dyn_smem((tc::gemm3xtf32_kernel<false, true>), tc::SmemLayout::total(256)));
      DQ_CHECK(raise_dyn_smem(tc::mlp_block_f16_kernel, tc::MlpSmem::total()));
      DQ_CHECK(raise_dyn_smem(tc::trunk_f16_kernel<false>, tc::TrSmem::total()));
      DQ_CHECK(raise_dyn_smem(tc::trunk_f16_kernel<true>, tc::TrSmem::total()));
#ifndef DQMC_EMU
      DQ_CHECK(raise_dyn_smem((tc::gemm3xtf32_kernel<true, false>), tc::SmemLayoutT<true>::total(256)));
      gemm_2cta = std::getenv("DQMC_GEMM_2CTA") != nullptr;
#endif
      if (const char* ev = std::getenv("DQMC_TC_F16")) f16_on = std::atoi(ev) != 0;
      if (const char* ev = std::getenv("DQMC_TC_FUSE_MLP")) fuse_mlp = std::atoi(ev) != 0;
      if (const char* ev = std::getenv("DQMC_TC_TRUNK")) fuse_trunk = std::atoi(ev) != 0;
      if (const char* ev = std::getenv("DQMC_TC_TRUNK_TS")) trunk_ts = std::atoi(ev) != 0;
      if (psif && !trans && d == 256 && H == 4 && N <= 32 && cfg.n_layers <= tc::kTrMaxLayers) {
        DQ_CHECK(cudaMalloc((void**)&d_trunk_maps, sizeof(CUtensorMap) * 8 * cfg.n_layers));
        DQ_CHECK(cudaMalloc((void**)&d_trunk_scratch, (size_t)n_sms * tc::kTrScratchPerCta));
        if (std::getenv("DQMC_TRUNK_TRACE")) {
          DQ_CHECK(cudaMalloc((void**)&d_trunk_trace, sizeof(long long) * 64));
          DQ_CHECK(cudaMemset(d_trunk_trace, 0, sizeof(long long) * 64));
        }
      }
#else
      err = "this build has no tcgen05 backend"; return 2;
#endif
    }
    return 0;
  }
  ~Engine() override {
    if (plan_only) return;  // nothing was allocated, no CUDA context
    cudaFree(d_params); cudaFree(d_params_t); cudaFree(d_stage); cudaFree(d_znuc); cudaFree(d_zval); cudaFree(d_ecp_mask);
    if (d_ecp_loc) cudaFree(d_ecp_loc);
    if (d_nl_params) cudaFree(d_nl_params);
    if (d_nl_nuc) cudaFree(d_nl_nuc);
    if (d_ph_tabs) cudaFree(d_ph_tabs);
    if (d_ph_nuc) cudaFree(d_ph_nuc);
#if !defined(DQMC_NO_TCGEN05)
    for (auto& kv : tcw) {
      if (kv.second.hi) cudaFree(kv.second.hi);
      if (kv.second.lo) cudaFree(kv.second.lo);
      if (kv.second.h16) cudaFree(kv.second.h16);
    }
    if (d_trunk_maps) cudaFree(d_trunk_maps);
    if (d_trunk_scratch) cudaFree(d_trunk_scratch);
#endif
  }
  const T* P(const std::string& n) const { return d_params + off(n); }

  int set_ph(int n_tab, int n_grid, double r_max, const double* tables, const int32_t* tab_of_nuc) override {
    if (J > 0 || cfg.ecp_loc_terms > 0) { err = "pseudo-Hamiltonian and Gaussian-type ECP are mutually exclusive"; return 2; }
    if (n_tab < 1 || n_grid < 2 || !(r_max > 0) || !tables || !tab_of_nuc) { err = "bad pseudo-Hamiltonian tables"; return 2; }
    for (int m = 0; m < M; ++m)
      if (tab_of_nuc[m] >= n_tab) { err = "pseudo-Hamiltonian table index out of range"; return 2; }
    std::vector<T> h((size_t)n_tab * 2 * n_grid);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (T)tables[i];
    if (d_ph_tabs) cudaFree(d_ph_tabs);
    if (d_ph_nuc) cudaFree(d_ph_nuc);
    DQ_CHECK(cudaMalloc((void**)&d_ph_tabs, sizeof(T) * h.size()));
    DQ_CHECK(cudaMemcpy(d_ph_tabs, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    DQ_CHECK(cudaMalloc((void**)&d_ph_nuc, sizeof(int) * M));
    DQ_CHECK(cudaMemcpy(d_ph_nuc, tab_of_nuc, sizeof(int) * M, cudaMemcpyHostToDevice));
    ph_G = n_grid; ph_rmax = r_max; ph_on = true;
    return 0;
  }
  PhArgs<T> ph_args(const T* QA) const {
    PhArgs<T> a;
    a.QA = QA; a.tabs = d_ph_tabs; a.tab_of_nuc = d_ph_nuc; a.G = ph_G; a.rmax = (T)ph_rmax;
    return a;
  }

  int set_params(const double* host, int64_t n, cudaStream_t st) override {
    if (n != total) { err = "parameter count mismatch"; return 2; }
    DQ_CHECK(cudaMemcpyAsync(d_stage, host, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    DQ_LAUNCH(convert_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const double*)d_stage, d_params, n);
    for (auto& e : entries)  // (vectors included: a [n][1] weight such as the Jastrow's last layer is its own transpose)
      if (e.rows >= 1 && e.cols >= 1)
        DQ_LAUNCH(transpose_kernel<T>, dim3((e.rows * e.cols + 255) / 256), dim3(256), 0, st, (const T*)(d_params + e.offset),
                  e.rows, e.cols, d_params_t + e.offset);
#if !defined(DQMC_NO_TCGEN05)
    if (use_tc()) {
      // largest |w| per matrix (the two spin heads of a per-spin layer share one scale: they are used in one launch)
      std::map<std::string, double> wmax;
      auto group = [](const std::string& n) {
        const size_t k = n.size();
        return (k > 3 && (n.compare(k - 3, 3, ".up") == 0 || n.compare(k - 3, 3, ".dn") == 0)) ? n.substr(0, k - 3) : n;
      };
      for (auto& e : entries) {
        double m = 0;
        for (int64_t i = 0; i < (int64_t)e.rows * e.cols; ++i) m = std::max(m, std::fabs(host[e.offset + i]));
        double& g = wmax[group(e.name)];
        g = std::max(g, m);
      }
      for (auto& e : entries) {
        bool is_w = e.name.find(".w") != std::string::npos || e.name.rfind("bf.", 0) == 0;
        if (!is_w || e.name == "emb.w" || e.rows % 32 != 0 || e.cols < 64) continue;
        int rc = prepare_tc_weight(e.name, (const float*)(d_params + e.offset), e.rows, e.cols, st, wmax[group(e.name)]);
        if (rc) return rc;
      }
      if (d_trunk_maps && trunk_weights_ok()) {
        std::vector<CUtensorMap> hm((size_t)8 * cfg.n_layers);
        for (int l = 0; l < cfg.n_layers; ++l) {
          const std::string pfx = "L" + std::to_string(l) + ".";
          int g = 0;
          for (const char* n : {"wqkv", "wo", "w1", "w2"}) {
            const TcWeight& w = tcw.at(pfx + n);
            hm[(size_t)8 * l + 2 * g] = w.m16h_128;
            hm[(size_t)8 * l + 2 * g + 1] = w.m16l_128;
            ++g;
          }
        }
        DQ_CHECK(cudaMemcpy(d_trunk_maps, hm.data(), sizeof(CUtensorMap) * hm.size(), cudaMemcpyHostToDevice));
      }
    }
#endif
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  // ---- workspace ---------------------------------------------------------------------------
  struct Ws {
    T *X, *O, *A, *M1, *QKV, *BF, *dsign, *dlog, *dlap, *dgrad;
    T* QA = nullptr;  // pseudo-Hamiltonian records [Bc][N][PH_STRIDE]
    T* Gadd = nullptr;  // additive-backflow factor g_i = cutoff * envelope norm with gradient / Laplacian [Bc][N][5]
    T *G0 = nullptr, *G1 = nullptr, *G2 = nullptr, *Hs = nullptr, *Ha = nullptr, *C = nullptr, *Fc = nullptr, *HT = nullptr,
      *E0 = nullptr, *E1 = nullptr, *ET0 = nullptr, *ET1 = nullptr, *W3 = nullptr,
      *Y0 = nullptr, *Y1 = nullptr, *Jb = nullptr;  // conv-GNN trunk
    size_t bytes;
  };
  int gnn_hmax() const {
    int h = 1;
    for (int i = 0; i < cfg.backflow_n; ++i) h = cfg.backflow_dims[i] > h ? cfg.backflow_dims[i] : h;
    return h;
  }
  size_t fermi_dmax() const { return (size_t)(4 * M > d ? 4 * M : d); }
  size_t fermi_emax() const { return (size_t)(cfg.edge_dim > 4 ? cfg.edge_dim : 4); }
  int gnn_dmax() const { return cfg.gnn_features && 4 * M > d ? 4 * M : d; }
  int gnn_emax() const {  // widest edge-side row (raw features, w / u hidden and output widths)
    int m = cfg.edge_dim > 4 ? cfg.edge_dim : 4;
    for (int l = 0; l < cfg.n_layers && l < 8; ++l)
      for (int i = 0; i < cfg.gnn_sub_n && i < 4; ++i) {
        m = cfg.gnn_w_dims[l][i] > m ? cfg.gnn_w_dims[l][i] : m;
        m = cfg.gnn_u_dims[l][i] > m ? cfg.gnn_u_dims[l][i] : m;
      }
    return m;
  }
  int gnn_hnode_max() const {
    int m = cfg.edge_dim;
    for (int l = 0; l < cfg.n_layers && l < 8; ++l)
      for (int i = 0; i < cfg.gnn_sub_n && i < 4; ++i) m = cfg.gnn_h_dims[l][i] > m ? cfg.gnn_h_dims[l][i] : m;
    return m;
  }
  int gnn_jsum() const {
    int j = d;
    for (int i = 0; i < cfg.jastrow_n; ++i) j += cfg.jastrow_dims[i];
    return j;
  }
  // Workspace plan == the carve itself: chunk_bytes() runs carve() on a dummy base, so a buffer added to carve() can never
  // be forgotten in the plan (round 1 planned dgrad for S > 1 only while carve() always took it).
  static char* plan_base() { return (char*)(uintptr_t)0x100000; }
  size_t chunk_bytes(int Bc, int S) const {
    const bool was = dry;
    const char* hw = dry_hwm;  // a size probe is not a carve of the caller's workspace
    const_cast<Engine*>(this)->dry = true;  // no guard writes through the dummy base
    const size_t n = carve(plan_base(), Bc, S).bytes;
    const_cast<Engine*>(this)->dry = was;
    dry_hwm = hw;
    return n;
  }
  Ws carve(void* base, int Bc, int S) const {
    Ws w;
    size_t rows = (size_t)Bc * N * S;
    char* p = (char*)base;
    auto take = [&](size_t n) { T* q = (T*)p; p += align_up(sizeof(T) * n); DQ_TAKE_GUARD(); return q; };
    if (gnn) {
      const size_t e = cfg.edge_dim, hm = gnn_hmax(), dm = gnn_dmax(), em = gnn_emax(), hn = gnn_hnode_max();
      const size_t pairs8 = (size_t)Bc * N * (N + (cfg.gnn_conv_ne ? M : 0)) * 8;
      w.X = take(rows * dm); w.O = take(rows * dm); w.G0 = take(rows * d); w.G1 = take(rows * d); w.G2 = take(rows * d);
      w.Fc = take(rows * (3 * dm + 3 * e)); w.Hs = take(rows * e); w.Ha = take(rows * e); w.HT = take(rows * hn);
      w.C = take(rows * 3 * e);
      w.E0 = take(pairs8 * em); w.E1 = take(pairs8 * em); w.ET0 = take(pairs8 * em); w.ET1 = take(pairs8 * em);
      w.W3 = take(pairs8 * 3 * e);
      w.Y0 = take(rows * hm); w.Y1 = take(rows * hm); w.Jb = take((size_t)Bc * S * gnn_jsum());
      w.A = w.M1 = w.QKV = nullptr;
      w.BF = take(rows * KN);
    } else if (cfg.kind == DQMC_FERMINET) {
      const size_t dm = fermi_dmax(), em = fermi_emax(), fin = 3 * dm + 2 * em;
      w.X = take(rows * dm); w.O = take(rows * dm); w.QKV = take(rows * fin);  // H, H2, F
      w.A = take(rows * N * em); w.M1 = take(rows * N * em);                   // E, E2
      w.BF = take(rows * BFW);
    } else {
      w.X = take(rows * d); w.O = take(rows * d); w.A = take(rows * d); w.M1 = take(rows * d);
      w.QKV = take(rows * 3 * d); w.BF = take(rows * BFW);
    }
    w.dsign = take((size_t)Bc * K); w.dlog = take((size_t)Bc * K); w.dlap = take((size_t)Bc * K);
    w.dgrad = take((size_t)Bc * K * (S > 1 ? T3 : 1));
    if (ph_on && S > 1) w.QA = take((size_t)Bc * N * PH_STRIDE);
    if (cfg.backflow_add) w.Gadd = take((size_t)Bc * N * 5);
    w.bytes = p - (char*)base;
    note_hwm(p);
    return w;
  }
  // largest walker chunk (<= B, <= the 32-bit row cap) whose carve fits wsb bytes
  int max_chunk(int64_t wsb, int S, int B) const {
    int64_t c = B;
    int64_t row_cap = (int64_t)2000000000 / ((int64_t)N * S * 3 * d);  // keep 32-bit row*ld products safe
    if (cfg.kind == DQMC_FERMINET) row_cap = (int64_t)2000000000 / ((int64_t)N * N * S * (3 * (int64_t)fermi_dmax() + 64));
    if (gnn) row_cap = (int64_t)2000000000 / ((int64_t)N * (N + M + S) * (8 * gnn_emax() + 3 * gnn_dmax() + 3 * cfg.edge_dim + KN));
    if (c > row_cap) c = row_cap;
    if (c < 1 || (int64_t)chunk_bytes(1, S) > wsb) return 0;
    if ((int64_t)chunk_bytes((int)c, S) <= wsb) return (int)c;
    int64_t lo = 1, hi = c;  // chunk_bytes is monotone in the chunk size
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) / 2;
      if ((int64_t)chunk_bytes((int)mid, S) <= wsb) lo = mid; else hi = mid;
    }
    return (int)lo;
  }
  // non-local ECP pass for nb walkers: virtual walkers r_virt[V][N][3], sign[V], log[V] + one
  // forward chunk over all V = nb * J * N * 12 virtual walkers
  int64_t ecp_prefix_bytes(int64_t nb) const {
    const int64_t V = nb * J * N * 12;
    return (int64_t)align_up(sizeof(T) * V * 3 * N) + 2 * (int64_t)align_up(sizeof(T) * V) +
           (int64_t)align_up(sizeof(T) * nb * N * K * N) +  // + envelope table of the group's base walkers
           (int64_t)align_up(sizeof(T) * nb * N * d);        // + their embedding rows
  }
  // walkers per ECP group are bounded by the 32-bit row cap of the plain-forward chunk: the plan never asks for more
  int64_t ecp_group_cap() const {
    const int64_t vper = (int64_t)J * N * 12;
    const int64_t vcap = 2000000000LL / ((int64_t)N * 3 * d);
    return std::max<int64_t>(1, vcap / vper);
  }
  int64_t ecp_bytes(int64_t nb) const {
    const int64_t V = nb * J * N * 12;
    return ecp_prefix_bytes(nb) + (int64_t)chunk_bytes((int)V, 1);
  }
  int64_t mcmc_prefix_bytes(int B) const {
    return (int64_t)align_up(sizeof(T) * (size_t)B * 3 * N) + 2 * (int64_t)align_up(sizeof(T) * (size_t)B) + 256;
  }
  int64_t force_prefix_bytes(int B) const {  // value_and_force: E, 6 stats, grad
    return (int64_t)(align_up(sizeof(T) * (size_t)B) + align_up(sizeof(T) * (size_t)6 * B) + align_up(sizeof(T) * (size_t)B * T3));
  }
  int64_t langevin_prefix_bytes(int B) const {
    return 2 * (int64_t)align_up(sizeof(T) * (size_t)B * 3 * N) + 2 * (int64_t)align_up(sizeof(T) * (size_t)B) + 256;
  }
  // bytes one reverse-pass chunk of Bc walkers carves: a dry pass of the chunk function itself
  int64_t vjp_chunk_bytes(int Bc) {
    const bool was = dry;
    const char* hw = dry_hwm;
    const int64_t cap = vjp_ws_cap;
    dry = true; dry_hwm = plan_base(); vjp_ws_cap = INT64_MAX;
    if (gnn) vjp_chunk_paulinet(nullptr, nullptr, 0, Bc, nullptr, nullptr, nullptr, nullptr, plan_base(), nullptr);
    else if (cfg.kind == DQMC_FERMINET) vjp_chunk_ferminet(nullptr, nullptr, 0, Bc, nullptr, nullptr, nullptr, nullptr, plan_base(), nullptr);
    else vjp_chunk(nullptr, nullptr, 0, Bc, nullptr, nullptr, nullptr, nullptr, plan_base(), nullptr);
    const int64_t n = dry_hwm - plan_base();
    dry = was; dry_hwm = hw; vjp_ws_cap = cap;
    return n;
  }
  int64_t ws_bytes(int B, int mode) override {
    if (B < 1) B = 1;
    if (mode == DQMC_MODE_VJP) return vjp_chunk_bytes(B);
    if (mode == DQMC_MODE_MCMC) return mcmc_prefix_bytes(B) + (int64_t)chunk_bytes(B, 1);
    if (mode == DQMC_MODE_LANGEVIN) return langevin_prefix_bytes(B) + force_prefix_bytes(B) + (int64_t)chunk_bytes(B, T3 + 2);
    int S = mode == DQMC_MODE_FORWARD ? 1 : T3 + 2;
    int64_t need = (int64_t)chunk_bytes(B, S);
    if (mode == DQMC_MODE_LOCAL_ENERGY && J > 0) need = std::max<int64_t>(need, ecp_bytes(std::min<int64_t>(B, ecp_group_cap())));
    return need;
  }
  // least workspace with which a call for B walkers proceeds (walkers chunked down to one at a time)
  int64_t ws_bytes_min(int B, int mode) override {
    if (B < 1) B = 1;
    const int S = T3 + 2;
    switch (mode) {
      case DQMC_MODE_VJP: return vjp_chunk_bytes(1);
      case DQMC_MODE_MCMC: return mcmc_prefix_bytes(B) + (int64_t)chunk_bytes(1, 1);
      case DQMC_MODE_LANGEVIN: return langevin_prefix_bytes(B) + force_prefix_bytes(B) + (int64_t)chunk_bytes(1, S);
      case DQMC_MODE_LOCAL_ENERGY:
        return std::max<int64_t>((int64_t)chunk_bytes(1, S), J > 0 ? ecp_prefix_bytes(1) + (int64_t)chunk_bytes(1, 1) : 0);
      default: return (int64_t)chunk_bytes(1, 1);
    }
  }
  // dqmc_debug_plan: walk the entry point of `mode` with a workspace of wsb bytes (<= 0: the planned size) on a dummy base
  // and report the highest offset it carves.  Host-only (works on plan-only engines).
  int debug_plan(int B, int mode, int64_t wsb, int64_t* planned, int64_t* carved) override {
    const int64_t pl = ws_bytes(B, mode);
    if (planned) *planned = pl;
    if (wsb <= 0) wsb = pl;
    const bool was = dry;
    dry = true; dry_hwm = plan_base();
    void* ws = plan_base();
    int rc = 0;
    switch (mode) {
      case DQMC_MODE_FORWARD: rc = forward(nullptr, nullptr, 0, B, nullptr, nullptr, ws, wsb, nullptr); break;
      case DQMC_MODE_LOCAL_ENERGY:
        rc = local_energy(nullptr, nullptr, 0, B, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ws, wsb, nullptr);
        break;
      case DQMC_MODE_VJP: rc = vjp_params(nullptr, nullptr, 0, B, nullptr, nullptr, nullptr, nullptr, ws, wsb, nullptr); break;
      case DQMC_MODE_MCMC:
        rc = mcmc(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, 1, 0.5, -1, 0, 0, 0, nullptr, nullptr, nullptr, ws, wsb,
                  nullptr, 0.0, nullptr, nullptr);
        break;
      case DQMC_MODE_LANGEVIN:
        rc = langevin(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, B, 1, 0.5, -1, 0, 0, 0, nullptr, nullptr,
                      nullptr, ws, wsb, nullptr);
        break;
      default: err = "unknown mode"; rc = 2;
    }
    if (carved) *carved = dry_hwm - plan_base();
    dry = was;
    return rc;
  }

  // ---- GEMM dispatch ------------------------------------------------------------------------
  // dense layers can absorb the tanh propagation in the tensor-core epilogue when whole slot
  // groups fit a 128-row tile with little padding
  bool can_fuse_act(int S) const {
    if (!use_tc() || std::getenv("DQMC_NO_FUSE_TANH")) return false;
    if (S == 1) return true;
    return S <= 128 && (128 / S) * S >= 112;
  }
  int gemm(const T* A, int lda, const char* w0, const char* w1, int zsplit, int ldw, const T* bias, const T* Res,
           int ldr, T* C, int ldc, int Mr, int Nc, int Kc, int S, int sliced, int Nel, cudaStream_t st, int act = 0,
           const T* bias1 = nullptr) {
    if (dry) return 0;  // planning pass
    const T* W0 = P(w0);
    const T* W1 = w1 ? P(w1) : nullptr;
#if !defined(DQMC_NO_TCGEN05)
    if constexpr (std::is_same<T, float>::value) {
      if (use_tc() && !bias1 && Kc % 32 == 0 && lda % 4 == 0 && ldc % 4 == 0 && tcw.count(w0) && (!w1 || tcw.count(w1))) {
        const TcWeight& t0 = tcw.at(w0);
        const TcWeight& t1 = w1 ? tcw.at(w1) : t0;
        tc::Params p;
        p.A = A; p.lda = lda; p.bias = bias; p.Res = Res; p.ldr = ldr; p.C = C; p.ldc = ldc; p.M = Mr; p.N = Nc;
        p.K = Kc; p.S = S; p.sliced = sliced; p.Nel = Nel; p.z_split = zsplit; p.BN = t0.BN; p.err_flag = nullptr;
        p.act = act;
        p.rpt = (act && S > 1) ? (128 / S) * S : tc::kBM;
        p.a_scale = 1.f; p.unscale = 1.f;
        // plain forwards: half operands (hi / lo), kind::f16 -- twice the MMA rate, half the shared-memory bytes per k
        const bool f16 = S == 1 && f16_on && Kc % 64 == 0 && t0.f16 && t1.f16 && t0.wscale == t1.wscale && !gemm_2cta;
        if (f16) { p.a_scale = kActScale; p.unscale = 1.f / (kActScale * t0.wscale); }
        int MT = (Mr + p.rpt - 1) / p.rpt, NT = (Nc + p.BN - 1) / p.BN;
        int n_tiles = (sliced ? Nel : 1) * MT * NT;
        int grid = n_tiles < n_sms ? n_tiles : n_sms;
#ifndef DQMC_EMU
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
        if (gemm_2cta && n_sms >= 2) {
          // CTA pairs: clusters of 2, each pair owns 256-row tiles; weight maps with half-tile boxes
          const int MT2 = (MT + 1) / 2;
          const int n_pt = (sliced ? Nel : 1) * MT2 * NT;
          int pairs = n_sms / 2;
          if (n_pt < pairs) pairs = n_pt;
          cudaLaunchConfig_t lc = {};
          lc.gridDim = dim3(2 * pairs); lc.blockDim = dim3(tc::kThreads);
          lc.dynamicSmemBytes = tc::SmemLayoutT<true>::total(p.BN); lc.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          lc.attrs = at; lc.numAttrs = 1;
          DQ_CHECK(cudaLaunchKernelEx(&lc, tc::gemm3xtf32_kernel<true, false>, t0.mh2, t0.ml2, t1.mh2, t1.ml2, p));
          ++launches;
        } else
#endif
        if (f16)
          DQ_LAUNCH((tc::gemm3xtf32_kernel<false, true>), dim3(grid), dim3(tc::kThreads), tc::SmemLayout::total(p.BN), st, t0.m16h,
                    t0.m16l, t1.m16h, t1.m16l, p);
        else
          DQ_LAUNCH((tc::gemm3xtf32_kernel<false, false>), dim3(grid), dim3(tc::kThreads), tc::SmemLayout::total(p.BN), st, t0.mh,
                    t0.ml, t1.mh, t1.ml, p);
#ifndef DQMC_EMU
        if (prof) {
          cudaEventRecord(e1, st);
          prof_note(e0, e1, 2.0 * (double)Mr * (sliced ? Nel : 1) * (double)Nc * (double)Kc, 0);
        }
#endif
        return 0;
      }
    }
#endif
    if (act) { err = "internal: fused activation requested on the CUDA-core GEMM"; return 5; }
    GemmArgs<T> g;
    g.A = A; g.lda = lda; g.W0 = W0; g.W1 = W1; g.z_split = zsplit; g.ldw = ldw; g.bias = bias; g.Res = Res;
    g.ldr = ldr; g.C = C; g.ldc = ldc; g.M = Mr; g.N = Nc; g.K = Kc; g.S = S; g.sliced = sliced; g.Nel = Nel;
    g.bias1 = bias1;
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((Nc + BN - 1) / BN, (Mr + BM - 1) / BM, sliced ? Nel : 1);
#ifndef DQMC_EMU
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
#endif
    DQ_LAUNCH((gemm_kernel<T, BM, BN, BK, TM, TN>), grid, dim3(256), 0, st, g);
#ifndef DQMC_EMU
    if (prof) {
      cudaEventRecord(e1, st);
      prof_note(e0, e1, 2.0 * (double)Mr * (sliced ? Nel : 1) * (double)Nc * (double)Kc, 0);
    }
#endif
    return 0;
  }

  // Fused MLP block of a plain forward: Out = A + tanh(tanh(A W1 + b1) W2 + b2), A = X + O Wo  (one launch; fused_tc.cuh)
  bool can_fuse_mlp(int S, const std::string& pfx) const {
#if !defined(DQMC_NO_TCGEN05)
    if (S != 1 || !use_tc() || !f16_on || !fuse_mlp || (d != 128 && d != 256)) return false;
    for (const char* n : {"wo", "w1", "w2"}) {
      auto it = tcw.find(pfx + n);
      if (it == tcw.end() || !it->second.f16_all) return false;
    }
    return true;
#else
    return false;
#endif
  }
  int mlp_block(const std::string& pfx, const T* O, const T* X, T* Out, int rows, cudaStream_t st) {
    if (dry) return 0;
#if !defined(DQMC_NO_TCGEN05)
    if constexpr (std::is_same<T, float>::value) {
      const TcWeight& wo = tcw.at(pfx + "wo");
      const TcWeight& w1 = tcw.at(pfx + "w1");
      const TcWeight& w2 = tcw.at(pfx + "w2");
      tc::MlpParams p;
      p.O = O; p.ldo = d; p.X = X; p.ldx = d; p.Out = Out; p.ldout = d; p.b1 = P(pfx + "b1"); p.b2 = P(pfx + "b2");
      p.M = rows; p.d = d; p.a_scale = kActScale;
      p.us0 = 1.f / (kActScale * wo.wscale); p.us1 = 1.f / (kActScale * w1.wscale); p.us2 = 1.f / (kActScale * w2.wscale);
      p.err_flag = nullptr;
      const int MT = (rows + 127) / 128;
      const int grid = MT < n_sms ? MT : n_sms;
#ifndef DQMC_EMU
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
#endif
      DQ_LAUNCH(tc::mlp_block_f16_kernel, dim3(grid), dim3(tc::kMlpThreads), tc::MlpSmem::total(), st, wo.m16h_all, wo.m16l_all,
                w1.m16h_all, w1.m16l_all, w2.m16h_all, w2.m16l_all, p);
#ifndef DQMC_EMU
      if (prof) {
        cudaEventRecord(e1, st);
        prof_note(e0, e1, 3 * 2.0 * (double)rows * (double)d * (double)d, 1);
      }
#endif
      return 0;
    }
#endif
    err = "internal: fused MLP block without the tensor-core backend";
    return 5;
  }

  int debug_mlp_block(int layer, const void* O, const void* X, void* Out, int rows, cudaStream_t st) override {
    const std::string pfx = "L" + std::to_string(layer) + ".";
    if (!can_fuse_mlp(1, pfx)) { err = "fused MLP block not available for this configuration"; return 2; }
    int rc = mlp_block(pfx, (const T*)O, (const T*)X, (T*)Out, rows, st);
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  // Whole trunk of a plain forward in one launch (trunk_tc.cuh): needs the shipped Psiformer shape (d = 256, 4 heads of 64)
  bool trunk_weights_ok() const {
#if !defined(DQMC_NO_TCGEN05)
    for (int l = 0; l < cfg.n_layers; ++l)
      for (const char* n : {"wqkv", "wo", "w1", "w2"}) {
        auto it = tcw.find("L" + std::to_string(l) + "." + n);
        if (it == tcw.end() || !it->second.f16_128) return false;
      }
    return true;
#else
    return false;
#endif
  }
  bool can_trunk(int S) const {
#if !defined(DQMC_NO_TCGEN05)
    return S == 1 && use_tc() && f16_on && fuse_trunk && d_trunk_maps && d_trunk_scratch && cfg.n_layers >= 1 && trunk_weights_ok();
#else
    return false;
#endif
  }
  int trunk_block(const T* X0, T* Out, int rows, cudaStream_t st, const T* Xbase = nullptr) {
    if (dry) return 0;
#if !defined(DQMC_NO_TCGEN05)
    if constexpr (std::is_same<T, float>::value) {
      tc::TrunkParams p;
      p.X0 = X0; p.ldx = d; p.Out = Out; p.ldout = d; p.maps = d_trunk_maps; p.scratch = d_trunk_scratch;
      p.Xbase = Xbase; p.v0 = Xbase ? (long long)ecp_v0 : 0; p.vper = Xbase ? ecp_vper : 0;
      int np2 = 1;
      while (np2 < N) np2 *= 2;  // walker slot of the tile: electrons rounded up to a power of two (<= 32)
      p.walkers = rows / N; p.N = N; p.NP = np2; p.L = cfg.n_layers; p.a_scale = kActScale;
      p.attn_scale = (float)(1.0 / std::sqrt((double)dh)); p.err_flag = nullptr; p.trace = d_trunk_trace;
      p.ablate = 0;
      if (const char* ev = std::getenv("DQMC_TRUNK_ABLATE")) p.ablate = std::atoi(ev);
      for (int l = 0; l < cfg.n_layers; ++l) {
        const std::string pfx = "L" + std::to_string(l) + ".";
        p.b1[l] = P(pfx + "b1"); p.b2[l] = P(pfx + "b2");
        int g = 0;
        for (const char* n : {"wqkv", "wo", "w1", "w2"}) p.us[l][g++] = 1.f / (kActScale * tcw.at(pfx + n).wscale);
      }
      const int G = 128 / np2, MT = (p.walkers + G - 1) / G;
      const int grid = MT < n_sms ? MT : n_sms;
#ifndef DQMC_EMU
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (prof) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
#endif
      // DQMC_TC_TRUNK_TS=0: A operand of the dense GEMMs from shared memory (SS form) instead of tensor memory
      if (trunk_ts) DQ_LAUNCH(tc::trunk_f16_kernel<true>, dim3(grid), dim3(tc::kTrThreads), tc::TrSmem::total(), st, p);
      else DQ_LAUNCH(tc::trunk_f16_kernel<false>, dim3(grid), dim3(tc::kTrThreads), tc::TrSmem::total(), st, p);
#ifndef DQMC_EMU
      if (prof) {
        cudaEventRecord(e1, st);
        // dense layers 12 d^2 and attention 4 N d (scores + weighted sum) flops per row and layer
        prof_note(e0, e1, (double)cfg.n_layers * 2.0 * (double)rows * (6.0 * d * d + 2.0 * N * d), 2);
      }
#endif
      return 0;
    }
#endif
    err = "internal: fused trunk without the tensor-core backend";
    return 5;
  }
  int debug_trunk(const void* X0, void* Out, int rows, cudaStream_t st) override {
    if (!can_trunk(1)) { err = "fused trunk not available for this configuration"; return 2; }
    if (rows % N != 0) { err = "debug_trunk: rows must be a multiple of the electron count"; return 2; }
    int rc = trunk_block((const T*)X0, (T*)Out, rows, st);
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
#ifndef DQMC_EMU
    if (d_trunk_trace) {  // print the stamps relative to the first one
      long long h[64];
      DQ_CHECK(cudaStreamSynchronize(st));
      DQ_CHECK(cudaMemcpy(h, d_trunk_trace, sizeof(h), cudaMemcpyDeviceToHost));
      std::fprintf(stderr, "trunk trace (clocks since the start of layer 1 of the traced tile):");
      for (int i = 0; i < 37; ++i) std::fprintf(stderr, " [%d]%lld", i, h[i] ? h[i] - h[0] : -1LL);
      std::fprintf(stderr, "\n");
    }
#endif
    return 0;
  }

  int debug_gemm(const char* wname, const char* bname, const void* A, const void* Res, void* C, int Mr, int S,
                 int sliced, int backend, cudaStream_t st) override {
    int64_t o = off(wname);
    if (o < 0) { err = "unknown weight"; return 2; }
    int rows = 0, cols = 0;
    for (auto& e : entries) if (e.name == wname) { rows = e.rows; cols = e.cols; }
    int saved = cfg.gemm_backend;
    cfg.gemm_backend = backend;
    int rc = gemm((const T*)A, rows, wname, sliced ? "bf.dn" : nullptr, cfg.n_up, cols, bname ? P(bname) : nullptr,
                  (const T*)Res, cols, (T*)C, cols, Mr, cols, rows, S, sliced, N, st);
    cfg.gemm_backend = saved;
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  // fp32 attention: pick the <N, dh> specialisation (compile-time index arithmetic) when there is one
  template <int NE, int DH>
  int attn_f32_go(const float* QKV, float* O, int Bc, int S, int tb, float scale, int smem, cudaStream_t st, bool setup) {
    if (setup) {
      DQ_CHECK(raise_dyn_smem((attn_fl_f32_kernel<NE, DH, false>), smem));
      DQ_CHECK(raise_dyn_smem((attn_fl_f32_kernel<NE, DH, true>), smem));
      return 0;
    }
    if (attn_fl_mma && S > 1)  // tangent chunks on the tensor cores (8 warp tasks per phase: 256 threads)
      DQ_LAUNCH((attn_fl_f32_kernel<NE, DH, true>), dim3(Bc, H), dim3(256), smem, st, QKV, 3 * d, O, d, N, S, dh, d, scale, tb);
    else
      DQ_LAUNCH((attn_fl_f32_kernel<NE, DH, false>), dim3(Bc, H), dim3(attn_fl_threads), smem, st, QKV, 3 * d, O, d, N, S, dh, d, scale, tb);
    return 0;
  }
  // generic forward-Laplacian attention (any dtype, extra key / value tokens); fp32 with tangents: tensor-core variant
  int launch_attn_generic(const T* QKV, T* O, int Bc, int S, int tb, T scale, const T* kn, const T* vn, cudaStream_t st) {
    if constexpr (std::is_same<T, float>::value) {
      if (attn_gen_mma && S > 1) {
        DQ_LAUNCH((attn_fl_kernel<T, true>), dim3(Bc, H), dim3(256), attn_smem_bytes<T>(N, dh, tb, Mn, 4), st, QKV, 3 * d, O, d, N, S,
                  dh, d, scale, tb, kn, vn, Mn);
        return 0;
      }
    }
    DQ_LAUNCH(attn_fl_kernel<T>, dim3(Bc, H), dim3(S > 1 ? attn_fl_threads : 128), attn_smem_bytes<T>(N, dh, tb, Mn), st, QKV, 3 * d,
              O, d, N, S, dh, d, scale, tb, kn, vn, Mn);
    return 0;
  }
  int launch_attn_f32(const float* QKV, float* O, int Bc, int S, int tb, float scale, int, int smem, cudaStream_t st,
                      bool setup) {
    if (dh == 64) {
      switch (N) {
        case 4: return attn_f32_go<4, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
        case 10: return attn_f32_go<10, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
        case 14: return attn_f32_go<14, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
        case 28: return attn_f32_go<28, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
        case 30: return attn_f32_go<30, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
        default: return attn_f32_go<0, 64>(QKV, O, Bc, S, tb, scale, smem, st, setup);
      }
    }
    return attn_f32_go<0, 0>(QKV, O, Bc, S, tb, scale, smem, st, setup);
  }

  // ---- one chunk of the wave-function pipeline ---------------------------------------------
  // FermiNet trunk (reference: conf/ansatz/ferminet.yaml; gnn/electron_gnn.py:160-259 with
  // Residual / NodeSum / EdgeSum update features and a shared edge MLP): leaves the final electron
  // embeddings in *Xout.
  int ferminet_trunk(const T* r, const T* R, int Rb, int Bc, int S, Ws& w, T** Xout, cudaStream_t st, const T* qa = nullptr) {
    const int rows = Bc * N * S, rowsE = Bc * N * N * S, de = cfg.edge_dim, d0 = 4 * M;
    const T isq2 = (T)0.70710678118654752440;
    DQ_LAUNCH(embed_kernel<T>, dim3(Bc * N), dim3(128), sizeof(T) * 5 * d0, st, r, R, Rb, N, M, cfg.n_up, S, 0, 0,
              (const T*)nullptr, d0, w.X, Bc * N, 1, qa);
    DQ_LAUNCH(edge_feat_kernel<T>, dim3((Bc * N * N + 127) / 128), dim3(128), 0, st, r, N, S, w.A, Bc * N * N, qa);
    T* Hc = w.X; T* Hn = w.O; T* Ec = w.A; T* En = w.M1;
    int dcur = d0, ecur = 4;
    for (int l = 0; l < cfg.n_layers; ++l) {
      std::string p = "F" + std::to_string(l) + ".";
      const int fin = 3 * dcur + 2 * ecur;
      DQ_LAUNCH(fermi_agg_kernel<T>, dim3(Bc * S, N), dim3(128), 0, st, (const T*)Hc, dcur, (const T*)Ec, ecur, N, cfg.n_up,
                S, w.QKV);
      int rc = gemm(w.QKV, fin, (p + "wg").c_str(), nullptr, 0, d, P(p + "bg"), nullptr, 0, Hn, d, rows, d, fin, S, 0, N, st);
      if (rc) return rc;
      DQ_LAUNCH(tanh_fl_kernel<T>, dim3(Bc * N, (d + 127) / 128), dim3(128), 0, st, Hn, d,
                (const T*)(dcur == d ? Hc : nullptr), dcur, S, d, dcur == d ? isq2 : T(1));
      if (l < cfg.n_layers - 1) {
        rc = gemm(Ec, ecur, (p + "wu").c_str(), nullptr, 0, de, P(p + "bu"), nullptr, 0, En, de, rowsE, de, ecur, S, 0, N, st);
        if (rc) return rc;
        DQ_LAUNCH(tanh_fl_kernel<T>, dim3(Bc * N * N, 1), dim3(32), 0, st, En, de, (const T*)(ecur == de ? Ec : nullptr),
                  ecur, S, de, ecur == de ? isq2 : T(1));
        T* t2 = Ec; Ec = En; En = t2;
        ecur = de;
      }
      T* t1 = Hc; Hc = Hn; Hn = t1;
      dcur = d;
    }
    *Xout = Hc;
    return 0;
  }

  // conv-GNN trunk (reference tests/conf/ansatz.yaml): embedding lookup, per layer edge filters w_t, node
  // transforms h_t, convolution over same / anti / ne edges, featurewise update sum_t g_t(conv_t) + residual;
  // then the Jastrow MLP on sum_i x_i and the hidden layers of the per-spin backflow MLPs (ssp).
  // MLP of `nl` Linear(+bias)+tanh layers on augmented rows (groups of Sg slots): in -> out, ping-pong through tmp
  int gnn_mlp(const T* in, int din, const std::string& base, const int* dims, int nl, bool bias, T* tmp, T* out, int rows_,
              int Sg, cudaStream_t st) {
    const T* cur = in;
    int dc = din;
    for (int i = 0; i < nl; ++i) {
      T* dst = (i == nl - 1) ? out : tmp;
      const std::string q = base + "." + std::to_string(i);
      int rc = gemm(cur, dc, (q + ".w").c_str(), nullptr, 0, dims[i], bias ? P(q + ".b") : nullptr, nullptr, 0, dst, dims[i], rows_,
                    dims[i], dc, Sg, 0, 1, st);
      if (rc) return rc;
      DQ_LAUNCH(act_fl_kernel<T>, dim3(rows_ / Sg, (dims[i] + 63) / 64), dim3(64), 0, st, dst, dims[i], (const T*)nullptr, 0, Sg,
                dims[i], T(1), 0);
      cur = dst; dc = dims[i];
    }
    return 0;
  }

  int paulinet_trunk(const T* r, const T* R, int Rb, int Bc, int S, Ws& w, T** Xbf, const T** jastrow, cudaStream_t st,
                     const T* qa = nullptr) {
    const int rows = Bc * N * S, e = cfg.edge_dim, groups = Bc * N, nl = cfg.gnn_sub_n > 0 ? cfg.gnn_sub_n : 1;
    const int Mne = cfg.gnn_conv_ne ? M : 0, NS = N + Mne, nt = cfg.gnn_conv_ne ? 3 : 2;
    const int pairs = Bc * N * NS, prow = pairs * 8;  // compact edge rows: 8 slots per (receiver, sender) pair
    const T isq2 = (T)0.70710678118654752440;
    int dcur = d, ecur = 4;
    if (cfg.gnn_features) {
      dcur = 4 * M;
      DQ_LAUNCH(embed_kernel<T>, dim3(Bc * N), dim3(128), sizeof(T) * 5 * dcur, st, r, R, Rb, N, M, cfg.n_up, S, 0, 0,
                (const T*)nullptr, dcur, w.X, Bc * N, 1, qa);
    } else {
      DQ_LAUNCH(gnn_embed_kernel<T>, dim3((Bc * N * d + 127) / 128), dim3(128), 0, st, P("emb.table"),
                cfg.n_elec_types > 0 ? cfg.n_elec_types : 1, N, cfg.n_up, S, d, w.X, Bc * N);
    }
    DQ_LAUNCH(gnn_edge_feat_kernel<T>, dim3((pairs + 63) / 64), dim3(64), 0, st, r, R, Rb, N, M, Mne, w.E0, pairs, S > 1 ? qa : (const T*)nullptr);
    T* X = w.X;
    T* Xn = w.O;
    T* E = w.E0;
    T* En = w.E1;
    const char* tn[3] = {"same", "anti", "ne"};
    for (int l = 0; l < cfg.n_layers; ++l) {
      const std::string p = "G" + std::to_string(l) + ".";
      // edge filters of every type on all pairs (compact rows), node transforms h_same / h_anti
      for (int t = 0; t < nt; ++t) {
        int rc = gnn_mlp(E, ecur, p + "w_" + tn[t], cfg.gnn_w_dims[l], nl, cfg.gnn_w_bias != 0, w.ET0, w.W3 + (size_t)t * prow * e,
                         prow, 8, st);
        if (rc) return rc;
      }
      int rc = gnn_mlp(X, dcur, p + "h_same", cfg.gnn_h_dims[l], nl, true, w.HT, w.Hs, rows, S, st);
      if (rc) return rc;
      rc = gnn_mlp(X, dcur, p + "h_anti", cfg.gnn_h_dims[l], nl, true, w.HT, w.Ha, rows, S, st);
      if (rc) return rc;
      DQ_LAUNCH(gnn_conv_kernel<T>, dim3(groups), dim3(128), 0, st, (const T*)w.W3, (const T*)(w.W3 + (size_t)prow * e),
                (const T*)(w.W3 + (size_t)2 * prow * e), (const T*)w.Hs, (const T*)w.Ha,
                cfg.gnn_conv_ne ? P(p + "hne") : (const T*)nullptr, N, Mne, cfg.n_up, S, e, w.C);
      T* Xout;
      if (cfg.gnn_concat) {
        // x <- [(x +) tanh(g([x, mean_up x, mean_down x, conv_*]))] (/ sqrt 2)   (electron_gnn.py:243-259)
        const int fin = 3 * dcur + nt * e;
        DQ_LAUNCH(gnn_concat_kernel<T>, dim3(Bc * S, N), dim3(128), 0, st, (const T*)X, dcur, (const T*)w.C, nt * e, N, cfg.n_up,
                  S, w.Fc);
        rc = gemm(w.Fc, fin, (p + "g.w").c_str(), nullptr, 0, d, cfg.gnn_g_bias ? P(p + "g.b") : nullptr, nullptr, 0, Xn, d, rows,
                  d, fin, S, 0, N, st);
        if (rc) return rc;
        const bool res = dcur == d;
        DQ_LAUNCH(act_fl_kernel<T>, dim3(groups, (d + 63) / 64), dim3(64), 0, st, Xn, d, res ? (const T*)X : (const T*)nullptr,
                  d, S, d, (res && cfg.gnn_res_norm) ? isq2 : T(1), 0);
        Xout = Xn; Xn = X;
      } else {
        // featurewise update: x <- x + sum_t tanh(g_t(conv_t))   (residual hkext.py:116-137)
        T* G[3] = {w.G0, w.G1, w.G2};
        const T* res = dcur == d ? X : nullptr;
        for (int t = 0; t < nt; ++t) {
          rc = gemm(w.C + t * e, nt * e, (p + "g_" + tn[t] + ".w").c_str(), nullptr, 0, d, P(p + "g_" + tn[t] + ".b"), nullptr,
                    0, G[t], d, rows, d, e, S, 0, N, st);
          if (rc) return rc;
          const bool last = t == nt - 1;
          DQ_LAUNCH(act_fl_kernel<T>, dim3(groups, (d + 63) / 64), dim3(64), 0, st, G[t], d, res, d, S, d,
                    (last && res && cfg.gnn_res_norm) ? isq2 : T(1), 0);
          res = G[t];
        }
        Xout = G[nt - 1];
        // the accumulated buffer becomes the new x; hand the old x buffer to the G slot
        T* old = X;
        if (nt == 3) w.G2 = old; else w.G1 = old;
      }
      X = Xout;
      dcur = d;
      if (cfg.gnn_deep_edges && l < cfg.n_layers - 1) {
        // shared edge MLP u on the compact rows + normalised residual (electron_gnn.py:160-192)
        rc = gnn_mlp(E, ecur, p + "u", cfg.gnn_u_dims[l], nl, true, w.ET0, w.ET1, prow, 8, st);
        if (rc) return rc;
        if (ecur == e) {
          // (e + u(e)) / sqrt 2: u's last tanh already applied -> plain scaled add on all slots
          DQ_LAUNCH((axpby_kernel<T>), dim3((unsigned)(((size_t)prow * e + 255) / 256)), dim3(256), 0, st, (const T*)E,
                    (const T*)w.ET1, isq2, En, (size_t)prow * e);
          T* tmp = E; E = En; En = tmp;
        } else {
          T* tmp = E; E = w.ET1; w.ET1 = tmp;
        }
        ecur = e;
      }
    }
    *jastrow = nullptr;
    if (cfg.jastrow_n > 0) {
      const int jr = Bc * S;
      T* cur = w.Jb;
      int din = d;
      DQ_LAUNCH(sum_electrons_kernel<T>, dim3((jr * d + 127) / 128), dim3(128), 0, st, (const T*)X, N, S, d, cur, jr * d);
      for (int i = 0; i < cfg.jastrow_n; ++i) {
        const int dout = cfg.jastrow_dims[i];
        T* nxt = cur + (size_t)jr * din;
        const bool last = i == cfg.jastrow_n - 1;
        const std::string q = "J" + std::to_string(i);
        int rc = gemm(cur, din, (q + ".w").c_str(), nullptr, 0, dout, last ? nullptr : P(q + ".b"), nullptr, 0, nxt, dout, jr,
                      dout, din, S, 0, 1, st);
        if (rc) return rc;
        if (!last) DQ_LAUNCH(act_fl_kernel<T>, dim3(Bc, (dout + 31) / 32), dim3(32), 0, st, nxt, dout, (const T*)nullptr, 0, S, dout, T(1), 1);
        cur = nxt; din = dout;
      }
      *jastrow = cur;  // [Bc][S] scalar rows
    }
    T* Y = X;
    int din = d;
    for (int i = 0; i < cfg.backflow_n; ++i) {
      const int dout = cfg.backflow_dims[i];
      const std::string q = std::to_string(i);
      T* out = (i & 1) ? w.Y1 : w.Y0;
      int rc = gemm(Y, din, ("bfh" + q + ".up").c_str(), ("bfh" + q + ".dn").c_str(), cfg.n_up, dout, P("bfb" + q + ".up"),
                    nullptr, 0, out, dout, Bc * S, dout, din, S, 1, N, st, 0, P("bfb" + q + ".dn"));
      if (rc) return rc;
      DQ_LAUNCH(act_fl_kernel<T>, dim3(groups, (dout + 31) / 32), dim3(32), 0, st, out, dout, (const T*)nullptr, 0, S, dout, T(1), 1);
      Y = out; din = dout;
    }
    *Xbf = Y;
    return 0;
  }

  int run_chunk(const T* r, const T* R, int Rb, int Bc, int S, int Bstat, T* sign, T* logp, T* E, T* stats, T* grad,
                void* wsbase, cudaStream_t st) {
    Ws w = carve(wsbase, Bc, S);
    const int rows = Bc * N * S;
    const T* qa = nullptr;  // pseudo-Hamiltonian: per-electron metric of the forward-Laplacian pass
    if (ph_on && ph_active && S > 1) {
      DQ_LAUNCH(ph_coeff_kernel<T>, dim3((Bc * N + 127) / 128), dim3(128), 0, st, r, R, Rb, N, M, ph_args(nullptr), w.QA,
                Bc * N);
      qa = w.QA;
    }
    if (gnn) {
      T* Xbf = nullptr;
      const T* jas = nullptr;
      int rc = paulinet_trunk(r, R, Rb, Bc, S, w, &Xbf, &jas, st, qa);
      if (rc) return rc;
      return tail(r, R, Rb, Bc, S, Bstat, sign, logp, E, stats, grad, w, Xbf, st, jas, qa);
    }
    if (cfg.kind == DQMC_FERMINET) {
      T* Xf = nullptr;
      int rc = ferminet_trunk(r, R, Rb, Bc, S, w, &Xf, st, qa);
      if (rc) return rc;
      return tail(r, R, Rb, Bc, S, Bstat, sign, logp, E, stats, grad, w, Xf, st, nullptr, qa);
    }
    const int F = 4 * M + 1;
    const bool compact = S == 1 && embed_fwd_ok && ecp_emb && can_trunk(S);
    if (compact) {
      // quadrature forwards of the non-local ECP: only the moved electron's embedding row is new, the whole-trunk kernel
      // takes the other rows from the base walkers' table
      int epb = (Bc / (2 * n_sms)) / 32 * 32;
      epb = epb < 32 ? 32 : (epb > 512 ? 512 : epb);
      DQ_LAUNCH(embed_fwd_kernel<T>, dim3((Bc + epb - 1) / epb), dim3(256), embed_fwd_smem_bytes<T>(M, d), st, r, R, Rb, N,
                M, cfg.n_up, 1, P("emb.w"), d, w.X, Bc, epb, (long long)ecp_v0, ecp_vper);
    } else if (S == 1 && embed_fwd_ok) {
      // plain forwards (Metropolis, ECP quadrature): register-tiled projection, W staged per block
      const int tot = Bc * N;
      int epb = (tot / (2 * n_sms)) / 32 * 32;
      epb = epb < 32 ? 32 : (epb > 512 ? 512 : epb);
      DQ_LAUNCH(embed_fwd_kernel<T>, dim3((tot + epb - 1) / epb), dim3(256), embed_fwd_smem_bytes<T>(M, d), st, r, R, Rb, N,
                M, cfg.n_up, 1, P("emb.w"), d, w.X, tot, epb, 0LL, 0);
    } else {
      const int epb = S == 1 ? 8 : 1;  // plain forwards: several electrons per block (tiny per-electron work)
      DQ_LAUNCH(embed_kernel<T>, dim3((Bc * N + epb - 1) / epb), dim3(128), sizeof(T) * 5 * F, st, r, R, Rb, N, M,
                cfg.n_up, S, 1, 1, P("emb.w"), d, w.X, Bc * N, epb, qa);
    }
    T* X = w.X;
    T* O = w.O;
    const T scale = (T)(1.0 / std::sqrt((double)dh));
    if (can_trunk(S)) {  // plain forward: every layer in one persistent tensor-core launch
      int rc = trunk_block(X, O, rows, st, compact ? ecp_emb : nullptr);
      if (rc) return rc;
      return tail(r, R, Rb, Bc, S, Bstat, sign, logp, E, stats, grad, w, O, st, nullptr, qa);
    }
    for (int l = 0; l < cfg.n_layers; ++l) {
      std::string p = "L" + std::to_string(l) + ".";
      gemm(X, d, (p + "wqkv").c_str(), nullptr, 0, 3 * d, nullptr, nullptr, 0, w.QKV, 3 * d, rows, 3 * d, d, S, 0, N, st);
      {
        const int tb = S > 1 ? attn_tb : 1;
        const T* kn = Mn > 0 ? P(p + "kn") : nullptr;
        const T* vn = Mn > 0 ? P(p + "vn") : nullptr;
        if constexpr (std::is_same<T, float>::value) {
          if (S == 1 && attn_mma_ok) {
            // tensor-core attention: a warp per (walker, head, 16-query tile), fragments straight from global memory
            const int n_pairs = Bc * H, tasks = n_pairs * ((N + 15) / 16);
            int nblk = (tasks + 3) / 4;
            if (nblk > n_sms * 8) nblk = n_sms * 8;
            const dim3 grid(nblk), block(128);
#define DQ_ATTN_MMA(NK_)                                                                                                     \
  DQ_LAUNCH(attn_fwd_mma_kernel<NK_>, grid, block, 0, st, (const float*)w.QKV, 3 * d, (float*)O, d, N, H, d, (float)scale, \
            n_pairs, (const float*)kn, (const float*)vn, Mn)
            switch ((N + Mn + 7) / 8) {
              case 1: DQ_ATTN_MMA(1); break;
              case 2: DQ_ATTN_MMA(2); break;
              case 3: DQ_ATTN_MMA(3); break;
              case 4: DQ_ATTN_MMA(4); break;
              case 5: DQ_ATTN_MMA(5); break;
              default: DQ_ATTN_MMA(6); break;
            }
#undef DQ_ATTN_MMA
          } else if (S == 1 && attn_fwd_ok && attn_fwd_pipelined) {
            // persistent blocks (one per SM), 6 warps each, K / V of the next pair prefetched by cp.async
            const int n_pairs = Bc * H, wpb = 6;
            const int smem = wpb * 4 * N * 64 * (int)sizeof(float);
            const int nblk = (n_pairs + wpb - 1) / wpb;
            const dim3 grid(nblk < n_sms ? nblk : n_sms), block(32 * wpb);
            if (N <= 8)
              DQ_LAUNCH(attn_fwd2_f32_kernel<8>, grid, block, smem, st, (const float*)w.QKV, 3 * d, (float*)O, d, N, H, d,
                        (float)scale, n_pairs);
            else if (N <= 16)
              DQ_LAUNCH(attn_fwd2_f32_kernel<16>, grid, block, smem, st, (const float*)w.QKV, 3 * d, (float*)O, d, N, H, d,
                        (float)scale, n_pairs);
            else
              DQ_LAUNCH(attn_fwd2_f32_kernel<32>, grid, block, smem, st, (const float*)w.QKV, 3 * d, (float*)O, d, N, H, d,
                        (float)scale, n_pairs);
          } else if (S == 1 && attn_fwd_ok) {
            const int n_pairs = Bc * H;
            const int smem = 4 * 2 * (N + Mn) * 64 * (int)sizeof(float);
            const dim3 grid((n_pairs + 3) / 4), block(128);
#define DQ_ATTN_FWD(NM_, EX_)                                                                                          \
  DQ_LAUNCH((attn_fwd_f32_kernel<NM_, EX_>), grid, block, smem, st, (const float*)w.QKV, 3 * d, (float*)O, d, N, H, d, \
            (float)scale, n_pairs, (const float*)kn, (const float*)vn, Mn)
            switch (Mn > 0 ? -1 : N) {  // exact-size instances for the benchmark molecules, padded generic ones otherwise
              case 4: DQ_ATTN_FWD(4, true); break;
              case 10: DQ_ATTN_FWD(10, true); break;
              case 14: DQ_ATTN_FWD(14, true); break;
              case 28: DQ_ATTN_FWD(28, true); break;
              case 30: DQ_ATTN_FWD(30, true); break;
              default:
                if (N + Mn <= 8) DQ_ATTN_FWD(8, false);
                else if (N + Mn <= 16) DQ_ATTN_FWD(16, false);
                else if (N + Mn <= 32) DQ_ATTN_FWD(32, false);
                else DQ_ATTN_FWD(48, false);
            }
#undef DQ_ATTN_FWD
          } else if (attn_f32) {
            if (launch_attn_f32((const float*)w.QKV, (float*)O, Bc, S, tb, (float)scale, 0,
                                (int)attn_f32_smem_bytes(N, dh, tb), st, false))
              return 1;
          } else {
            int rc = launch_attn_generic((const T*)w.QKV, O, Bc, S, tb, scale, kn, vn, st);
            if (rc) return rc;
          }
        } else {
          int rc = launch_attn_generic((const T*)w.QKV, O, Bc, S, tb, scale, kn, vn, st);
          if (rc) return rc;
        }
      }
      if (can_fuse_mlp(S, p)) {
        // plain forward: attention projection + residual and both MLP layers in ONE launch, result in place of O
        int rc = mlp_block(p, O, X, O, rows, st);
        if (rc) return rc;
        T* tmp = X; X = O; O = tmp;
        continue;
      }
      gemm(O, d, (p + "wo").c_str(), nullptr, 0, d, nullptr, X, d, w.A, d, rows, d, d, S, 0, N, st);
      if (can_fuse_act(S)) {
        // MLP with the tanh (and its Jacobian/Laplacian propagation) inside the GEMM epilogues
        gemm(w.A, d, (p + "w1").c_str(), nullptr, 0, d, P(p + "b1"), nullptr, 0, w.M1, d, rows, d, d, S, 0, N, st, 1);
        gemm(w.M1, d, (p + "w2").c_str(), nullptr, 0, d, P(p + "b2"), w.A, d, O, d, rows, d, d, S, 0, N, st, 1);
      } else {
        gemm(w.A, d, (p + "w1").c_str(), nullptr, 0, d, P(p + "b1"), nullptr, 0, w.M1, d, rows, d, d, S, 0, N, st);
        DQ_LAUNCH(tanh_fl_kernel<T>, dim3(Bc * N, (d + 127) / 128), dim3(128), 0, st, w.M1, d, (const T*)nullptr, 0, S, d,
                  T(1));
        gemm(w.M1, d, (p + "w2").c_str(), nullptr, 0, d, P(p + "b2"), nullptr, 0, O, d, rows, d, d, S, 0, N, st);
        DQ_LAUNCH(tanh_fl_kernel<T>, dim3(Bc * N, (d + 127) / 128), dim3(128), 0, st, O, d, (const T*)w.A, d, S, d, T(1));
      }
      T* tmp = X; X = O; O = tmp;
    }
    return tail(r, R, Rb, Bc, S, Bstat, sign, logp, E, stats, grad, w, X, st, nullptr, qa);
  }

  // backflow heads -> Slater determinants -> det sum / cusp / potentials (shared by all trunks)
  int tail(const T* r, const T* R, int Rb, int Bc, int S, int Bstat, T* sign, T* logp, T* E, T* stats, T* grad, Ws& w,
           T* X, cudaStream_t st, const T* jastrow = nullptr, const T* qa = nullptr) {
    // per-spin backflow heads: rows of electron e across walkers, weights by spin
    gemm(X, bf_in, "bf.up", "bf.dn", cfg.n_up, BFW, gnn ? P("bfb.up") : nullptr, nullptr, 0, w.BF, BFW, Bc * S, BFW, bf_in, S, 1, N,
         st, 0, gnn ? P("bfb.dn") : nullptr);
    if (cfg.mult_act == 1)  // default mult_act 1 + 2 tanh(x / 4) of the BackflowOp (nn_wave_function.py:14-33)
      DQ_LAUNCH(act_fl_kernel<T>, dim3(Bc * N, (KN + 127) / 128), dim3(128), 0, st, w.BF, KN, (const T*)nullptr, 0, S, KN, T(1), 2);
    const int full_det = cfg.factorized_det ? 0 : 1;
    const int mult_on = cfg.backflow_add == 1 ? 0 : 1;
    const T* gadd = nullptr;
    if (cfg.backflow_add) {
      // additive branch (wf/nn_wave_function.py:26-32): add_act on its head, electron-local factor cutoff * |envelope|
      if (qa) { err = "additive backflow with a pseudo-Hamiltonian is not supported"; return 2; }
      DQ_LAUNCH(act_fl_kernel<T>, dim3(Bc * N, (KN + 127) / 128), dim3(128), 0, st, w.BF + add_off, BFW, (const T*)nullptr, 0, S, KN,
                T(1), 3);
      DQ_LAUNCH(bf_add_factor_kernel<T>, dim3((Bc * N + 63) / 64), dim3(64), 0, st, r, R, Rb, N, M, cfg.n_up, K, P("env.pi_up"),
                P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), env_rep, full_det, w.Gadd, Bc * N);
      gadd = w.Gadd;
    }
    if (mos_out) {  // Ansatz.apply(..., return_mos=True): orbital matrices instead of determinants
      const size_t tot = (size_t)Bc * K * N * N;
      DQ_LAUNCH(orbitals_kernel<T>, dim3((unsigned)((tot + 255) / 256)), dim3(256), 0, st, r, R, Rb, N, M, cfg.n_up, K,
                P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)w.BF, BFW, env_rep, full_det,
                mos_out, tot, gadd, add_off, mult_on);
      return 0;
    }
    const int sl_wpb = slater_warps_per_block<T>(N);
    if ((N <= 4 || (N <= 6 && std::is_same<T, float>::value)) && !qa && !gadd && !std::getenv("DQMC_SLATER_GENERIC")) {
      const int tot = Bc * K;
#define DQ_SL_SMALL(NS_)                                                                                           \
  DQ_LAUNCH((slater_small_kernel<T, NS_>), dim3((tot + 63) / 64), dim3(64), 0, st, r, R, Rb, M, cfg.n_up, K, S, tot,    \
            P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)w.BF, KN, w.dsign, w.dlog, \
            w.dgrad, w.dlap, env_rep, full_det)
      switch (N) {
        case 2: DQ_SL_SMALL(2); break;
        case 3: DQ_SL_SMALL(3); break;
        case 4: DQ_SL_SMALL(4); break;
        case 5: DQ_SL_SMALL(5); break;
        default: DQ_SL_SMALL(6); break;
      }
#undef DQ_SL_SMALL
    } else if (S == 1 && N <= 32 && slater_fwd2_ok && !gadd) {
      const int nthr = 32 * K < 256 ? 32 * K : 256;
      const int grid = Bc < 3 * n_sms ? Bc : 3 * n_sms;
      // register-resident LU: the row array is sized to the electron count where an exact instance exists (every padded
      // column costs a shuffle + FMA per elimination step)
#define DQ_SL_FWD2(NMV)                                                                                                        \
  DQ_LAUNCH((slater_fwd2_kernel<T, NMV>), dim3(grid), dim3(nthr), slater_fwd2_smem_bytes<T>(N, M, K), st, r, R, Rb, N, M,      \
            cfg.n_up, K, Bc, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)w.BF, KN, w.dsign, \
            w.dlog, env_rep, full_det, ecp_env, (long long)ecp_v0, ecp_vper)
      if (N == 14) DQ_SL_FWD2(14);
      else if (N <= 16) DQ_SL_FWD2(16);
      else if (N == 28) DQ_SL_FWD2(28);
      else if (N == 30) DQ_SL_FWD2(30);
      else DQ_SL_FWD2(32);
#undef DQ_SL_FWD2
    } else if (S == 1 && N <= 32 && !gadd && !std::getenv("DQMC_SLATER_GENERIC")) {
      const int wpb = K < 8 ? K : 8;
      DQ_LAUNCH(slater_fwd_reg_kernel<T>, dim3(Bc), dim3(32 * wpb), sizeof(T) * N * M, st, r, R, Rb, N, M, cfg.n_up, K,
                P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)w.BF, KN, w.dsign, w.dlog,
                env_rep, full_det);
    } else
    DQ_LAUNCH(slater_kernel<T>, dim3((Bc * K + sl_wpb - 1) / sl_wpb), dim3(32 * sl_wpb), slater_smem_bytes<T>(N), st, r, R,
              Rb, N, M, cfg.n_up, K, S, Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)w.BF, BFW, w.dsign, w.dlog,
              w.dgrad, w.dlap, env_rep, full_det, qa, gadd, add_off, mult_on);
    FinalizeCfg fc;
    fc.N = N; fc.M = M; fc.n_up = cfg.n_up; fc.K = K; fc.S = S; fc.cusp_kind = cfg.cusp_kind;
    fc.cusp_same_scale = cfg.cusp_same_scale; fc.cusp_anti_scale = cfg.cusp_anti_scale;
    fc.ecp_terms = cfg.ecp_loc_terms;
    fc.nuc_cusp_kind = cfg.nuc_cusp_kind;
    DQ_LAUNCH(finalize_kernel<T>, dim3(Bc), dim3(128), finalize_smem_bytes<T>(N, K), st, fc, r, R, Rb,
              (const T*)w.dsign, (const T*)w.dlog, (const T*)w.dgrad, (const T*)w.dlap, P("cusp.alpha"),
              (const T*)d_zval, (const T*)d_ecp_loc, (const int*)d_ecp_mask, Bstat, sign, logp, E, stats, grad,
              cfg.conf_linear ? P("conf.w") : (const T*)nullptr, jastrow,
              cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr, qa ? ph_args(qa) : PhArgs<T>());
    return 0;
  }

  int run_batched(const T* r, const T* R, int Rb, int B, int S, T* sign, T* logp, T* E, T* stats, T* grad, void* ws,
                  int64_t wsb, cudaStream_t st) {
    int Bc = max_chunk(wsb, S, B);
    if (Bc < 1) { err = "workspace too small for a single walker"; return 3; }
    if ((int64_t)chunk_bytes(Bc, S) > wsb) { err = "internal: carved workspace exceeds the planned size"; return 3; }
    if (dry) { carve(ws, Bc, S); return 0; }  // planning pass: record the extent of the largest chunk
    for (int b0 = 0; b0 < B; b0 += Bc) {
      int nb = std::min(Bc, B - b0);
      ecp_v0 = b0;  // index of the chunk's first walker in the caller's batch (quadrature forwards: virtual-walker index)
      int rc = run_chunk(r + (size_t)b0 * 3 * N, R + (Rb ? (size_t)b0 * 3 * M : 0), Rb, nb, S, B, sign + b0, logp + b0,
                         E ? E + b0 : nullptr, stats ? stats + b0 : nullptr, grad ? grad + (size_t)b0 * T3 : nullptr, ws,
                         st);
      if (rc) return rc;
      rc = check_guards();
      if (rc) return rc;
    }
    return 0;
  }

  // ---- Metropolis-adjusted Langevin sweep: value + drift from the forward-Laplacian pass ----
  // state = {r, sign, log, force}; force == clean_force(grad log|psi|) with the CURRENT tau (electron_samplers.py:197-211)
  int value_and_force(const T* r, const T* R, int Rb, int B, const T* tau, T* sign, T* logp, T* force, char* p, int64_t rest,
                      cudaStream_t st) {
    T* E = (T*)p; p += align_up(sizeof(T) * (size_t)B);
    T* stt = (T*)p; p += align_up(sizeof(T) * (size_t)6 * B);
    T* grad = (T*)p; p += align_up(sizeof(T) * (size_t)B * T3);
    rest -= force_prefix_bytes(B);
    note_hwm(p);
    if (rest < 0) { err = "workspace too small (Langevin force buffers)"; return 3; }
    int rc = run_batched(r, R, Rb, B, T3 + 2, sign, logp, E, stt, grad, p, rest, st);
    if (rc) return rc;
    DQ_LAUNCH(langevin_force_kernel<T>, dim3((B * N + 127) / 128), dim3(128), 0, st, (const T*)grad, r, R, Rb, (const T*)d_znuc, tau,
              N, M, B * N, force);
    return 0;
  }
  int langevin(void* r_, void* sign_, void* logp_, void* force_, int32_t* age, void* tau_, const void* R_, int Rb, int B, int n_sub,
               double target, int max_age, uint64_t seed, uint64_t step0, uint64_t woff, const void* nn, const void* nu,
               void* stats_, void* ws, int64_t wsb, cudaStream_t st) override {
    T* r = (T*)r_; T* sign = (T*)sign_; T* logp = (T*)logp_; T* force = (T*)force_; T* tau = (T*)tau_; T* stats = (T*)stats_;
    const T* R = (const T*)R_;
    if (n_sub == 0) return langevin_update(r, R, Rb, B, tau, sign, logp, force, ws, wsb, st);
    char* p = (char*)ws;
    T* rp = (T*)p; p += align_up(sizeof(T) * (size_t)B * 3 * N);
    T* fp = (T*)p; p += align_up(sizeof(T) * (size_t)B * 3 * N);
    T* sp = (T*)p; p += align_up(sizeof(T) * (size_t)B);
    T* lp = (T*)p; p += align_up(sizeof(T) * (size_t)B);
    int* cnt = (int*)p; p += 256;
    int64_t rest = wsb - (p - (char*)ws);
    note_hwm(p);
    if (rest < 0) { err = "workspace too small (Langevin proposal buffers)"; return 3; }
    DQ_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), st));
    const int ne = B * 3 * N;
    for (int s = 0; s < n_sub; ++s) {
      const T* nns = nn ? (const T*)nn + (size_t)s * ne : nullptr;
      const T* nus = nu ? (const T*)nu + (size_t)s * B : nullptr;
      DQ_LAUNCH(langevin_propose_kernel<T>, dim3((ne / 2 + 1 + 127) / 128), dim3(128), 0, st, (const T*)r, (const T*)force, rp,
                (const T*)tau, nns, seed, step0 + (uint64_t)s, woff * (uint64_t)(3 * N), ne);
      int rc = value_and_force(rp, R, Rb, B, tau, sp, lp, fp, p, rest, st);
      if (rc) return rc;
      DQ_LAUNCH(langevin_accept_kernel<T>, dim3((B + 127) / 128), dim3(128), 0, st, r, (const T*)rp, force, (const T*)fp, sign,
                (const T*)sp, logp, (const T*)lp, age, (const T*)tau, nus, seed, step0 + (uint64_t)s, woff, max_age, B, N, cnt);
      DQ_LAUNCH(tau_kernel<T>, dim3(1), dim3(32), 0, st, tau, cnt, B, (T)target, stats);
    }
    DQ_LAUNCH(sampler_stats_kernel<T>, dim3(1), dim3(256), 0, st, (const T*)r, (const T*)logp, (const int*)age, (const T*)tau, B,
              N, stats);
    DQ_CHECK(cudaGetLastError());
    return 0;
  }
  // n_sub == 0 with force output: (re)compute psi and the drift of the current walkers (sampler.update)
  int langevin_update(const T* r, const T* R, int Rb, int B, const T* tau, T* sign, T* logp, T* force, void* ws, int64_t wsb,
                      cudaStream_t st) {
    int rc = value_and_force(r, R, Rb, B, tau, sign, logp, force, (char*)ws, wsb, st);
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  // ---- parameter VJP of the plain forward (Psiformer): SURVEY.md 8(f) N1 ------------------------
  const T* PT(const std::string& n) const { return d_params_t + off(n); }
  // C = (Res) + A @ W with a raw weight pointer (CUDA-core kernel; used by the reverse pass with transposed weights)
  int gemm_raw(const T* A, int lda, const T* W0, const T* W1, int zsplit, int ldw, const T* Res, int ldr, T* C, int ldc,
               int Mr, int Nc, int Kc, int sliced, cudaStream_t st) {
    GemmArgs<T> g;
    g.A = A; g.lda = lda; g.W0 = W0; g.W1 = W1; g.z_split = zsplit; g.ldw = ldw; g.bias = nullptr; g.Res = Res;
    g.ldr = ldr; g.C = C; g.ldc = ldc; g.M = Mr; g.N = Nc; g.K = Kc; g.S = 1; g.sliced = sliced; g.Nel = N;
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((Nc + BN - 1) / BN, (Mr + BM - 1) / BM, sliced ? N : 1);
    DQ_LAUNCH((gemm_kernel<T, BM, BN, BK, TM, TN>), grid, dim3(256), 0, st, g);
    return 0;
  }
  // dW += A^T dY over `rows` rows (optionally only electrons lo <= i < hi of every walker), db += column sums
  void wgrad(const T* A, int lda, const T* dY, int ldy, int rows, int Kc, int Nc, T* dW, int lo, int hi, cudaStream_t st) {
    int nz = (rows + 4095) / 4096;
    if (nz > 256) nz = 256;
    if (nz < 1) nz = 1;
    const int rpb = ((rows + nz - 1) / nz + 31) / 32 * 32;
    DQ_LAUNCH(gemm_tn_kernel<T>, dim3((Nc + 31) / 32, (Kc + 31) / 32, (rows + rpb - 1) / rpb), dim3(256), 0, st, A, lda, dY, ldy,
              rows, Kc, Nc, rpb, hi > lo ? N : 0, lo, hi, dW, Nc);
  }
  void bgrad(const T* dZ, int ld, int rows, int Nc, T* db, cudaStream_t st, int lo = 0, int hi = 0) {
    int ny = (rows + 2047) / 2048;
    if (ny > 128) ny = 128;
    if (ny < 1) ny = 1;
    const int rpb = (rows + ny - 1) / ny;
    DQ_LAUNCH(colsum_kernel<T>, dim3((Nc + 127) / 128, (rows + rpb - 1) / rpb), dim3(128), 0, st, dZ, ld, rows, Nc, rpb, db,
              hi > lo ? N : 0, lo, hi);
  }

  int vjp_chunk(const T* r, const T* R, int Rb, int Bc, const T* wts, T* sign, T* logp, T* G, void* wsbase, cudaStream_t st) {
    const int L = cfg.n_layers, rows = Bc * N, F = 4 * M + 1;
    char* p = (char*)wsbase;
    auto take = [&](size_t n) { T* q = (T*)p; p += align_up(sizeof(T) * n); DQ_TAKE_GUARD(); return q; };
    std::vector<T*> X(L + 1), QKV(L), O(L), A(L), M1(L);
    for (int l = 0; l <= L; ++l) X[l] = take((size_t)rows * d);
    for (int l = 0; l < L; ++l) { QKV[l] = take((size_t)rows * 3 * d); O[l] = take((size_t)rows * d); A[l] = take((size_t)rows * d); M1[l] = take((size_t)rows * d); }
    T* BF = take((size_t)rows * KN); T* dBF = take((size_t)rows * KN);
    T* dsign = take((size_t)Bc * K); T* dlog = take((size_t)Bc * K); T* dld = take((size_t)Bc * K);
    T* dXn = take((size_t)rows * d); T* dZ = take((size_t)rows * d); T* dM1 = take((size_t)rows * d);
    T* dA = take((size_t)rows * d); T* dO = take((size_t)rows * d); T* dQKV = take((size_t)rows * 3 * d);
    T* dX = take((size_t)rows * d); T* Feat = take((size_t)rows * F);
    const T scale = (T)(1.0 / std::sqrt((double)dh));
    // ---- forward with every layer's activations kept --------------------------------------------------------
    DQ_LAUNCH(embed_kernel<T>, dim3((rows + 7) / 8), dim3(128), sizeof(T) * 5 * F, st, r, R, Rb, N, M, cfg.n_up, 1, 1, 1,
              P("emb.w"), d, X[0], rows, 8, (const T*)nullptr);
    for (int l = 0; l < L; ++l) {
      const std::string q = "L" + std::to_string(l) + ".";
      gemm(X[l], d, (q + "wqkv").c_str(), nullptr, 0, 3 * d, nullptr, nullptr, 0, QKV[l], 3 * d, rows, 3 * d, d, 1, 0, N, st);
      const T* kn = Mn > 0 ? P(q + "kn") : nullptr;
      const T* vn = Mn > 0 ? P(q + "vn") : nullptr;
      DQ_LAUNCH(attn_fl_kernel<T>, dim3(Bc, H), dim3(128), attn_smem_bytes<T>(N, dh, 1, Mn), st, (const T*)QKV[l], 3 * d, O[l], d, N,
                1, dh, d, scale, 1, kn, vn, Mn);
      gemm(O[l], d, (q + "wo").c_str(), nullptr, 0, d, nullptr, X[l], d, A[l], d, rows, d, d, 1, 0, N, st);
      gemm(A[l], d, (q + "w1").c_str(), nullptr, 0, d, P(q + "b1"), nullptr, 0, M1[l], d, rows, d, d, 1, 0, N, st);
      DQ_LAUNCH(tanh_fl_kernel<T>, dim3(rows, (d + 127) / 128), dim3(128), 0, st, M1[l], d, (const T*)nullptr, 0, 1, d, T(1));
      gemm(M1[l], d, (q + "w2").c_str(), nullptr, 0, d, P(q + "b2"), nullptr, 0, X[l + 1], d, rows, d, d, 1, 0, N, st);
      DQ_LAUNCH(tanh_fl_kernel<T>, dim3(rows, (d + 127) / 128), dim3(128), 0, st, X[l + 1], d, (const T*)A[l], d, 1, d, T(1));
    }
    gemm(X[L], d, "bf.up", "bf.dn", cfg.n_up, KN, nullptr, nullptr, 0, BF, KN, Bc, KN, d, 1, 1, N, st);
    const int sl_wpb = slater_warps_per_block<T>(N);
    DQ_LAUNCH(slater_kernel<T>, dim3((Bc * K + sl_wpb - 1) / sl_wpb), dim3(32 * sl_wpb), slater_smem_bytes<T>(N), st, r, R, Rb, N,
              M, cfg.n_up, K, 1, Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN,
              dsign, dlog, (T*)nullptr, (T*)nullptr, env_rep, 1, (const T*)nullptr, (const T*)nullptr, 0, 1);
    FinalizeCfg fc;
    fc.N = N; fc.M = M; fc.n_up = cfg.n_up; fc.K = K; fc.S = 1; fc.cusp_kind = cfg.cusp_kind;
    fc.cusp_same_scale = cfg.cusp_same_scale; fc.cusp_anti_scale = cfg.cusp_anti_scale; fc.ecp_terms = 0;
    fc.nuc_cusp_kind = cfg.nuc_cusp_kind;
    DQ_LAUNCH(finalize_kernel<T>, dim3(Bc), dim3(128), finalize_smem_bytes<T>(N, K), st, fc, r, R, Rb, (const T*)dsign,
              (const T*)dlog, (const T*)nullptr, (const T*)nullptr, P("cusp.alpha"), (const T*)d_zval, (const T*)nullptr,
              (const int*)d_ecp_mask, Bc, sign, logp, (T*)nullptr, (T*)nullptr, (T*)nullptr, (const T*)nullptr, (const T*)nullptr,
              cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr, PhArgs<T>());
    // ---- reverse ------------------------------------------------------------------------------------------------
    DQ_LAUNCH(finalize_bwd_kernel<T>, dim3((Bc + 127) / 128), dim3(128), 0, st, r, N, cfg.n_up, K, Bc, (const T*)dsign,
              (const T*)dlog, wts, cfg.cusp_kind, (T)cfg.cusp_same_scale, (T)cfg.cusp_anti_scale, P("cusp.alpha"), dld,
              G + off("cusp.alpha"), R, Rb, M, cfg.nuc_cusp_kind, cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr,
              cfg.nuc_cusp_kind ? G + off("cusp.nuc") : (T*)nullptr, (const T*)nullptr, (T*)nullptr);
    {
      const size_t pw = slater_bwd_smem_per_warp<T>(N);
      int wpb = (int)((96 * 1024) / pw);
      wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
      DQ_LAUNCH(slater_bwd_kernel<T>, dim3((Bc * K + wpb - 1) / wpb), dim3(32 * wpb), pw * wpb, st, r, R, Rb, N, M, cfg.n_up, K,
                Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN, (const T*)dld, dBF,
                G + off("env.pi_up"), G + off("env.pi_dn"), G + off("env.zeta_up"), G + off("env.zeta_dn"), env_rep, 1);
    }
    // backflow heads: dX_L = dBF W_spin^T, dW_spin += X_L[spin rows]^T dBF[spin rows]
    gemm_raw(dBF, KN, PT("bf.up"), PT("bf.dn"), cfg.n_up, d, nullptr, 0, dXn, d, Bc, d, KN, 1, st);
    wgrad(X[L], d, dBF, KN, rows, d, KN, G + off("bf.up"), 0, cfg.n_up, st);
    wgrad(X[L], d, dBF, KN, rows, d, KN, G + off("bf.dn"), cfg.n_up, N, st);
    const size_t nel = (size_t)rows * d;
    for (int l = L - 1; l >= 0; --l) {
      const std::string q = "L" + std::to_string(l) + ".";
      // X_{l+1} = A + tanh(M1 W2 + b2)
      DQ_LAUNCH(tanh_bwd_kernel<T>, dim3((unsigned)((nel + 255) / 256)), dim3(256), 0, st, (const T*)dXn, (const T*)X[l + 1],
                (const T*)A[l], dZ, nel);
      bgrad(dZ, d, rows, d, G + off(q + "b2"), st);
      wgrad(M1[l], d, dZ, d, rows, d, d, G + off(q + "w2"), 0, 0, st);
      gemm_raw(dZ, d, PT(q + "w2"), nullptr, 0, d, nullptr, 0, dM1, d, rows, d, d, 0, st);
      // M1 = tanh(A W1 + b1)
      DQ_LAUNCH(tanh_bwd_kernel<T>, dim3((unsigned)((nel + 255) / 256)), dim3(256), 0, st, (const T*)dM1, (const T*)M1[l],
                (const T*)nullptr, dZ, nel);
      bgrad(dZ, d, rows, d, G + off(q + "b1"), st);
      wgrad(A[l], d, dZ, d, rows, d, d, G + off(q + "w1"), 0, 0, st);
      gemm_raw(dZ, d, PT(q + "w1"), nullptr, 0, d, dXn, d, dA, d, rows, d, d, 0, st);  // dA = dX_{l+1} + dZ1 W1^T
      // A = X + O Wo
      wgrad(O[l], d, dA, d, rows, d, d, G + off(q + "wo"), 0, 0, st);
      gemm_raw(dA, d, PT(q + "wo"), nullptr, 0, d, nullptr, 0, dO, d, rows, d, d, 0, st);
      DQ_LAUNCH(attn_bwd_kernel<T>, dim3(Bc, H), dim3(128), attn_bwd_smem_bytes<T>(N, dh, Mn), st, (const T*)QKV[l], 3 * d,
                (const T*)dO, d, N, dh, d, scale, dQKV, Mn > 0 ? P(q + "kn") : (const T*)nullptr,
                Mn > 0 ? P(q + "vn") : (const T*)nullptr, Mn, Mn > 0 ? G + off(q + "kn") : (T*)nullptr,
                Mn > 0 ? G + off(q + "vn") : (T*)nullptr);
      wgrad(X[l], d, dQKV, 3 * d, rows, d, 3 * d, G + off(q + "wqkv"), 0, 0, st);
      gemm_raw(dQKV, 3 * d, PT(q + "wqkv"), nullptr, 0, d, dA, d, dX, d, rows, d, 3 * d, 0, st);  // dX_l = dA + dQKV Wqkv^T
      T* t = dXn; dXn = dX; dX = t;
    }
    DQ_LAUNCH(embed_feat_kernel<T>, dim3((rows * M + 127) / 128), dim3(128), 0, st, r, R, Rb, N, M, cfg.n_up, Feat, rows);
    wgrad(Feat, F, dXn, d, rows, F, d, G + off("emb.w"), 0, 0, st);
    note_hwm(p);
    if ((int64_t)(p - (char*)wsbase) > vjp_ws_cap) { err = "internal: reverse-pass buffers exceed the planned workspace"; return 3; }
    return 0;
  }

  // FermiNet reverse pass (conf/ansatz/ferminet.yaml): node update g on concat[h, spin means of h, spin means of the
  // incoming edges], shared edge MLP u, residuals / sqrt(2).  The raw input features carry no parameters, so the
  // chain stops at the first layer's weights.
  int vjp_chunk_ferminet(const T* r, const T* R, int Rb, int Bc, const T* wts, T* sign, T* logp, T* G, void* wsbase,
                         cudaStream_t st) {
    const int L = cfg.n_layers, rows = Bc * N, rowsE = Bc * N * N, de = cfg.edge_dim, d0 = 4 * M;
    const T isq2 = (T)0.70710678118654752440;
    char* p = (char*)wsbase;
    auto take = [&](size_t n) { T* q = (T*)p; p += align_up(sizeof(T) * n); DQ_TAKE_GUARD(); return q; };
    std::vector<T*> Hs(L + 1), Es(L), Fs(L);
    std::vector<int> dH(L + 1), dEd(L);
    dH[0] = d0;
    for (int l = 1; l <= L; ++l) dH[l] = d;
    for (int l = 0; l < L; ++l) dEd[l] = l == 0 ? 4 : de;
    for (int l = 0; l <= L; ++l) Hs[l] = take((size_t)rows * dH[l]);
    for (int l = 0; l < L; ++l) { Es[l] = take((size_t)rowsE * dEd[l]); Fs[l] = take((size_t)rows * (3 * dH[l] + 2 * dEd[l])); }
    const int fmax = 3 * (d > d0 ? d : d0) + 2 * (de > 4 ? de : 4);
    T* BF = take((size_t)rows * KN); T* dBF = take((size_t)rows * KN);
    T* dsign = take((size_t)Bc * K); T* dlog = take((size_t)Bc * K); T* dld = take((size_t)Bc * K);
    T* dXa = take((size_t)rows * d); T* dXb = take((size_t)rows * d); T* dZ = take((size_t)rows * d);
    T* dF = take((size_t)rows * fmax);
    T* dEa = take((size_t)rowsE * de); T* dEb = take((size_t)rowsE * de); T* dZe = take((size_t)rowsE * de);
    // ---- forward, activations kept -----------------------------------------------------------------------------
    DQ_LAUNCH(embed_kernel<T>, dim3(Bc * N), dim3(128), sizeof(T) * 5 * d0, st, r, R, Rb, N, M, cfg.n_up, 1, 0, 0,
              (const T*)nullptr, d0, Hs[0], Bc * N, 1, (const T*)nullptr);
    DQ_LAUNCH(edge_feat_kernel<T>, dim3((rowsE + 127) / 128), dim3(128), 0, st, r, N, 1, Es[0], rowsE, (const T*)nullptr);
    for (int l = 0; l < L; ++l) {
      const std::string q = "F" + std::to_string(l) + ".";
      const int dc = dH[l], ec = dEd[l], fin = 3 * dc + 2 * ec;
      DQ_LAUNCH(fermi_agg_kernel<T>, dim3(Bc, N), dim3(128), 0, st, (const T*)Hs[l], dc, (const T*)Es[l], ec, N, cfg.n_up, 1, Fs[l]);
      int rc = gemm(Fs[l], fin, (q + "wg").c_str(), nullptr, 0, d, P(q + "bg"), nullptr, 0, Hs[l + 1], d, rows, d, fin, 1, 0, N, st);
      if (rc) return rc;
      DQ_LAUNCH(tanh_fl_kernel<T>, dim3(rows, (d + 127) / 128), dim3(128), 0, st, Hs[l + 1], d,
                (const T*)(dc == d ? Hs[l] : nullptr), dc, 1, d, dc == d ? isq2 : T(1));
      if (l < L - 1) {
        rc = gemm(Es[l], ec, (q + "wu").c_str(), nullptr, 0, de, P(q + "bu"), nullptr, 0, Es[l + 1], de, rowsE, de, ec, 1, 0, N, st);
        if (rc) return rc;
        DQ_LAUNCH(tanh_fl_kernel<T>, dim3(rowsE, 1), dim3(32), 0, st, Es[l + 1], de, (const T*)(ec == de ? Es[l] : nullptr), ec,
                  1, de, ec == de ? isq2 : T(1));
      }
    }
    gemm(Hs[L], d, "bf.up", "bf.dn", cfg.n_up, KN, nullptr, nullptr, 0, BF, KN, Bc, KN, d, 1, 1, N, st);
    const int sl_wpb = slater_warps_per_block<T>(N);
    DQ_LAUNCH(slater_kernel<T>, dim3((Bc * K + sl_wpb - 1) / sl_wpb), dim3(32 * sl_wpb), slater_smem_bytes<T>(N), st, r, R, Rb, N,
              M, cfg.n_up, K, 1, Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN,
              dsign, dlog, (T*)nullptr, (T*)nullptr, env_rep, 1, (const T*)nullptr, (const T*)nullptr, 0, 1);
    FinalizeCfg fc;
    fc.N = N; fc.M = M; fc.n_up = cfg.n_up; fc.K = K; fc.S = 1; fc.cusp_kind = cfg.cusp_kind;
    fc.cusp_same_scale = cfg.cusp_same_scale; fc.cusp_anti_scale = cfg.cusp_anti_scale; fc.ecp_terms = 0;
    fc.nuc_cusp_kind = cfg.nuc_cusp_kind;
    DQ_LAUNCH(finalize_kernel<T>, dim3(Bc), dim3(128), finalize_smem_bytes<T>(N, K), st, fc, r, R, Rb, (const T*)dsign,
              (const T*)dlog, (const T*)nullptr, (const T*)nullptr, P("cusp.alpha"), (const T*)d_zval, (const T*)nullptr,
              (const int*)d_ecp_mask, Bc, sign, logp, (T*)nullptr, (T*)nullptr, (T*)nullptr, (const T*)nullptr, (const T*)nullptr,
              cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr, PhArgs<T>());
    // ---- reverse -------------------------------------------------------------------------------------------------
    DQ_LAUNCH(finalize_bwd_kernel<T>, dim3((Bc + 127) / 128), dim3(128), 0, st, r, N, cfg.n_up, K, Bc, (const T*)dsign,
              (const T*)dlog, wts, cfg.cusp_kind, (T)cfg.cusp_same_scale, (T)cfg.cusp_anti_scale, P("cusp.alpha"), dld,
              G + off("cusp.alpha"), R, Rb, M, cfg.nuc_cusp_kind, cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr,
              cfg.nuc_cusp_kind ? G + off("cusp.nuc") : (T*)nullptr, (const T*)nullptr, (T*)nullptr);
    {
      const size_t pw = slater_bwd_smem_per_warp<T>(N);
      int wpb = (int)((96 * 1024) / pw);
      wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
      DQ_LAUNCH(slater_bwd_kernel<T>, dim3((Bc * K + wpb - 1) / wpb), dim3(32 * wpb), pw * wpb, st, r, R, Rb, N, M, cfg.n_up, K,
                Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN, (const T*)dld, dBF,
                G + off("env.pi_up"), G + off("env.pi_dn"), G + off("env.zeta_up"), G + off("env.zeta_dn"), env_rep, 1);
    }
    T* dHn = dXa;   // gradient w.r.t. H_{l+1}
    T* dHc = dXb;   // gradient w.r.t. H_l (being built)
    T* dEn = dEa;   // gradient w.r.t. E_{l+1} (valid for l < L - 1)
    T* dEc = dEb;
    gemm_raw(dBF, KN, PT("bf.up"), PT("bf.dn"), cfg.n_up, d, nullptr, 0, dHn, d, Bc, d, KN, 1, st);
    wgrad(Hs[L], d, dBF, KN, rows, d, KN, G + off("bf.up"), 0, cfg.n_up, st);
    wgrad(Hs[L], d, dBF, KN, rows, d, KN, G + off("bf.dn"), cfg.n_up, N, st);
    for (int l = L - 1; l >= 0; --l) {
      const std::string q = "F" + std::to_string(l) + ".";
      const int dc = dH[l], ec = dEd[l], fin = 3 * dc + 2 * ec;
      const bool res_h = dc == d, res_e = ec == de;
      const size_t nel = (size_t)rows * d;
      // H_{l+1} = s (H_l + tanh(F Wg + bg))  |  tanh(F Wg + bg)
      DQ_LAUNCH(tanh_res_bwd_kernel<T>, dim3((unsigned)((nel + 255) / 256)), dim3(256), 0, st, (const T*)dHn, (const T*)Hs[l + 1],
                (const T*)(res_h ? Hs[l] : nullptr), res_h ? isq2 : T(1), dZ, nel);
      bgrad(dZ, d, rows, d, G + off(q + "bg"), st);
      wgrad(Fs[l], fin, dZ, d, rows, fin, d, G + off(q + "wg"), 0, 0, st);
      if (l > 0) {
        gemm_raw(dZ, d, PT(q + "wg"), nullptr, 0, fin, nullptr, 0, dF, fin, rows, fin, d, 0, st);
        DQ_LAUNCH(fermi_agg_bwd_kernel<T>, dim3(Bc, N), dim3(128), 0, st, (const T*)dF, dc, ec, N, cfg.n_up,
                  (const T*)(res_h ? dHn : nullptr), isq2, dHc, dEc);
      }
      if (l < L - 1) {
        // E_{l+1} = s (E_l + tanh(E_l Wu + bu))  |  tanh(E_l Wu + bu)
        const size_t nee = (size_t)rowsE * de;
        DQ_LAUNCH(tanh_res_bwd_kernel<T>, dim3((unsigned)((nee + 255) / 256)), dim3(256), 0, st, (const T*)dEn, (const T*)Es[l + 1],
                  (const T*)(res_e ? Es[l] : nullptr), res_e ? isq2 : T(1), dZe, nee);
        bgrad(dZe, de, rowsE, de, G + off(q + "bu"), st);
        wgrad(Es[l], ec, dZe, de, rowsE, ec, de, G + off(q + "wu"), 0, 0, st);
        if (l > 0) {
          gemm_raw(dZe, de, PT(q + "wu"), nullptr, 0, ec, dEc, ec, dEc, ec, rowsE, ec, de, 0, st);  // dE_l += dZe Wu^T
          if (res_e) {
            const size_t ne2 = (size_t)rowsE * ec;
            DQ_LAUNCH(axpy_kernel<T>, dim3((unsigned)((ne2 + 255) / 256)), dim3(256), 0, st, (const T*)dEn, isq2, dEc, ne2);
          }
        }
      }
      T* t1 = dHn; dHn = dHc; dHc = t1;
      T* t2 = dEn; dEn = dEc; dEc = t2;
    }
    note_hwm(p);
    if ((int64_t)(p - (char*)wsbase) > vjp_ws_cap) { err = "internal: reverse-pass buffers exceed the planned workspace"; return 3; }
    return 0;
  }
  // ---- conv-GNN reverse pass: the reference's test ansatz (tests/conf/ansatz.yaml: hk.Embed embeddings, 'featurewise'
  // update over same / anti / ne convolutions, no deep edge features), Jastrow and per-spin backflow MLPs (ssp), default
  // mult_act, spin-factorised determinants, hk.Linear determinant weights.  Value layouts of kernels_bwd.cuh.
  struct Tape {
    std::vector<T*> a;     // a[0] input, a[k] output of layer k (after its activation)
    std::vector<int> dim;  // widths
  };
  template <class Take>
  int tape_fwd(const T* in, int din, const int* dims, int nl, const std::string& base, bool bias, int act, bool last_linear,
               int rows_, Take& take, Tape& t, cudaStream_t st) {
    t.a.assign(1, const_cast<T*>(in));
    t.dim.assign(1, din);
    for (int i = 0; i < nl; ++i) {
      T* out = take((size_t)rows_ * dims[i]);
      const std::string q = base + std::to_string(i);
      const bool b_i = bias && (!last_linear || base[0] != 'J' || i < nl - 1);  // Jastrow: bias 'not_last'
      int rc = gemm(t.a.back(), t.dim.back(), (q + ".w").c_str(), nullptr, 0, dims[i], b_i ? P(q + ".b") : nullptr, nullptr, 0, out,
                    dims[i], rows_, dims[i], t.dim.back(), 1, 0, 1, st);
      if (rc) return rc;
      if (i < nl - 1 || !last_linear)
        DQ_LAUNCH(act_fl_kernel<T>, dim3(rows_, (dims[i] + 63) / 64), dim3(64), 0, st, out, dims[i], (const T*)nullptr, 0, 1, dims[i],
                  T(1), act);
      t.a.push_back(out);
      t.dim.push_back(dims[i]);
    }
    return 0;
  }
  // dY = gradient w.r.t. the tape's output (destroyed).  s0 / s1: scratch of rows_ x max width.  dIn (nullable) receives
  // (accumulate: is increased by) the gradient w.r.t. the tape's input.
  int tape_bwd(const Tape& t, T* dY, const std::string& base, bool bias, int act, bool last_linear, int rows_, T* s0, T* s1,
               T* dIn, bool accumulate, T* G, cudaStream_t st) {
    const int nl = (int)t.a.size() - 1;
    T* cur = dY;
    for (int i = nl - 1; i >= 0; --i) {
      const int dout = t.dim[i + 1], din = t.dim[i];
      const std::string q = base + std::to_string(i);
      const size_t n = (size_t)rows_ * dout;
      if (i < nl - 1 || !last_linear)
        DQ_LAUNCH(act_bwd_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, cur, (const T*)t.a[i + 1], act, n);
      const bool b_i = bias && (!last_linear || base[0] != 'J' || i < nl - 1);
      if (b_i) bgrad(cur, dout, rows_, dout, G + off(q + ".b"), st);
      wgrad(t.a[i], din, cur, dout, rows_, din, dout, G + off(q + ".w"), 0, 0, st);
      if (i > 0) {
        T* nxt = (cur == s0) ? s1 : s0;
        gemm_raw(cur, dout, PT(q + ".w"), nullptr, 0, din, nullptr, 0, nxt, din, rows_, din, dout, 0, st);
        cur = nxt;
      } else if (dIn) {
        gemm_raw(cur, dout, PT(q + ".w"), nullptr, 0, din, accumulate ? dIn : nullptr, din, dIn, din, rows_, din, dout, 0, st);
      }
    }
    return 0;
  }
  int vjp_chunk_paulinet(const T* r, const T* R, int Rb, int Bc, const T* wts, T* sign, T* logp, T* G, void* wsbase,
                         cudaStream_t st) {
    const int L = cfg.n_layers, rows = Bc * N, e = cfg.edge_dim, nl = cfg.gnn_sub_n > 0 ? cfg.gnn_sub_n : 1;
    const int Mne = cfg.gnn_conv_ne ? M : 0, NS = N + Mne, nt = cfg.gnn_conv_ne ? 3 : 2, pairs = Bc * N * NS;
    const int n_types = cfg.n_elec_types > 0 ? cfg.n_elec_types : 1;
    const bool deep = cfg.gnn_deep_edges != 0;
    const T isq2 = (T)0.70710678118654752440;
    const char* tn[3] = {"same", "anti", "ne"};
    char* p = (char*)wsbase;
    auto take = [&](size_t n) { T* q = (T*)p; p += align_up(sizeof(T) * n); DQ_TAKE_GUARD(); return q; };
    // ---- forward, everything kept --------------------------------------------------------------------------------
    std::vector<T*> X(L + 1), C(L), Fc(L), E(L);
    std::vector<int> xd(L + 1), ed(L);
    std::vector<std::array<T*, 3>> Gt(L);
    std::vector<std::array<Tape, 3>> Wt(L);
    std::vector<std::array<Tape, 2>> Ht(L);
    std::vector<Tape> Ut(L);
    xd[0] = cfg.gnn_features ? 4 * M : d;
    X[0] = take((size_t)rows * xd[0]);
    if (cfg.gnn_features)  // raw nucleus-electron features [|d|, d]: no parameters
      DQ_LAUNCH(embed_kernel<T>, dim3(rows), dim3(128), sizeof(T) * 5 * xd[0], st, r, R, Rb, N, M, cfg.n_up, 1, 0, 0, (const T*)nullptr,
                xd[0], X[0], rows, 1, (const T*)nullptr);
    else
      DQ_LAUNCH(gnn_embed_kernel<T>, dim3((rows * d + 127) / 128), dim3(128), 0, st, P("emb.table"), n_types, N, cfg.n_up, 1, d, X[0], rows);
    E[0] = take((size_t)pairs * 4);
    ed[0] = 4;
    DQ_LAUNCH(gnn_edge_val_kernel<T>, dim3((pairs + 127) / 128), dim3(128), 0, st, r, R, Rb, N, M, Mne, E[0], pairs);
    for (int l = 0; l < L; ++l) {
      const std::string q = "G" + std::to_string(l) + ".";
      for (int t = 0; t < nt; ++t) {
        int rc = tape_fwd(E[l], ed[l], cfg.gnn_w_dims[l], nl, q + "w_" + tn[t] + ".", cfg.gnn_w_bias != 0, 0, false, pairs, take, Wt[l][t], st);
        if (rc) return rc;
      }
      for (int t = 0; t < 2; ++t) {
        int rc = tape_fwd(X[l], xd[l], cfg.gnn_h_dims[l], nl, q + "h_" + tn[t] + ".", true, 0, false, rows, take, Ht[l][t], st);
        if (rc) return rc;
      }
      C[l] = take((size_t)rows * nt * e);
      DQ_LAUNCH(gnn_conv_val_kernel<T>, dim3(rows), dim3(64), 0, st, (const T*)Wt[l][0].a.back(), (const T*)Wt[l][1].a.back(),
                (const T*)(nt == 3 ? Wt[l][2].a.back() : nullptr), (const T*)Ht[l][0].a.back(), (const T*)Ht[l][1].a.back(),
                nt == 3 ? P(q + "hne") : (const T*)nullptr, N, Mne, cfg.n_up, e, C[l]);
      xd[l + 1] = d;
      if (cfg.gnn_concat) {  // x <- [(x +) tanh(g([x, mean_up x, mean_down x, conv_*]))] (/ sqrt 2)
        const int fin = 3 * xd[l] + nt * e;
        Fc[l] = take((size_t)rows * fin);
        DQ_LAUNCH(gnn_concat_kernel<T>, dim3(Bc, N), dim3(128), 0, st, (const T*)X[l], xd[l], (const T*)C[l], nt * e, N, cfg.n_up, 1, Fc[l]);
        X[l + 1] = take((size_t)rows * d);
        int rc = gemm(Fc[l], fin, (q + "g.w").c_str(), nullptr, 0, d, cfg.gnn_g_bias ? P(q + "g.b") : nullptr, nullptr, 0, X[l + 1], d,
                      rows, d, fin, 1, 0, N, st);
        if (rc) return rc;
        const bool res = xd[l] == d;
        DQ_LAUNCH(act_fl_kernel<T>, dim3(rows, (d + 63) / 64), dim3(64), 0, st, X[l + 1], d, res ? (const T*)X[l] : (const T*)nullptr, d, 1,
                  d, (res && cfg.gnn_res_norm) ? isq2 : T(1), 0);
      } else {  // featurewise: x <- x + sum_t tanh(g_t(conv_t)); each G_t holds the running sum
        const T* res = xd[l] == d ? X[l] : nullptr;
        for (int t = 0; t < nt; ++t) {
          Gt[l][t] = take((size_t)rows * d);
          int rc = gemm(C[l] + t * e, nt * e, (q + "g_" + tn[t] + ".w").c_str(), nullptr, 0, d, P(q + "g_" + tn[t] + ".b"), nullptr, 0,
                        Gt[l][t], d, rows, d, e, 1, 0, N, st);
          if (rc) return rc;
          DQ_LAUNCH(act_fl_kernel<T>, dim3(rows, (d + 63) / 64), dim3(64), 0, st, Gt[l][t], d, res, d, 1, d, T(1), 0);
          res = Gt[l][t];
        }
        if (cfg.gnn_res_norm) { err = "dqmc_wf_vjp_params: normalised featurewise residual not supported"; return 2; }
        X[l + 1] = Gt[l][nt - 1];
      }
      if (deep && l < L - 1) {  // shared edge MLP u + normalised residual (electron_gnn.py:160-192)
        int rc = tape_fwd(E[l], ed[l], cfg.gnn_u_dims[l], nl, q + "u.", true, 0, false, pairs, take, Ut[l], st);
        if (rc) return rc;
        ed[l + 1] = e;
        if (ed[l] == e) {
          E[l + 1] = take((size_t)pairs * e);
          DQ_LAUNCH((axpby_kernel<T>), dim3((unsigned)(((size_t)pairs * e + 255) / 256)), dim3(256), 0, st, (const T*)E[l],
                    (const T*)Ut[l].a.back(), isq2, E[l + 1], (size_t)pairs * e);
        } else {
          E[l + 1] = Ut[l].a.back();
        }
      } else if (l < L - 1) {
        E[l + 1] = E[l]; ed[l + 1] = ed[l];
      }
    }
    // Jastrow on sum_i x_i
    Tape Jt;
    T* Js = nullptr;
    if (cfg.jastrow_n > 0) {
      Js = take((size_t)Bc * d);
      DQ_LAUNCH(sum_electrons_kernel<T>, dim3((Bc * d + 127) / 128), dim3(128), 0, st, (const T*)X[L], N, 1, d, Js, Bc * d);
      int rc = tape_fwd(Js, d, cfg.jastrow_dims, cfg.jastrow_n, "J", true, 1, true, Bc, take, Jt, st);
      if (rc) return rc;
    }
    // per-spin backflow MLPs: hidden layers (ssp), then the orbital head (+ default mult_act)
    std::vector<T*> Y(cfg.backflow_n + 1);
    std::vector<int> yd(cfg.backflow_n + 1);
    Y[0] = X[L]; yd[0] = d;
    for (int i = 0; i < cfg.backflow_n; ++i) {
      const int dout = cfg.backflow_dims[i];
      const std::string q = std::to_string(i);
      Y[i + 1] = take((size_t)rows * dout); yd[i + 1] = dout;
      int rc = gemm(Y[i], yd[i], ("bfh" + q + ".up").c_str(), ("bfh" + q + ".dn").c_str(), cfg.n_up, dout, P("bfb" + q + ".up"), nullptr,
                    0, Y[i + 1], dout, Bc, dout, yd[i], 1, 1, N, st, 0, P("bfb" + q + ".dn"));
      if (rc) return rc;
      DQ_LAUNCH(act_fl_kernel<T>, dim3(rows, (dout + 31) / 32), dim3(32), 0, st, Y[i + 1], dout, (const T*)nullptr, 0, 1, dout, T(1), 1);
    }
    T* BF = take((size_t)rows * KN); T* dBF = take((size_t)rows * KN);
    int rc = gemm(Y.back(), yd.back(), "bf.up", "bf.dn", cfg.n_up, KN, P("bfb.up"), nullptr, 0, BF, KN, Bc, KN, yd.back(), 1, 1, N, st, 0,
                  P("bfb.dn"));
    if (rc) return rc;
    if (cfg.mult_act == 1)
      DQ_LAUNCH(act_fl_kernel<T>, dim3(rows, (KN + 127) / 128), dim3(128), 0, st, BF, KN, (const T*)nullptr, 0, 1, KN, T(1), 2);
    T* dsign = take((size_t)Bc * K); T* dlog = take((size_t)Bc * K); T* dld = take((size_t)Bc * K);
    const int full_det = cfg.factorized_det ? 0 : 1;
    const int sl_wpb = slater_warps_per_block<T>(N);
    DQ_LAUNCH(slater_kernel<T>, dim3((Bc * K + sl_wpb - 1) / sl_wpb), dim3(32 * sl_wpb), slater_smem_bytes<T>(N), st, r, R, Rb, N,
              M, cfg.n_up, K, 1, Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN,
              dsign, dlog, (T*)nullptr, (T*)nullptr, env_rep, full_det, (const T*)nullptr, (const T*)nullptr, 0, 1);
    FinalizeCfg fc;
    fc.N = N; fc.M = M; fc.n_up = cfg.n_up; fc.K = K; fc.S = 1; fc.cusp_kind = cfg.cusp_kind;
    fc.cusp_same_scale = cfg.cusp_same_scale; fc.cusp_anti_scale = cfg.cusp_anti_scale; fc.ecp_terms = 0;
    fc.nuc_cusp_kind = cfg.nuc_cusp_kind;
    DQ_LAUNCH(finalize_kernel<T>, dim3(Bc), dim3(128), finalize_smem_bytes<T>(N, K), st, fc, r, R, Rb, (const T*)dsign,
              (const T*)dlog, (const T*)nullptr, (const T*)nullptr, P("cusp.alpha"), (const T*)d_zval, (const T*)nullptr,
              (const int*)d_ecp_mask, Bc, sign, logp, (T*)nullptr, (T*)nullptr, (T*)nullptr,
              cfg.conf_linear ? P("conf.w") : (const T*)nullptr, cfg.jastrow_n > 0 ? (const T*)Jt.a.back() : (const T*)nullptr,
              cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr, PhArgs<T>());
    // ---- reverse ---------------------------------------------------------------------------------------------------
    DQ_LAUNCH(finalize_bwd_kernel<T>, dim3((Bc + 127) / 128), dim3(128), 0, st, r, N, cfg.n_up, K, Bc, (const T*)dsign,
              (const T*)dlog, wts, 0 /*fixed cusp exponent: no gradient*/, (T)cfg.cusp_same_scale, (T)cfg.cusp_anti_scale,
              P("cusp.alpha"), dld, (T*)nullptr, R, Rb, M, cfg.nuc_cusp_kind, cfg.nuc_cusp_kind ? P("cusp.nuc") : (const T*)nullptr,
              cfg.nuc_cusp_kind ? G + off("cusp.nuc") : (T*)nullptr, cfg.conf_linear ? P("conf.w") : (const T*)nullptr,
              cfg.conf_linear ? G + off("conf.w") : (T*)nullptr);
    {
      const size_t pw = slater_bwd_smem_per_warp<T>(N);
      int wpb = (int)((96 * 1024) / pw);
      wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
      DQ_LAUNCH(slater_bwd_kernel<T>, dim3((Bc * K + wpb - 1) / wpb), dim3(32 * wpb), pw * wpb, st, r, R, Rb, N, M, cfg.n_up, K,
                Bc * K, P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), (const T*)BF, KN, (const T*)dld, dBF,
                G + off("env.pi_up"), G + off("env.pi_dn"), G + off("env.zeta_up"), G + off("env.zeta_dn"), env_rep, full_det);
    }
    // scratch for the reverse sweep
    const int xm = std::max(d, xd[0]);
    const int hm = std::max(gnn_hmax(), xm), hn = std::max(gnn_hnode_max(), e), em = gnn_emax();
    const int fmax = 3 * xm + nt * e;
    T* dXa = take((size_t)rows * xm); T* dXb = take((size_t)rows * xm);
    T* sr0 = take((size_t)rows * std::max(hm, hn)); T* sr1 = take((size_t)rows * std::max(hm, hn));
    T* dCb = take((size_t)rows * nt * e); T* dCt = take((size_t)rows * e);
    T* dFb = cfg.gnn_concat ? take((size_t)rows * fmax) : nullptr;
    T* dWb[3] = {take((size_t)pairs * e), take((size_t)pairs * e), take((size_t)pairs * e)};
    T* sp0 = take((size_t)pairs * em); T* sp1 = take((size_t)pairs * em);
    T* dHb[2] = {take((size_t)rows * e), take((size_t)rows * e)};
    T* dEa = deep ? take((size_t)pairs * e) : nullptr;
    T* dEb = deep ? take((size_t)pairs * e) : nullptr;
    T* dUb = deep ? take((size_t)pairs * e) : nullptr;
    // orbital head
    if (cfg.mult_act == 1) {
      const size_t n = (size_t)rows * KN;
      DQ_LAUNCH(act_bwd_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, dBF, (const T*)BF, 2, n);
    }
    bgrad(dBF, KN, rows, KN, G + off("bfb.up"), st, 0, cfg.n_up);
    bgrad(dBF, KN, rows, KN, G + off("bfb.dn"), st, cfg.n_up, N);
    wgrad(Y.back(), yd.back(), dBF, KN, rows, yd.back(), KN, G + off("bf.up"), 0, cfg.n_up, st);
    wgrad(Y.back(), yd.back(), dBF, KN, rows, yd.back(), KN, G + off("bf.dn"), cfg.n_up, N, st);
    T* cur = cfg.backflow_n > 0 ? sr0 : dXa;
    gemm_raw(dBF, KN, PT("bf.up"), PT("bf.dn"), cfg.n_up, yd.back(), nullptr, 0, cur, yd.back(), Bc, yd.back(), KN, 1, st);
    for (int i = cfg.backflow_n - 1; i >= 0; --i) {
      const int dout = yd[i + 1], din = yd[i];
      const std::string q = std::to_string(i);
      const size_t n = (size_t)rows * dout;
      DQ_LAUNCH(act_bwd_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, cur, (const T*)Y[i + 1], 1, n);
      bgrad(cur, dout, rows, dout, G + off("bfb" + q + ".up"), st, 0, cfg.n_up);
      bgrad(cur, dout, rows, dout, G + off("bfb" + q + ".dn"), st, cfg.n_up, N);
      wgrad(Y[i], din, cur, dout, rows, din, dout, G + off("bfh" + q + ".up"), 0, cfg.n_up, st);
      wgrad(Y[i], din, cur, dout, rows, din, dout, G + off("bfh" + q + ".dn"), cfg.n_up, N, st);
      T* nxt = i == 0 ? dXa : (cur == sr0 ? sr1 : sr0);
      gemm_raw(cur, dout, PT("bfh" + q + ".up"), PT("bfh" + q + ".dn"), cfg.n_up, din, nullptr, 0, nxt, din, Bc, din, dout, 1, st);
      cur = nxt;
    }
    T* dXn = dXa;  // gradient w.r.t. X_L
    T* dXc = dXb;
    if (cfg.jastrow_n > 0) {  // d log|psi| / d jastrow = w_b
      T* dJ = take((size_t)Bc);  // [Bc] scalars
      T* js0 = take((size_t)Bc * d); T* js1 = take((size_t)Bc * d); T* dJs = take((size_t)Bc * d);
      DQ_CHECK(cudaMemcpyAsync(dJ, wts, sizeof(T) * Bc, cudaMemcpyDeviceToDevice, st));
      tape_bwd(Jt, dJ, "J", true, 1, true, Bc, js0, js1, dJs, false, G, st);
      const size_t n = (size_t)rows * d;
      DQ_LAUNCH(bcast_add_kernel<T>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, (const T*)dJs, N, d, dXn, n);
    }
    T* dEn = dEa;  // gradient w.r.t. E_{l+1} (deep edge features only)
    T* dEc = dEb;
    for (int l = L - 1; l >= 0; --l) {
      const std::string q = "G" + std::to_string(l) + ".";
      const size_t nd = (size_t)rows * d;
      const bool need_dx = l > 0 || !cfg.gnn_features;  // layer 0 of the raw-feature variant has nothing trainable upstream
      if (cfg.gnn_concat) {
        const int fin = 3 * xd[l] + nt * e;
        const bool res = xd[l] == d;
        const T sc = (res && cfg.gnn_res_norm) ? isq2 : T(1);
        DQ_LAUNCH(tanh_res_bwd_kernel<T>, dim3((unsigned)((nd + 255) / 256)), dim3(256), 0, st, (const T*)dXn, (const T*)X[l + 1],
                  (const T*)(res ? X[l] : nullptr), sc, sr0, nd);
        if (cfg.gnn_g_bias) bgrad(sr0, d, rows, d, G + off(q + "g.b"), st);
        wgrad(Fc[l], fin, sr0, d, rows, fin, d, G + off(q + "g.w"), 0, 0, st);
        gemm_raw(sr0, d, PT(q + "g.w"), nullptr, 0, fin, nullptr, 0, dFb, fin, rows, fin, d, 0, st);
        // dF -> dX_l (own row + spin means + residual) and dC (the convolution columns); F = [x, mean_up, mean_down, conv_*]
        if (nt != 2) { err = "dqmc_wf_vjp_params: concatenate update with nucleus-electron convolutions not supported"; return 2; }
        DQ_LAUNCH(fermi_agg_bwd_kernel<T>, dim3(Bc, N), dim3(128), 0, st, (const T*)dFb, xd[l], e, N, cfg.n_up,
                  (const T*)(res ? dXn : nullptr), sc, dXc, (T*)nullptr);
        const size_t nc = (size_t)rows * nt * e;
        DQ_LAUNCH(slice_cols_kernel<T>, dim3((unsigned)((nc + 255) / 256)), dim3(256), 0, st, (const T*)dFb, fin, 3 * xd[l], nt * e, dCb, nc);
      } else {
        // X_{l+1} = X_l + sum_t tanh(z_t): tanh outputs are the differences of the running sums
        if (xd[l] == d) DQ_CHECK(cudaMemcpyAsync(dXc, dXn, sizeof(T) * nd, cudaMemcpyDeviceToDevice, st));  // residual
        else DQ_CHECK(cudaMemsetAsync(dXc, 0, sizeof(T) * (size_t)rows * xd[l], st));
        for (int t = 0; t < nt; ++t) {
          const T* prev = t == 0 ? (xd[l] == d ? X[l] : nullptr) : Gt[l][t - 1];
          DQ_LAUNCH(tanh_bwd_kernel<T>, dim3((unsigned)((nd + 255) / 256)), dim3(256), 0, st, (const T*)dXn, (const T*)Gt[l][t], prev, sr0, nd);
          bgrad(sr0, d, rows, d, G + off(q + "g_" + tn[t] + ".b"), st);
          const size_t ne_ = (size_t)rows * e;  // conv_t is a column slice of C: copy it out for the weight gradient
          DQ_LAUNCH(slice_cols_kernel<T>, dim3((unsigned)((ne_ + 255) / 256)), dim3(256), 0, st, (const T*)C[l], nt * e, t * e, e, dCt, ne_);
          wgrad(dCt, e, sr0, d, rows, e, d, G + off(q + "g_" + tn[t] + ".w"), 0, 0, st);
          gemm_raw(sr0, d, PT(q + "g_" + tn[t] + ".w"), nullptr, 0, e, nullptr, 0, dCb + t * e, nt * e, rows, e, d, 0, st);
        }
      }
      DQ_CHECK(cudaMemsetAsync(dHb[0], 0, sizeof(T) * (size_t)rows * e, st));
      DQ_CHECK(cudaMemsetAsync(dHb[1], 0, sizeof(T) * (size_t)rows * e, st));
      DQ_LAUNCH(gnn_conv_bwd_kernel<T>, dim3(rows), dim3(64), 0, st, (const T*)dCb, (const T*)Wt[l][0].a.back(),
                (const T*)Wt[l][1].a.back(), (const T*)(nt == 3 ? Wt[l][2].a.back() : nullptr), (const T*)Ht[l][0].a.back(),
                (const T*)Ht[l][1].a.back(), nt == 3 ? P(q + "hne") : (const T*)nullptr, N, Mne, cfg.n_up, e, dWb[0], dWb[1], dWb[2],
                dHb[0], dHb[1], nt == 3 ? G + off(q + "hne") : (T*)nullptr);
      // gradient w.r.t. this layer's edge features E_l (only the deep-edge variant carries it; E_0 has nothing upstream)
      const bool need_de = deep && l > 0;
      if (need_de) DQ_CHECK(cudaMemsetAsync(dEc, 0, sizeof(T) * (size_t)pairs * ed[l], st));
      for (int t = 0; t < nt; ++t)
        tape_bwd(Wt[l][t], dWb[t], q + "w_" + tn[t] + ".", cfg.gnn_w_bias != 0, 0, false, pairs, sp0, sp1, need_de ? dEc : nullptr, true, G, st);
      for (int t = 0; t < 2; ++t)
        tape_bwd(Ht[l][t], dHb[t], q + "h_" + tn[t] + ".", true, 0, false, rows, sr0, sr1, need_dx ? dXc : nullptr, true, G, st);
      if (deep && l < L - 1) {  // E_{l+1} = s (E_l + u(E_l))  |  u(E_l); dEn holds the gradient w.r.t. E_{l+1}
        const bool res_e = ed[l] == e;
        const size_t ne2 = (size_t)pairs * e;
        DQ_LAUNCH((axpby_kernel<T>), dim3((unsigned)((ne2 + 255) / 256)), dim3(256), 0, st, (const T*)dEn, (const T*)dEn, res_e ? isq2 / T(2) : T(0.5),
                  dUb, ne2);  // dU = s dE_{l+1}
        tape_bwd(Ut[l], dUb, q + "u.", true, 0, false, pairs, sp0, sp1, need_de ? dEc : nullptr, true, G, st);
        if (need_de && res_e) DQ_LAUNCH(axpy_kernel<T>, dim3((unsigned)((ne2 + 255) / 256)), dim3(256), 0, st, (const T*)dEn, isq2, dEc, ne2);
      }
      T* tmp = dXn; dXn = dXc; dXc = tmp;
      T* tmq = dEn; dEn = dEc; dEc = tmq;
    }
    if (!cfg.gnn_features)
      DQ_LAUNCH(embed_table_bwd_kernel<T>, dim3((d + 63) / 64, 64), dim3(64), 0, st, (const T*)dXn, n_types, N, cfg.n_up, d, rows,
                G + off("emb.table"));
    note_hwm(p);
    if ((int64_t)(p - (char*)wsbase) > vjp_ws_cap) { err = "internal: reverse-pass buffers exceed the planned workspace"; return 3; }
    return 0;
  }

  int vjp_params(const void* r_, const void* R_, int Rb, int B, const void* weights, void* sign, void* logp,
                 void* grad_params, void* ws, int64_t wsb, cudaStream_t st) override {

    if (cfg.backflow_add) { err = "dqmc_wf_vjp_params: additive backflow branch has no reverse pass"; return 2; }
    const T* r = (const T*)r_;
    const T* R = (const T*)R_;
    DQ_CHECK(cudaMemsetAsync(grad_params, 0, sizeof(T) * total, st));
    if (B == 0) return 0;  // empty batch: zero gradient
    // walkers per chunk: activations of every layer stay resident for the reverse pass (64 buffers, 256 B alignment each)
    const bool fermi = cfg.kind == DQMC_FERMINET;
    // largest chunk whose buffers (measured by a dry pass of the chunk function) fit the caller's workspace
    int64_t Bc = B;
    if (vjp_chunk_bytes(1) > wsb) { err = "workspace too small for a single walker (vjp)"; return 3; }
    if (vjp_chunk_bytes((int)Bc) > wsb) {
      int64_t lo = 1, hi = Bc;
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) / 2;
        if (vjp_chunk_bytes((int)mid) <= wsb) lo = mid; else hi = mid;
      }
      Bc = lo;
    }
    vjp_ws_cap = wsb;
    if (!fermi && !gnn)
    DQ_CHECK(raise_dyn_smem(attn_bwd_kernel<T>, (int)attn_bwd_smem_bytes<T>(N, dh, Mn)));
    {  // the same warps-per-block rule as at the launch sites (at most 4 warps, at most 96 KiB)
      const size_t pw = slater_bwd_smem_per_warp<T>(N);
      int wpb = (int)((96 * 1024) / pw);
      wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
      if (pw * wpb > 227 * 1024) { err = "system too large for the shared-memory tiling of the reverse pass (N)"; return 2; }
      DQ_CHECK(raise_dyn_smem(slater_bwd_kernel<T>, (int)(pw * wpb)));
    }
    for (int b0 = 0; b0 < B; b0 += (int)Bc) {
      const int nb = (int)std::min<int64_t>(Bc, B - b0);
      const T* rc_ = r + (size_t)b0 * 3 * N;
      const T* Rc_ = R + (Rb ? (size_t)b0 * 3 * M : 0);
      int rc = gnn ? vjp_chunk_paulinet(rc_, Rc_, Rb, nb, (const T*)weights + b0, (T*)sign + b0, (T*)logp + b0, (T*)grad_params, ws, st)
             : fermi ? vjp_chunk_ferminet(rc_, Rc_, Rb, nb, (const T*)weights + b0, (T*)sign + b0, (T*)logp + b0, (T*)grad_params, ws, st)
                     : vjp_chunk(rc_, Rc_, Rb, nb, (const T*)weights + b0, (T*)sign + b0, (T*)logp + b0, (T*)grad_params, ws, st);
      if (rc) return rc;
      rc = check_guards();
      if (rc) return rc;
      if (dry) break;  // planning pass: the first chunk is the largest
    }
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  int stats_pack(const void* E, const void* stats, int B, double* out, cudaStream_t st) override {
    DQ_LAUNCH(stats_pack_kernel<T>, dim3(1), dim3(1024), 0, st, (const T*)E, (const T*)stats, B, out);
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  // orbital matrices out[B][K][N][N] (electron i, orbital mu) of a plain forward
  int orbitals(const void* r_, const void* R_, int Rb, int B, void* out, void* ws, int64_t wsb, cudaStream_t st) override {
    const T* r = (const T*)r_;
    const T* R = (const T*)R_;
    const int Bc = max_chunk(wsb, 1, B);
    if (Bc < 1) { err = "workspace too small for a single walker"; return 3; }
    int rc = 0;
    for (int b0 = 0; b0 < B && !rc; b0 += Bc) {
      const int nb = std::min(Bc, B - b0);
      mos_out = (T*)out + (size_t)b0 * K * N * N;
      rc = run_chunk(r + (size_t)b0 * 3 * N, R + (Rb ? (size_t)b0 * 3 * M : 0), Rb, nb, 1, B, nullptr, nullptr, nullptr, nullptr,
                     nullptr, ws, st);
    }
    mos_out = nullptr;
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  int forward(const void* r, const void* R, int Rb, int B, void* sign, void* logp, void* ws, int64_t wsb,
              cudaStream_t st) override {
    int rc = run_batched((const T*)r, (const T*)R, Rb, B, 1, (T*)sign, (T*)logp, nullptr, nullptr, nullptr, ws, wsb, st);
    if (rc) return rc;
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  int local_energy(const void* r_, const void* R_, int Rb, int B, uint64_t seed, const void* twist, void* E,
                   void* stats, void* sign, void* logp, void* grad, void* ws, int64_t wsb, cudaStream_t st) override {
    const T* r = (const T*)r_;
    const T* R = (const T*)R_;
    ph_active = ph_on;
    int rc = run_batched(r, R, Rb, B, T3 + 2, (T*)sign, (T*)logp, (T*)E, (T*)stats, (T*)grad, ws, wsb, st);
    ph_active = false;
    if (rc) return rc;
    if (J > 0) {
      // non-local ECP: virtual walkers (12 quadrature points x electrons x ECP nuclei)
      const int64_t vper = (int64_t)J * N * 12;
      int64_t Be = std::min<int64_t>(B, ecp_group_cap());
      if (ecp_bytes(Be) > wsb) {  // largest walker group whose virtual walkers fit the workspace
        int64_t lo = 0, hi = Be;
        while (hi - lo > 1) {
          int64_t mid = (lo + hi) / 2;
          if (ecp_bytes(mid) <= wsb) lo = mid; else hi = mid;
        }
        Be = lo;
      }
      if (Be < 1) Be = 1;  // one walker's virtual walkers do not fit at once: the plain-forward pass chunks them
      if (ecp_prefix_bytes(Be) + (int64_t)chunk_bytes(1, 1) > wsb) { err = "workspace too small for the non-local ECP pass"; return 3; }
      for (int b0 = 0; b0 < B; b0 += (int)Be) {
        int nb = (int)std::min<int64_t>(Be, B - b0);
        int64_t V = (int64_t)nb * vper;
        char* p = (char*)ws;
        T* rv = (T*)p; p += align_up(sizeof(T) * V * 3 * N);
        T* sv = (T*)p; p += align_up(sizeof(T) * V);
        T* lv = (T*)p; p += align_up(sizeof(T) * V);
        T* envt = (T*)p; p += align_up(sizeof(T) * (size_t)nb * N * K * N);
        T* embt = (T*)p; p += align_up(sizeof(T) * (size_t)nb * N * d);
        note_hwm(p);
        const T* rb = r + (size_t)b0 * 3 * N;
        const T* Rbp = R + (Rb ? (size_t)b0 * 3 * M : 0);
        const T* tw = twist ? (const T*)twist + (size_t)b0 * J * N : nullptr;
        DQ_LAUNCH(ecp_points_kernel<T>, dim3(nb * J * N), dim3(64), 0, st, rb, Rbp, Rb, N, M, J, (const int*)d_nl_nuc, tw,
                  seed, (uint64_t)b0, rv);
        if (Rb) { err = "non-local ECP with per-walker nuclei is not supported"; return 2; }
        const bool use_table = slater_fwd2_ok && N <= 32 && !dry && !std::getenv("DQMC_ECP_ENV_TABLE_OFF");
        if (use_table) {  // the quadrature forwards take the unmoved electrons' envelopes from the base walkers' table
          DQ_LAUNCH(env_table_kernel<T>, dim3(nb), dim3(256), sizeof(T) * N * M, st, rb, R, N, M, cfg.n_up, K * N,
                    P("env.pi_up"), P("env.pi_dn"), P("env.zeta_up"), P("env.zeta_dn"), cfg.n_env_per_nuc > 1 ? cfg.n_env_per_nuc : 1,
                    envt);
          ecp_env = envt; ecp_vper = (int)vper;
          if (embed_fwd_ok && can_trunk(1) && !std::getenv("DQMC_ECP_EMB_TABLE_OFF")) {
            DQ_LAUNCH(embed_fwd_kernel<T>, dim3((nb * N + 31) / 32), dim3(256), embed_fwd_smem_bytes<T>(M, d), st, rb, R, 0, N, M,
                      cfg.n_up, 1, P("emb.w"), d, embt, nb * N, 32, 0LL, 0);
            ecp_emb = embt;
          }
        }
        rc = run_batched(rv, R, 0, (int)V, 1, sv, lv, nullptr, nullptr, nullptr, p, wsb - (p - (char*)ws), st);
        ecp_env = nullptr; ecp_emb = nullptr;
        if (rc) return rc;
        DQ_LAUNCH(ecp_accumulate_kernel<T>, dim3((nb + 3) / 4), dim3(128), 0, st, rb, Rbp, Rb, N, M, J,
                  (const int*)d_nl_nuc, (const T*)d_nl_params, cfg.ecp_nl_lmax_p1, cfg.ecp_nl_terms,
                  (const T*)sign + b0, (const T*)logp + b0, (const T*)sv, (const T*)lv, nb, B, (T*)E + b0,
                  (T*)stats + b0);
        if (dry) break;  // planning pass: the first group is the largest
      }
    }
    DQ_CHECK(cudaGetLastError());
    return 0;
  }

  int mcmc(void* r_, void* sign_, void* logp_, int32_t* age, void* tau_, const void* R_, int Rb, int B, int n_sub,
           double target, int max_age, uint64_t seed, uint64_t step0, uint64_t woff, const void* nn, const void* nu,
           void* stats_, void* ws, int64_t wsb, cudaStream_t st, double p_exchange, const int32_t* ex_flags,
           const int32_t* ex_idx) override {
    // p_exchange > 0 (or ex_flags given): some sub-steps are spin-exchange steps (OppositeSpinExchangeSampler,
    // electron_samplers.py:286-330): the whole batch swaps one up / down pair per walker, plain Metropolis acceptance without
    // max_age override and without step-size adaptation.  ex_flags[n_sub] (host) / ex_idx[n_sub][B][2] (device): injected.
    T* r = (T*)r_; T* sign = (T*)sign_; T* logp = (T*)logp_; T* tau = (T*)tau_; T* stats = (T*)stats_;
    const T* R = (const T*)R_;
    char* p = (char*)ws;
    T* rp = (T*)p; p += align_up(sizeof(T) * (size_t)B * 3 * N);
    T* sp = (T*)p; p += align_up(sizeof(T) * (size_t)B);
    T* lp = (T*)p; p += align_up(sizeof(T) * (size_t)B);
    int* cnt = (int*)p; p += 256;
    int64_t rest = wsb - (p - (char*)ws);
    note_hwm(p);
    if (rest < 0) { err = "workspace too small (proposal buffers)"; return 3; }
    DQ_CHECK(cudaMemsetAsync(cnt, 0, sizeof(int), st));
    const int ne = B * 3 * N;
    for (int s = 0; s < n_sub; ++s) {
      const T* nns = nn ? (const T*)nn + (size_t)s * ne : nullptr;
      const T* nus = nu ? (const T*)nu + (size_t)s * B : nullptr;
      bool exchange = false;
      if (ex_flags) exchange = ex_flags[s] != 0;
      else if (p_exchange > 0.0) {  // one decision per sub-step for the whole batch (as the reference's lax.cond on a scalar)
        uint32_t w4[4];
        Philox::gen(seed ^ 0xA0761D6478BD642Full, 0, step0 + (uint64_t)s, w4);
        exchange = Philox::u01(w4[0], w4[1]) < p_exchange;
      }
      if (exchange)
        DQ_LAUNCH(exchange_propose_kernel<T>, dim3((B + 127) / 128), dim3(128), 0, st, (const T*)r, rp,
                  ex_idx ? ex_idx + (size_t)s * 2 * B : (const int32_t*)nullptr, seed, step0 + (uint64_t)s, woff, cfg.n_up, N, B);
      else
      DQ_LAUNCH(propose_kernel<T>, dim3((ne / 2 + 1 + 127) / 128), dim3(128), 0, st, (const T*)r, rp, (const T*)tau, nns,
                seed, step0 + (uint64_t)s, woff * (uint64_t)(3 * N), ne);
      int rc = run_batched(rp, R, Rb, B, 1, sp, lp, nullptr, nullptr, nullptr, p, rest, st);
      if (rc) return rc;
      DQ_LAUNCH(accept_kernel<T>, dim3((B + 127) / 128), dim3(128), 0, st, r, (const T*)rp, sign, (const T*)sp, logp,
                (const T*)lp, age, nus, seed, step0 + (uint64_t)s, woff, exchange ? -1 : max_age, B, N, cnt);
      DQ_LAUNCH(tau_kernel<T>, dim3(1), dim3(32), 0, st, tau, cnt, B, exchange ? T(0) : (T)target, stats);
    }
    DQ_LAUNCH(sampler_stats_kernel<T>, dim3(1), dim3(256), 0, st, (const T*)r, (const T*)logp, (const int*)age,
              (const T*)tau, B, N, stats);
    DQ_CHECK(cudaGetLastError());
    return 0;
  }
};

}  // namespace dq

// ================================= C ABI ======================================================
struct dqmc_engine {
  dq::EngineBase* e;
};

#define DQ_NEED_DEVICE(h)                                                                        \
  do {                                                                                            \
    if ((h)->e->plan_only) {                                                                      \
      (h)->e->err = "plan-only engine (created with device < 0): no compute entry points";        \
      return 2;                                                                                   \
    }                                                                                             \
  } while (0)

extern "C" {

const char* dqmc_version(void) { return "dqmc_b200 0.1 (sm_100a)"; }

int dqmc_create(const dqmc_config* cfg, int device, dqmc_handle* out) {
  if (!cfg || !out) return 2;
  dq::EngineBase* e = nullptr;
  int rc = 0;
  const bool plan = device < 0;  // plan-only engine: workspace planning without a CUDA device (dqmc_debug_plan)
  if (cfg->dtype == DQMC_F64) {
    auto* x = new dq::Engine<double>();
    x->cfg = *cfg; x->device = device; x->dry = x->plan_only = plan; rc = x->init(); e = x;
  } else if (cfg->dtype == DQMC_F32) {
    auto* x = new dq::Engine<float>();
    x->cfg = *cfg; x->device = device; x->dry = x->plan_only = plan; rc = x->init(); e = x;
  } else {
    return 2;
  }
  if (rc) {
    std::fprintf(stderr, "dqmc_create failed: %s\n", e->err.c_str());
    delete e;
    return rc;
  }
  *out = new dqmc_engine{e};
  return 0;
}
int dqmc_destroy(dqmc_handle h) {
  if (!h) return 2;
  delete h->e;
  delete h;
  return 0;
}
const char* dqmc_last_error(dqmc_handle h) { return h ? h->e->err.c_str() : "null handle"; }
int dqmc_param_count(dqmc_handle h) { return h ? (int)h->e->entries.size() : -1; }
int64_t dqmc_param_total(dqmc_handle h) { return h ? h->e->total : -1; }
int dqmc_param_entry(dqmc_handle h, int idx, char* name, int name_len, int64_t* offset, int32_t* rows, int32_t* cols) {
  if (!h || idx < 0 || idx >= (int)h->e->entries.size()) return 2;
  auto& en = h->e->entries[idx];
  if (name && name_len > 0) {
    std::strncpy(name, en.name.c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  if (offset) *offset = en.offset;
  if (rows) *rows = en.rows;
  if (cols) *cols = en.cols;
  return 0;
}
int dqmc_set_params(dqmc_handle h, const double* host_params, int64_t n, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  return h->e->set_params(host_params, n, (cudaStream_t)stream);
}
int64_t dqmc_workspace_bytes(dqmc_handle h, int32_t n_walkers, int32_t mode) {
  return h ? h->e->ws_bytes(n_walkers, mode) : -1;
}
int64_t dqmc_workspace_bytes_min(dqmc_handle h, int32_t n_walkers, int32_t mode) {
  return h ? h->e->ws_bytes_min(n_walkers, mode) : -1;
}
int dqmc_wf_forward(dqmc_handle h, const void* r, const void* R, int32_t R_batched, int32_t n_walkers, void* out_sign,
                    void* out_log, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 0) { h->e->err = "negative walker count"; return 2; }
  if (n_walkers == 0) return 0;  // empty batch: nothing to evaluate
  return h->e->forward(r, R, R_batched, n_walkers, out_sign, out_log, workspace, workspace_bytes, (cudaStream_t)stream);
}
int dqmc_wf_orbitals(dqmc_handle h, const void* r, const void* R, int32_t R_batched, int32_t n_walkers, void* out_orbitals,
                     void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 0) { h->e->err = "negative walker count"; return 2; }
  if (n_walkers == 0) return 0;
  return h->e->orbitals(r, R, R_batched, n_walkers, out_orbitals, workspace, workspace_bytes, (cudaStream_t)stream);
}
int dqmc_local_energy(dqmc_handle h, const void* r, const void* R, int32_t R_batched, int32_t n_walkers, uint64_t seed,
                      const void* ecp_twist, void* out_E, void* out_stats, void* out_sign, void* out_log,
                      void* out_grad, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 0) { h->e->err = "negative walker count"; return 2; }
  if (n_walkers == 0) return 0;  // empty batch: nothing to evaluate
  return h->e->local_energy(r, R, R_batched, n_walkers, seed, ecp_twist, out_E, out_stats, out_sign, out_log, out_grad,
                            workspace, workspace_bytes, (cudaStream_t)stream);
}
int dqmc_mcmc_sweep(dqmc_handle h, void* r, void* sign, void* log, int32_t* age, void* tau, const void* R,
                    int32_t R_batched, int32_t n_walkers, int32_t n_sub, double target_acceptance, int32_t max_age,
                    uint64_t seed, uint64_t step0, uint64_t walker_offset, const void* noise_normal,
                    const void* noise_uniform, void* out_stats, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 1) { h->e->err = "the sampler needs at least one walker"; return 2; }
  return h->e->mcmc(r, sign, log, age, tau, R, R_batched, n_walkers, n_sub, target_acceptance, max_age, seed, step0,
                    walker_offset, noise_normal, noise_uniform, out_stats, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}
int dqmc_mcmc_sweep_exchange(dqmc_handle h, void* r, void* sign, void* log, int32_t* age, void* tau, const void* R,
                             int32_t R_batched, int32_t n_walkers, int32_t n_sub, double target_acceptance, int32_t max_age,
                             uint64_t seed, uint64_t step0, uint64_t walker_offset, const void* noise_normal,
                             const void* noise_uniform, double exchange_step_probability, const int32_t* exchange_flags,
                             const int32_t* exchange_idx, void* out_stats, void* workspace, int64_t workspace_bytes,
                             void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 1) { h->e->err = "the sampler needs at least one walker"; return 2; }
  if (h->e->cfg.n_up < 1 || h->e->cfg.n_down < 1) { h->e->err = "spin exchange needs electrons of both spins"; return 2; }
  return h->e->mcmc(r, sign, log, age, tau, R, R_batched, n_walkers, n_sub, target_acceptance, max_age, seed, step0,
                    walker_offset, noise_normal, noise_uniform, out_stats, workspace, workspace_bytes, (cudaStream_t)stream,
                    exchange_step_probability, exchange_flags, exchange_idx);
}
int dqmc_langevin_sweep(dqmc_handle h, void* r, void* sign, void* log, void* force, int32_t* age, void* tau, const void* R,
                        int32_t R_batched, int32_t n_walkers, int32_t n_sub, double target_acceptance, int32_t max_age,
                        uint64_t seed, uint64_t step0, uint64_t walker_offset, const void* noise_normal,
                        const void* noise_uniform, void* out_stats, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 1) { h->e->err = "the sampler needs at least one walker"; return 2; }
  return h->e->langevin(r, sign, log, force, age, tau, R, R_batched, n_walkers, n_sub, target_acceptance, max_age, seed, step0,
                        walker_offset, noise_normal, noise_uniform, out_stats, workspace, workspace_bytes, (cudaStream_t)stream);
}
int dqmc_wf_vjp_params(dqmc_handle h, const void* r, const void* R, int32_t R_batched, int32_t n_walkers, const void* weights,
                       void* out_sign, void* out_log, void* out_grad_params, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 0) { h->e->err = "negative walker count"; return 2; }
  return h->e->vjp_params(r, R, R_batched, n_walkers, weights, out_sign, out_log, out_grad_params, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}
int dqmc_set_pseudo_hamiltonian(dqmc_handle h, int32_t n_tab, int32_t n_grid, double r_max, const double* tables,
                                const int32_t* tab_of_nuc) {
  if (!h) return 2;
  if (!h->e->plan_only) cudaSetDevice(h->e->device);
  return h->e->set_ph(n_tab, n_grid, r_max, tables, tab_of_nuc);
}
int dqmc_debug_plan(dqmc_handle h, int32_t n_walkers, int32_t mode, int64_t workspace_bytes, int64_t* planned_bytes,
                    int64_t* carved_bytes) {
  if (!h) return 2;
  if (n_walkers < 1) { h->e->err = "dqmc_debug_plan needs at least one walker"; return 2; }
  return h->e->debug_plan(n_walkers, mode, workspace_bytes, planned_bytes, carved_bytes);
}
int dqmc_stats_pack(dqmc_handle h, const void* E_loc, const void* stats, int32_t n_walkers, double* out11, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  if (n_walkers < 1 || !E_loc || !out11) { h->e->err = "dqmc_stats_pack: bad arguments"; return 2; }
  return h->e->stats_pack(E_loc, stats, n_walkers, out11, (cudaStream_t)stream);
}
int64_t dqmc_launch_count(dqmc_handle h) { return h ? h->e->launches : -1; }

int dqmc_debug_gemm(dqmc_handle h, const char* weight, const char* bias, const void* A, const void* Res, void* C,
                    int32_t rows, int32_t S, int32_t sliced, int32_t backend, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  return h->e->debug_gemm(weight, bias, A, Res, C, rows, S, sliced, backend, (cudaStream_t)stream);
}

int dqmc_debug_mlp_block(dqmc_handle h, int32_t layer, const void* O, const void* X, void* Out, int32_t rows, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  return h->e->debug_mlp_block(layer, O, X, Out, rows, (cudaStream_t)stream);
}
int dqmc_debug_trunk(dqmc_handle h, const void* X0, void* Out, int32_t rows, void* stream) {
  if (!h) return 2;
  DQ_NEED_DEVICE(h);
  return h->e->debug_trunk(X0, Out, rows, (cudaStream_t)stream);
}

int dqmc_profile_begin(dqmc_handle h) {
  if (!h) return 2;
  h->e->prof = true; h->e->prof_flops = 0; h->e->prof_n = 0;
  for (int c = 0; c < 3; ++c) { h->e->prof_cls_flops[c] = 0; h->e->prof_cls_n[c] = 0; }
  h->e->prof_cls.clear();
  return 0;
}
int dqmc_profile_end_classes(dqmc_handle h, double* ms3, double* flops3, int64_t* n3) {
  if (!h) return 2;
  double ms[3] = {0, 0, 0};
#ifndef DQMC_EMU
  for (size_t i = 0; i + 1 < h->e->prof_ev.size(); i += 2) {
    cudaEventSynchronize(h->e->prof_ev[i + 1]);
    float t = 0;
    cudaEventElapsedTime(&t, h->e->prof_ev[i], h->e->prof_ev[i + 1]);
    ms[h->e->prof_cls[i / 2]] += t;
    cudaEventDestroy(h->e->prof_ev[i]); cudaEventDestroy(h->e->prof_ev[i + 1]);
  }
  h->e->prof_ev.clear();
#endif
  h->e->prof_cls.clear();
  h->e->prof = false;
  for (int c = 0; c < 3; ++c) {
    if (ms3) ms3[c] = ms[c];
    if (flops3) flops3[c] = h->e->prof_cls_flops[c];
    if (n3) n3[c] = h->e->prof_cls_n[c];
  }
  return 0;
}
int dqmc_profile_end(dqmc_handle h, double* gemm_ms, double* gemm_flops, int64_t* n_gemm) {
  double ms[3], fl[3];
  int64_t n[3];
  const int rc = dqmc_profile_end_classes(h, ms, fl, n);
  if (rc) return rc;
  if (gemm_ms) *gemm_ms = ms[0] + ms[1] + ms[2];
  if (gemm_flops) *gemm_flops = fl[0] + fl[1] + fl[2];
  if (n_gemm) *n_gemm = n[0] + n[1] + n[2];
  return 0;
}

}  // extern "C"
