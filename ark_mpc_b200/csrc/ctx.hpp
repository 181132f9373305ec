// Internal: the context object and the helpers shared by the translation units behind the C ABI
// (arkmpc_b200.cu: scalar-field gates and the host-buffer path; arkmpc_curve.cu: point gates).
// Not installed; not part of the ABI.
#pragma once
#include "../../include/arkmpc_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "ctab.hpp"
#include "fr_kernels.cuh"

namespace arkctx {
constexpr int kSlots = 3;                 // chunk pipeline depth of the host-buffer path
constexpr size_t kChunkElems = 1u << 19;  // elements per staged chunk (per 2^20-gate two-party step: 2^17 17.1 ms, 2^18 16.5 ms, 2^19 16.2 ms, 2^20 16.6 ms; profiles/r02z18_summary.txt)
constexpr int kMaxPartialBlocks = 1024;
constexpr int kNumCurves = 2;
}  // namespace arkctx

// Threading (gate closures run on the executor thread or on arbitrary rayon workers concurrently,
// /root/reference/online-phase/src/fabric/executor/multi_threaded/executor.rs:208-217, and handles are cloned across tokio
// tasks, fabric/result.rs:262-266): a context may be called from any number of threads at once.  Every entry point holds the
// context's (recursive) lock for the duration of the call — calls only enqueue work, so the lock is short except for the few
// that synchronise (sum_is_zero, the host-buffer path) — and all work goes to the ONE context stream, so the order in which
// calls return is the order in which the GPU runs them: a gate whose inputs were produced by calls that have RETURNED on
// other threads sees their results without any event.  The last error string is per thread.
struct arkmpc_ctx {
  std::recursive_mutex mu;
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // current (own or caller's)
  cudaStream_t slot_stream[arkctx::kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t slot_event[arkctx::kSlots] = {nullptr, nullptr, nullptr};
  std::atomic<uint64_t> launches{0};
  char* partials = nullptr;  // 2 * kMaxPartialBlocks field elements (also reused for point partial sums)
  int* flag_dev = nullptr;
  int* flag_host = nullptr;  // pinned
  bool use_tma = false;      // ARKMPC_RECOMBINE=tma
  bool hint_independent = false;  // arkmpc_ctx_hint_independent: consumed by the next Beaver kernel launch
  bool l2_keep = true;       // K1 stores its masks evict-last in L2 (ARKMPC_L2_KEEP=0 disables)
  bool pdl = true;           // Beaver K1/K2 launched with programmatic stream serialization (ARKMPC_PDL=0 disables)
  bool full_grids = true;    // element-wise kernels: one element per thread instead of a persistent wave (ARKMPC_GRID=persistent reverts)
  int xy_mode = 2;           // host-buffer path, how x.share / y.share reach the mask kernel: 0 flat AoS copy, 1 strided DMA copy, 2 zero-copy reads of pinned memory (ARKMPC_XY=flat|2d|zc)
  size_t chunk_elems = arkctx::kChunkElems;  // host-buffer path staging granularity (ARKMPC_CHUNK_LOG2 overrides)
  void* gtab[arkctx::kNumCurves] = {nullptr, nullptr};  // fixed-base tables, built on first use (arkmpc_curve.cu)
  std::mutex gtab_mutex;
  void* tab_scratch = nullptr;  // window-table records of the variable-base multiplications (curve_kernels.cuh), L2-resident
  void* tab_masks = nullptr;    // per-SM slot claim masks
  void* ntt_tw[2] = {nullptr, nullptr};  // twiddle table + constants of the last forward [0] and inverse [1] transform (arkmpc_ntt.cu):
  cudaStream_t ntt_stream[2] = {nullptr, nullptr};  // the stream each table was built on
  long ntt_key[2] = {-1, -1};            // FFT -> batch_mul -> IFFT (authenticated_poly.rs:377-401) alternates directions without rebuilding
  void* nccl = nullptr;      // ncclComm_t of arkmpc_nccl_init (arkmpc_comm.cu)
  void* stage_ring = nullptr;  // pinned staging ring of arkmpc_memcpy_h2d (arkmpc_mem.cu), borrowed from the device's pool on first use
  int nccl_world = 0, nccl_rank = -1;
};

namespace arkctx {

// arkmpc_mem.cu: the device-memory cache records events on every live context's current stream, so contexts register with it
// and change `stream` under its lock
void mem_register(arkmpc_ctx* ctx);
void mem_unregister(arkmpc_ctx* ctx);
void mem_set_stream(arkmpc_ctx* ctx, cudaStream_t s);

inline std::string& thread_error() {
  static thread_local std::string e;
  return e;
}
inline int fail(arkmpc_ctx*, int code, const std::string& msg) {
  thread_error() = msg;
  return code;
}

// Held for the duration of one ABI call: the context lock, and the calling thread's current device switched to the
// context's device and RESTORED on return (a multi-GPU host thread keeps its own current device).
struct CallGuard {
  std::unique_lock<std::recursive_mutex> lk;
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit CallGuard(arkmpc_ctx* ctx) : lk(ctx->mu) {
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != ctx->device) err = cudaSetDevice(ctx->device); else prev = -1;
  }
  ~CallGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  CallGuard(const CallGuard&) = delete;
  CallGuard& operator=(const CallGuard&) = delete;
};

#define ARK_CUDA(ctx, expr)                                                                               \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      return arkctx::fail(ctx, _e == cudaErrorMemoryAllocation ? ARKMPC_ERR_OOM : ARKMPC_ERR_CUDA,        \
                          std::string(#expr) + ": " + cudaGetErrorString(_e));                            \
    }                                                                                                     \
  } while (0)

#define ARK_REQUIRE(ctx, cond, msg)                                   \
  do {                                                                \
    if (!(cond)) return arkctx::fail(ctx, ARKMPC_ERR_INVALID, msg);   \
  } while (0)

inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

inline ark::Vec vec(const void* p, uint32_t stride = 32) { return ark::Vec{static_cast<const char*>(p), stride}; }
inline ark::MVec mvec(void* p, uint32_t stride = 32) { return ark::MVec{static_cast<char*>(p), stride}; }

inline ark::fe8 load_host_fe(const uint64_t* h) {
  ark::fe8 r;
  for (int j = 0; j < 4; j++) {
    r.v[2 * j] = (uint32_t)h[j];
    r.v[2 * j + 1] = (uint32_t)(h[j] >> 32);
  }
  return r;
}

// constant-multiplier table of a batch-constant scalar given as the reference's 4 x u64 Montgomery image
template <class F>
inline ark::CTab host_ctab(const uint64_t* h) {
  ark::CTab t;
  ark::ctab_build<F>(t, h);
  return t;
}

// persistent grid: enough blocks of `block` threads to cover n, capped at blocks_per_sm resident blocks per SM
inline unsigned grid_for(const arkmpc_ctx* ctx, size_t n, int blocks_per_sm, int block = ark::kBlock) {
  size_t need = (n + block - 1) / block;
  size_t cap = (size_t)ctx->sm_count * blocks_per_sm;
  return (unsigned)(need < cap ? (need ? need : 1) : cap);
}
// Streaming element-wise kernels: one element per thread by default (the hardware block scheduler refills SMs as blocks
// retire; measured faster than a persistent wave for every kernel with a multiplication, profiles/r01e_bench_extra.txt); the
// grid-stride loops in the kernels still cover any n.
inline unsigned grid_stream(const arkmpc_ctx* ctx, size_t n, int blocks_per_sm, int block = ark::kBlock) {
  if (!ctx->full_grids) return grid_for(ctx, n, blocks_per_sm, block);
  size_t need = (n + block - 1) / block;
  return (unsigned)(need < (1u << 30) ? (need ? need : 1) : (1u << 30));
}

// consume the independence hint (one launch)
inline int take_hint(arkmpc_ctx* ctx) {
  const bool h = ctx->hint_independent && ctx->pdl;
  ctx->hint_independent = false;
  return h ? 1 : 0;
}

inline int post_launch(arkmpc_ctx* ctx, const char* what) {
  ctx->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, ARKMPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return ARKMPC_OK;
}

#define ARK_FIELD_SWITCH(ctx, field, ...)                                        \
  switch (field) {                                                               \
    case ARKMPC_BN254_FR: { using F = ark::Bn254Fr; __VA_ARGS__; } break;        \
    case ARKMPC_CURVE25519_FR: { using F = ark::Curve25519Fr; __VA_ARGS__; } break; \
    default: return arkctx::fail(ctx, ARKMPC_ERR_INVALID, "unknown field id");   \
  }

// First statement of every entry point (function scope: the guard lives until the call returns).
#define ARK_CHECK_CTX(ctx)                                                                                                      \
  if (!(ctx)) return ARKMPC_ERR_INVALID;                                                                                        \
  arkctx::CallGuard _ark_guard(ctx);                                                                                            \
  if (_ark_guard.err != cudaSuccess)                                                                                            \
    return arkctx::fail(ctx, ARKMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_ark_guard.err))

}  // namespace arkctx
