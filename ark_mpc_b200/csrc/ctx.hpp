// Internal: the context object and the helpers shared by the translation units behind the C ABI
// (arkmpc_b200.cu: scalar-field gates and the host-buffer path; arkmpc_curve.cu: point gates).
// Not installed; not part of the ABI.
#pragma once
#include "../../include/arkmpc_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "fr_kernels.cuh"

namespace arkctx {
constexpr int kSlots = 3;                 // chunk pipeline depth of the host-buffer path
constexpr size_t kChunkElems = 1u << 18;  // elements per staged chunk (tools/e2e_sweep.sh: 2^16 17.6 ms, 2^18 16.6 ms per 2^20-gate step)
constexpr int kMaxPartialBlocks = 1024;
constexpr int kNumCurves = 2;
}  // namespace arkctx

struct arkmpc_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // current (own or caller's)
  cudaStream_t slot_stream[arkctx::kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t slot_event[arkctx::kSlots] = {nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  char* partials = nullptr;  // 2 * kMaxPartialBlocks field elements (also reused for point partial sums)
  int* flag_dev = nullptr;
  int* flag_host = nullptr;  // pinned
  bool use_tma = false;      // ARKMPC_RECOMBINE=tma
  bool pdl = true;           // Beaver K1/K2 launched with programmatic stream serialization (ARKMPC_PDL=0 disables)
  bool full_grids = true;    // element-wise kernels: one element per thread instead of a persistent wave (ARKMPC_GRID=persistent reverts)
  size_t chunk_elems = arkctx::kChunkElems;  // host-buffer path staging granularity (ARKMPC_CHUNK_LOG2 overrides)
  void* gtab[arkctx::kNumCurves] = {nullptr, nullptr};  // fixed-base tables, built on first use (arkmpc_curve.cu)
  std::mutex gtab_mutex;
  void* ntt_tw = nullptr;    // twiddle table + constants of the last (field, log2n, direction) transform (arkmpc_ntt.cu)
  long ntt_key = -1;
  std::string last_error;
};

namespace arkctx {

inline int fail(arkmpc_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}

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

#define ARK_CHECK_CTX(ctx)                                                                                                      \
  do {                                                                                                                          \
    if (!(ctx)) return ARKMPC_ERR_INVALID;                                                                                      \
    cudaError_t _e = cudaSetDevice((ctx)->device);                                                                              \
    if (_e != cudaSuccess) return arkctx::fail(ctx, ARKMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_e));  \
  } while (0)

}  // namespace arkctx
