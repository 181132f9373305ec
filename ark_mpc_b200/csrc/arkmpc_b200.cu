// C ABI of libarkmpc_b200 (declared in include/arkmpc_b200.h): contexts, memory, launch wrappers.
// No torch types, no exceptions across the boundary.  Kernels live in fr_kernels.cuh / beaver.cuh / fp256.cuh.
#include "ctx.hpp"

using namespace ark;
using namespace arkctx;

namespace {

// one gate per thread (the grid-stride loop in the kernels still covers n beyond the grid limit)
inline unsigned full_grid(size_t n) {
  size_t need = (n + kBlock - 1) / kBlock;
  return (unsigned)(need < (1u << 30) ? (need ? need : 1) : (1u << 30));
}

// Launch with the programmatic-stream-serialization attribute (see pdl_prologue in fr_kernels.cuh) unless disabled.
template <class... KArgs, class... Args>
void launch_pdl(const arkmpc_ctx* ctx, void (*kern)(KArgs...), unsigned grid, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kBlock);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);  // errors surface through cudaGetLastError in post_launch
}

// ---- launch helpers shared by the device-pointer ABI and the host-buffer path ----
template <class F>
int launch_mask(arkmpc_ctx* ctx, cudaStream_t s, size_t n, Vec x, Vec y, Vec a, Vec b, MVec d, MVec e) {
  launch_pdl(ctx, beaver_mask_kernel<F>, grid_stream(ctx, n, 8), s, n, x, y, a, b, d, e, take_hint(ctx), ctx->l2_keep ? 1 : 0);
  return post_launch(ctx, "beaver_mask_kernel");
}

template <class F, int PARTY, bool OPEN>
int launch_recombine_tma(arkmpc_ctx* ctx, cudaStream_t s, size_t n, const RecombineArgs& g) {
  auto kern = beaver_recombine_tma_kernel<F, PARTY, OPEN>;
  static thread_local int configured_dev = -1;
  if (configured_dev != ctx->device) {  // per (instantiation, thread): opt in to > 48 KiB of dynamic shared memory
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
    if (e != cudaSuccess) return fail(ctx, ARKMPC_ERR_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured_dev = ctx->device;
  }
  const size_t tiles = (n + kTmaTile - 1) / kTmaTile;
  const size_t need = (tiles + kTmaWarps - 1) / kTmaWarps;
  const size_t cap = (size_t)ctx->sm_count * 2;
  kern<<<(unsigned)(need < cap ? need : cap), kBlock, kTmaSmemBytes, s>>>(n, g);
  return post_launch(ctx, "beaver_recombine_tma_kernel");
}

inline bool recombine_is_planar(const RecombineArgs& g, bool open) {
  const uint32_t st[] = {g.d_mine.stride, g.e_mine.stride, g.d_peer.stride, g.e_peer.stride, g.a_s.stride, g.a_m.stride,
                         g.b_s.stride,    g.b_m.stride,    g.c_s.stride,    g.c_m.stride};
  for (uint32_t v : st)
    if (v != 32) return false;
  (void)open;
  return true;
}

// Default: the LDG kernel.  Measured on B200 (profiles/r01b_ab.txt) the recombination is bound by integer issue
// (536 IMAD.WIDE at ~4 issue cycles + ~320 other instructions per gate), not by memory latency, so TMA staging buys
// nothing (88.4 us vs 84.5 us per 2^20 gates); ARKMPC_RECOMBINE=tma selects the staged kernel for planar operands.
template <class F>
int launch_recombine(arkmpc_ctx* ctx, cudaStream_t s, int party, size_t n, RecombineArgs g, bool open) {
  g.independent = take_hint(ctx);
  if (ctx->use_tma && recombine_is_planar(g, open)) {
    if (party == 0) return open ? launch_recombine_tma<F, 0, true>(ctx, s, n, g) : launch_recombine_tma<F, 0, false>(ctx, s, n, g);
    return open ? launch_recombine_tma<F, 1, true>(ctx, s, n, g) : launch_recombine_tma<F, 1, false>(ctx, s, n, g);
  }
  const unsigned grid = full_grid(n);
  if (party == 0) {
    if (open) launch_pdl(ctx, beaver_recombine_kernel<F, 0, true>, grid, s, n, g);
    else launch_pdl(ctx, beaver_recombine_kernel<F, 0, false>, grid, s, n, g);
  } else {
    if (open) launch_pdl(ctx, beaver_recombine_kernel<F, 1, true>, grid, s, n, g);
    else launch_pdl(ctx, beaver_recombine_kernel<F, 1, false>, grid, s, n, g);
  }
  return post_launch(ctx, "beaver_recombine_kernel");
}

}  // namespace

extern "C" {

int arkmpc_abi_version(void) { return ARKMPC_ABI_VERSION; }

const char* arkmpc_status_string(int status) {
  switch (status) {
    case ARKMPC_OK: return "ok";
    case ARKMPC_ERR_INVALID: return "invalid argument";
    case ARKMPC_ERR_CUDA: return "CUDA error";
    case ARKMPC_ERR_NO_DEVICE: return "no usable CUDA device";
    case ARKMPC_ERR_OOM: return "out of device memory";
    case ARKMPC_ERR_UNSUPPORTED: return "unsupported in this build";
    case ARKMPC_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
  }
}

int arkmpc_device_count(int* count) {
  if (!count) return ARKMPC_ERR_INVALID;
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return ARKMPC_ERR_NO_DEVICE; }
  *count = c;
  return ARKMPC_OK;
}

int arkmpc_ctx_create(int device, arkmpc_ctx** out) {
  if (!out) return ARKMPC_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return ARKMPC_ERR_NO_DEVICE; }
  if (device < 0 || device >= count) return ARKMPC_ERR_INVALID;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ARKMPC_ERR_CUDA;
  if (prop.major != 10) return ARKMPC_ERR_NO_DEVICE;  // the library carries sm_100a SASS only
  arkmpc_ctx* ctx = new (std::nothrow) arkmpc_ctx();
  if (!ctx) return ARKMPC_ERR_OOM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  CallGuard guard(ctx);  // switches to `device`, restores the caller's current device on return
  bool ok = guard.err == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; ok && i < kSlots; i++) {
    ok = cudaStreamCreateWithFlags(&ctx->slot_stream[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->slot_event[i], cudaEventDisableTiming) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&ctx->partials, 2 * kMaxPartialBlocks * 32) == cudaSuccess;
  ok = ok && cudaMalloc(&ctx->flag_dev, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMallocHost(&ctx->flag_host, sizeof(int)) == cudaSuccess;
  if (ok) {
    // keep stream-ordered allocations cached between host-path sessions
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  if (!ok) {
    cudaGetLastError();
    guard.lk.unlock();
    arkmpc_ctx_destroy(ctx);
    return ARKMPC_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  {
    const char* v = getenv("ARKMPC_RECOMBINE");
    ctx->use_tma = v && strcmp(v, "tma") == 0;
    const char* l2 = getenv("ARKMPC_L2_KEEP");
    if (l2) ctx->l2_keep = strcmp(l2, "0") != 0;
    const char* pd = getenv("ARKMPC_PDL");
    if (pd && strcmp(pd, "0") == 0) ctx->pdl = false;
    const char* gm = getenv("ARKMPC_GRID");
    if (gm && strcmp(gm, "persistent") == 0) ctx->full_grids = false;
    // Default: copy the whole AoS images.  Measured on B200 / PCIe 5 (tools/_zc.cu, profiles/r02b_zero_copy.txt): device reads of
    // pinned host memory move whole 64-byte blocks, so fetching only the 32-byte share halves at stride 64 takes exactly as long
    // as reading everything (25.5 GB/s useful = 51 GB/s on the wire), and a strided DMA copy is slower still (22 GB/s useful).
    ctx->xy_mode = 0;
    const char* xy = getenv("ARKMPC_XY");
    if (xy && strcmp(xy, "flat") == 0) ctx->xy_mode = 0;
    if (xy && strcmp(xy, "2d") == 0) ctx->xy_mode = 1;
    if (xy && strcmp(xy, "zc") == 0) ctx->xy_mode = 2;
    const char* c = getenv("ARKMPC_CHUNK_LOG2");
    if (c && atoi(c) >= 10 && atoi(c) <= 24) ctx->chunk_elems = (size_t)1 << atoi(c);
  }
  mem_register(ctx);
  *out = ctx;
  return ARKMPC_OK;
}

int arkmpc_ctx_destroy(arkmpc_ctx* ctx) {
  if (!ctx) return ARKMPC_OK;
  {
  CallGuard guard(ctx);
  // what this context still runs finishes here; only then does the memory cache stop ordering reuse after its stream
  if (ctx->stream && ctx->stream != ctx->own_stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  mem_unregister(ctx);
  arkmpc_nccl_destroy(ctx);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  for (int i = 0; i < kSlots; i++) {
    if (ctx->slot_stream[i]) { cudaStreamSynchronize(ctx->slot_stream[i]); cudaStreamDestroy(ctx->slot_stream[i]); }
    if (ctx->slot_event[i]) cudaEventDestroy(ctx->slot_event[i]);
  }
  if (ctx->partials) cudaFree(ctx->partials);
  if (ctx->flag_dev) cudaFree(ctx->flag_dev);
  if (ctx->flag_host) cudaFreeHost(ctx->flag_host);
  for (int c = 0; c < kNumCurves; c++)
    if (ctx->gtab[c]) cudaFree(ctx->gtab[c]);
  for (int d = 0; d < 2; d++)
    if (ctx->ntt_tw[d]) cudaFree(ctx->ntt_tw[d]);
  if (ctx->tab_scratch) cudaFree(ctx->tab_scratch);
  if (ctx->tab_masks) cudaFree(ctx->tab_masks);
  }  // the guard (and the lock it holds) must be gone before the context is
  delete ctx;
  return ARKMPC_OK;
}

int arkmpc_ctx_set_stream(arkmpc_ctx* ctx, void* cuda_stream) {
  if (!ctx) return ARKMPC_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  mem_set_stream(ctx, static_cast<cudaStream_t>(cuda_stream));
  return ARKMPC_OK;
}
int arkmpc_ctx_hint_independent(arkmpc_ctx* ctx) {
  if (!ctx) return ARKMPC_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  ctx->hint_independent = true;
  return ARKMPC_OK;
}
int arkmpc_ctx_reset_stream(arkmpc_ctx* ctx) {
  if (!ctx) return ARKMPC_ERR_INVALID;
  std::lock_guard<std::recursive_mutex> lk(ctx->mu);
  mem_set_stream(ctx, ctx->own_stream);
  return ARKMPC_OK;
}
void* arkmpc_ctx_get_stream(arkmpc_ctx* ctx) { return ctx ? ctx->stream : nullptr; }
int arkmpc_ctx_device(arkmpc_ctx* ctx) { return ctx ? ctx->device : -1; }
int arkmpc_ctx_sm_count(arkmpc_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
uint64_t arkmpc_ctx_launch_count(arkmpc_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
const char* arkmpc_last_error(arkmpc_ctx* ctx) { return ctx ? thread_error().c_str() : "null context"; }

int arkmpc_ctx_sync(arkmpc_ctx* ctx) {
  ARK_CHECK_CTX(ctx);
  ARK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ARKMPC_OK;
}

// ---- memory ---- (arkmpc_malloc / arkmpc_free / arkmpc_host_alloc / arkmpc_host_free / arkmpc_memcpy_h2d: arkmpc_mem.cu)
int arkmpc_memcpy_d2h(arkmpc_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  ARK_CHECK_CTX(ctx);
  if (bytes == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, dst_host && src_dev, "null pointer");
  ARK_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return ARKMPC_OK;
}
int arkmpc_memcpy_d2d(arkmpc_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes) {
  ARK_CHECK_CTX(ctx);
  if (bytes == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, dst_dev && src_dev, "null pointer");
  ARK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return ARKMPC_OK;
}

// ---- layout ----
int arkmpc_share_unzip(arkmpc_ctx* ctx, size_t n, const uint64_t* aos_dev, uint64_t* share_plane, uint64_t* mac_plane) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, aos_dev && share_plane && mac_plane, "null pointer");
  ARK_REQUIRE(ctx, aligned32(aos_dev) && aligned32(share_plane) && aligned32(mac_plane), "planes must be 32-byte aligned");
  const char* base = reinterpret_cast<const char*>(aos_dev);
  copy_planes_kernel<<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(base, 64), vec(base + 32, 64), mvec(share_plane), mvec(mac_plane));
  return post_launch(ctx, "share_unzip");
}
int arkmpc_share_zip(arkmpc_ctx* ctx, size_t n, const uint64_t* share_plane, const uint64_t* mac_plane, uint64_t* aos_dev) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, aos_dev && share_plane && mac_plane, "null pointer");
  ARK_REQUIRE(ctx, aligned32(aos_dev) && aligned32(share_plane) && aligned32(mac_plane), "planes must be 32-byte aligned");
  char* base = reinterpret_cast<char*>(aos_dev);
  copy_planes_kernel<<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(share_plane), vec(mac_plane), mvec(base, 64), mvec(base + 32, 64));
  return post_launch(ctx, "share_zip");
}

// ---- Beaver multiplication ----
int arkmpc_fr_beaver_mask(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* x_share, const uint64_t* y_share,
                          const uint64_t* a_share, const uint64_t* b_share, uint64_t* d_mine, uint64_t* e_mine) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, x_share && y_share && a_share && b_share && d_mine && e_mine, "null pointer");
  ARK_REQUIRE(ctx, aligned32(x_share) && aligned32(y_share) && aligned32(a_share) && aligned32(b_share) && aligned32(d_mine) && aligned32(e_mine),
              "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, return launch_mask<F>(ctx, ctx->stream, n, vec(x_share), vec(y_share), vec(a_share), vec(b_share), mvec(d_mine), mvec(e_mine)));
  return ARKMPC_OK;
}

int arkmpc_fr_beaver_recombine(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* d_mine, const uint64_t* e_mine, const uint64_t* d_peer, const uint64_t* e_peer,
                               const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac,
                               const uint64_t* c_share, const uint64_t* c_mac, uint64_t* out_share, uint64_t* out_mac,
                               uint64_t* d_open, uint64_t* e_open) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, key_host && d_mine && e_mine && d_peer && e_peer && a_share && a_mac && b_share && b_mac && c_share && c_mac && out_share && out_mac,
              "null pointer");
  ARK_REQUIRE(ctx, (d_open == nullptr) == (e_open == nullptr), "d_open and e_open must both be given or both be NULL");
  const void* ptrs[] = {d_mine, e_mine, d_peer, e_peer, a_share, a_mac, b_share, b_mac, c_share, c_mac, out_share, out_mac, d_open, e_open};
  for (const void* p : ptrs) ARK_REQUIRE(ctx, aligned32(p), "planes must be 32-byte aligned");
  RecombineArgs g;
  g.d_mine = vec(d_mine); g.e_mine = vec(e_mine); g.d_peer = vec(d_peer); g.e_peer = vec(e_peer);
  g.a_s = vec(a_share); g.a_m = vec(a_mac); g.b_s = vec(b_share); g.b_m = vec(b_mac); g.c_s = vec(c_share); g.c_m = vec(c_mac);
  g.out_s = mvec(out_share); g.out_m = mvec(out_mac); g.d_open = mvec(d_open); g.e_open = mvec(e_open);
  ARK_FIELD_SWITCH(ctx, field, { g.key = host_ctab<F>(key_host); return launch_recombine<F>(ctx, ctx->stream, party_id, n, g, d_open != nullptr); });
  return ARKMPC_OK;
}

// ---- multi-GPU open gather ----
int arkmpc_ipc_export(arkmpc_ctx* ctx, void* dev_ptr, uint8_t* handle_out) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, dev_ptr && handle_out, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == ARKMPC_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  ARK_CUDA(ctx, cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, sizeof h);
  return ARKMPC_OK;
}
int arkmpc_ipc_import(arkmpc_ctx* ctx, const uint8_t* handle, void** peer_dev_ptr) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, handle && peer_dev_ptr, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  *peer_dev_ptr = nullptr;
  ARK_CUDA(ctx, cudaIpcOpenMemHandle(peer_dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ARKMPC_OK;
}
int arkmpc_ipc_release(arkmpc_ctx* ctx, void* peer_dev_ptr) {
  ARK_CHECK_CTX(ctx);
  if (peer_dev_ptr) ARK_CUDA(ctx, cudaIpcCloseMemHandle(peer_dev_ptr));
  return ARKMPC_OK;
}

int arkmpc_fr_beaver_recombine_gather(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                      const uint64_t* d_mine, const uint64_t* e_mine, const uint64_t* d_peer, const uint64_t* e_peer,
                                      const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac,
                                      const uint64_t* c_share, const uint64_t* c_mac, uint64_t* out_share, uint64_t* out_mac,
                                      int world, int rank, uint64_t* const* gather_d, uint64_t* const* gather_e) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank");
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, key_host && gather_d && gather_e, "null pointer");
  const void* ptrs[] = {d_mine, e_mine, d_peer, e_peer, a_share, a_mac, b_share, b_mac, c_share, c_mac, out_share, out_mac};
  for (const void* p : ptrs) ARK_REQUIRE(ctx, p && aligned32(p), "null or misaligned plane");
  RecombineArgs g;
  g.d_mine = vec(d_mine); g.e_mine = vec(e_mine); g.d_peer = vec(d_peer); g.e_peer = vec(e_peer);
  g.a_s = vec(a_share); g.a_m = vec(a_mac); g.b_s = vec(b_share); g.b_m = vec(b_mac); g.c_s = vec(c_share); g.c_m = vec(c_mac);
  g.out_s = mvec(out_share); g.out_m = mvec(out_mac); g.d_open = mvec(nullptr); g.e_open = mvec(nullptr);
  GatherArgs q;
  q.world = world;
  for (int k = 0; k < kMaxPeers; k++) {
    q.d[k] = q.e[k] = nullptr;
    if (k < world) {
      ARK_REQUIRE(ctx, gather_d[k] && gather_e[k] && aligned32(gather_d[k]) && aligned32(gather_e[k]), "null or misaligned gather plane");
      q.d[k] = reinterpret_cast<char*>(gather_d[k]) + (size_t)rank * n * 32;
      q.e[k] = reinterpret_cast<char*>(gather_e[k]) + (size_t)rank * n * 32;
    }
  }
  const unsigned grid = full_grid(n);
  ARK_FIELD_SWITCH(ctx, field, {
    g.key = host_ctab<F>(key_host);
    g.independent = take_hint(ctx);
    if (party_id == 0) launch_pdl(ctx, beaver_recombine_gather_kernel<F, 0>, grid, ctx->stream, n, g, q);
    else launch_pdl(ctx, beaver_recombine_gather_kernel<F, 1>, grid, ctx->stream, n, g, q);
  });
  return post_launch(ctx, "beaver_recombine_gather_kernel");
}

// ---- public-scalar vector gates ----
#define ARK_BINARY_ENTRY(NAME, OP)                                                                                         \
  int NAME(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {                   \
    ARK_CHECK_CTX(ctx);                                                                                                    \
    if (n == 0) return ARKMPC_OK;                                                                                          \
    ARK_REQUIRE(ctx, a && b && out, "null pointer");                                                                       \
    ARK_REQUIRE(ctx, aligned32(a) && aligned32(b) && aligned32(out), "planes must be 32-byte aligned");                    \
    ARK_FIELD_SWITCH(ctx, field, (fr_binary_kernel<F, OP><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(a), vec(b), mvec(out)))); \
    return post_launch(ctx, #NAME);                                                                                        \
  }
ARK_BINARY_ENTRY(arkmpc_fr_add, Bin::Add)
ARK_BINARY_ENTRY(arkmpc_fr_sub, Bin::Sub)
ARK_BINARY_ENTRY(arkmpc_fr_mul, Bin::Mul)
#undef ARK_BINARY_ENTRY

int arkmpc_fr_neg(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a && out, "null pointer");
  ARK_REQUIRE(ctx, aligned32(a) && aligned32(out), "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, (fr_neg_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(a), mvec(out))));
  return post_launch(ctx, "arkmpc_fr_neg");
}

int arkmpc_fr_scale(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* s_host, uint64_t* out) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a && out && s_host, "null pointer");
  ARK_REQUIRE(ctx, aligned32(a) && aligned32(out), "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, (fr_scale_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(a), host_ctab<F>(s_host), mvec(out))));
  return post_launch(ctx, "arkmpc_fr_scale");
}

int arkmpc_fr_to_mont(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* plain, uint64_t* mont) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, plain && mont, "null pointer");
  ARK_REQUIRE(ctx, aligned32(plain) && aligned32(mont), "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, {
    fe8 r2;
    Fp<F>::set_r2(r2);
    uint64_t r2h[4];
    for (int j = 0; j < 4; j++) r2h[j] = (uint64_t)r2.v[2 * j] | ((uint64_t)r2.v[2 * j + 1] << 32);
    fr_scale_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(plain), host_ctab<F>(r2h), mvec(mont));
  });
  return post_launch(ctx, "arkmpc_fr_to_mont");
}

int arkmpc_fr_from_mont(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* mont, uint64_t* plain) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, plain && mont, "null pointer");
  ARK_REQUIRE(ctx, aligned32(plain) && aligned32(mont), "planes must be 32-byte aligned");
  const uint64_t one[4] = {1, 0, 0, 0};
  ARK_FIELD_SWITCH(ctx, field, (fr_scale_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(mont), host_ctab<F>(one), mvec(plain))));
  return post_launch(ctx, "arkmpc_fr_from_mont");
}

// ---- linear gates on shares ----
#define ARK_SHARE_BINARY_ENTRY(NAME, OP)                                                                                   \
  int NAME(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* b_share, \
           const uint64_t* b_mac, uint64_t* out_share, uint64_t* out_mac) {                                               \
    ARK_CHECK_CTX(ctx);                                                                                                    \
    if (n == 0) return ARKMPC_OK;                                                                                          \
    ARK_REQUIRE(ctx, a_share && a_mac && b_share && b_mac && out_share && out_mac, "null pointer");                        \
    ARK_REQUIRE(ctx, aligned32(a_share) && aligned32(a_mac) && aligned32(b_share) && aligned32(b_mac) && aligned32(out_share) && aligned32(out_mac), \
                "planes must be 32-byte aligned");                                                                         \
    ARK_FIELD_SWITCH(ctx, field, (fr_share_binary_kernel<F, OP><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(          \
                                     n, vec(a_share), vec(a_mac), vec(b_share), vec(b_mac), mvec(out_share), mvec(out_mac)))); \
    return post_launch(ctx, #NAME);                                                                                        \
  }
ARK_SHARE_BINARY_ENTRY(arkmpc_fr_share_add, Bin::Add)
ARK_SHARE_BINARY_ENTRY(arkmpc_fr_share_sub, Bin::Sub)
#undef ARK_SHARE_BINARY_ENTRY

int arkmpc_fr_share_neg(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                        uint64_t* out_share, uint64_t* out_mac) {
  int rc = arkmpc_fr_neg(ctx, field, n, a_share, out_share);
  if (rc != ARKMPC_OK) return rc;
  return arkmpc_fr_neg(ctx, field, n, a_mac, out_mac);
}

static int share_add_public_impl(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n, const uint64_t* a_share,
                                 const uint64_t* a_mac, const uint64_t* v, uint64_t* out_share, uint64_t* out_mac, bool sub) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, key_host && a_share && a_mac && v && out_share && out_mac, "null pointer");
  ARK_REQUIRE(ctx, aligned32(a_share) && aligned32(a_mac) && aligned32(v) && aligned32(out_share) && aligned32(out_mac), "planes must be 32-byte aligned");
  const unsigned grid = grid_stream(ctx, n, 8);
  cudaStream_t s = ctx->stream;
  ARK_FIELD_SWITCH(ctx, field, {
    const CTab key = host_ctab<F>(key_host);
    if (party_id == 0) {
      if (sub) fr_share_add_public_kernel<F, 0, true><<<grid, kBlock, 0, s>>>(n, vec(a_share), vec(a_mac), vec(v), key, mvec(out_share), mvec(out_mac));
      else fr_share_add_public_kernel<F, 0, false><<<grid, kBlock, 0, s>>>(n, vec(a_share), vec(a_mac), vec(v), key, mvec(out_share), mvec(out_mac));
    } else {
      if (sub) fr_share_add_public_kernel<F, 1, true><<<grid, kBlock, 0, s>>>(n, vec(a_share), vec(a_mac), vec(v), key, mvec(out_share), mvec(out_mac));
      else fr_share_add_public_kernel<F, 1, false><<<grid, kBlock, 0, s>>>(n, vec(a_share), vec(a_mac), vec(v), key, mvec(out_share), mvec(out_mac));
    }
  });
  return post_launch(ctx, "share_add_public");
}
int arkmpc_fr_share_add_public(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n, const uint64_t* a_share,
                               const uint64_t* a_mac, const uint64_t* v, uint64_t* out_share, uint64_t* out_mac) {
  return share_add_public_impl(ctx, field, party_id, key_host, n, a_share, a_mac, v, out_share, out_mac, false);
}
int arkmpc_fr_share_sub_public(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n, const uint64_t* a_share,
                               const uint64_t* a_mac, const uint64_t* v, uint64_t* out_share, uint64_t* out_mac) {
  return share_add_public_impl(ctx, field, party_id, key_host, n, a_share, a_mac, v, out_share, out_mac, true);
}

int arkmpc_fr_share_mul_public(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* v,
                               uint64_t* out_share, uint64_t* out_mac) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a_share && a_mac && v && out_share && out_mac, "null pointer");
  ARK_REQUIRE(ctx, aligned32(a_share) && aligned32(a_mac) && aligned32(v) && aligned32(out_share) && aligned32(out_mac), "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, (fr_share_mul_public_kernel<F><<<grid_stream(ctx, n, 4), kBlock, 0, ctx->stream>>>(
                                   n, vec(a_share), vec(a_mac), vec(v), mvec(out_share), mvec(out_mac))));
  return post_launch(ctx, "arkmpc_fr_share_mul_public");
}

// ---- open_authenticated pieces ----
int arkmpc_fr_mac_check(arkmpc_ctx* ctx, int field, const uint64_t* key_host, size_t n, const uint64_t* opened, const uint64_t* mac, uint64_t* check) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, key_host && opened && mac && check, "null pointer");
  ARK_REQUIRE(ctx, aligned32(opened) && aligned32(mac) && aligned32(check), "planes must be 32-byte aligned");
  ARK_FIELD_SWITCH(ctx, field, (fr_mac_check_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(opened), vec(mac), host_ctab<F>(key_host), mvec(check))));
  return post_launch(ctx, "arkmpc_fr_mac_check");
}

int arkmpc_fr_sum_is_zero(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* mine, const uint64_t* peer, int* all_zero_host) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, all_zero_host, "null pointer");
  *all_zero_host = 1;
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, mine && peer, "null pointer");
  ARK_REQUIRE(ctx, aligned32(mine) && aligned32(peer), "planes must be 32-byte aligned");
  *ctx->flag_host = 1;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_dev, ctx->flag_host, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ARK_FIELD_SWITCH(ctx, field, (fr_sum_is_zero_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(mine), vec(peer), ctx->flag_dev)));
  int rc = post_launch(ctx, "arkmpc_fr_sum_is_zero");
  if (rc != ARKMPC_OK) return rc;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ARK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *all_zero_host = *ctx->flag_host;
  return ARKMPC_OK;
}

int arkmpc_fr_validate(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, int* all_canonical_host) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, all_canonical_host, "null pointer");
  *all_canonical_host = 1;
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a && aligned32(a), "null or misaligned plane");
  *ctx->flag_host = 1;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_dev, ctx->flag_host, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  ARK_FIELD_SWITCH(ctx, field, (fr_validate_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(a), ctx->flag_dev)));
  int rc = post_launch(ctx, "arkmpc_fr_validate");
  if (rc != ARKMPC_OK) return rc;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ARK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *all_canonical_host = *ctx->flag_host;
  return ARKMPC_OK;
}

int arkmpc_fr_to_bytes_be(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint8_t* out_dev) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a && out_dev, "null pointer");
  ARK_REQUIRE(ctx, aligned32(a) && aligned32(out_dev), "planes must be 32-byte aligned");
  const uint64_t one[4] = {1, 0, 0, 0};
  ARK_FIELD_SWITCH(ctx, field, (fr_to_bytes_be_kernel<F><<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(a), host_ctab<F>(one), mvec(out_dev))));
  return post_launch(ctx, "arkmpc_fr_to_bytes_be");
}

// ---- sums ----
static int sum_impl(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a0, const uint64_t* a1, uint64_t* out0, uint64_t* out1) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, a0 && out0 && aligned32(a0) && aligned32(out0), "null or misaligned pointer");
  const bool two = a1 != nullptr;
  if (two) ARK_REQUIRE(ctx, out1 && aligned32(a1) && aligned32(out1), "null or misaligned pointer");
  unsigned grid = grid_for(ctx, n, 4);
  if (grid > kMaxPartialBlocks) grid = kMaxPartialBlocks;
  char* p0 = ctx->partials;
  char* p1 = ctx->partials + (size_t)kMaxPartialBlocks * 32;
  cudaStream_t s = ctx->stream;
  // n == 0 gives the additive identity, as an empty Rust `sum()` does
  ARK_FIELD_SWITCH(ctx, field, {
    if (two) {
      fr_sum_kernel<F, 2><<<grid, kBlock, 0, s>>>(n, vec(a0), vec(a1), mvec(p0), mvec(p1), true);
      fr_sum_kernel<F, 2><<<1, kBlock, 0, s>>>(grid, vec(p0), vec(p1), mvec(out0), mvec(out1), false);
    } else {
      fr_sum_kernel<F, 1><<<grid, kBlock, 0, s>>>(n, vec(a0), vec(a0), mvec(p0), mvec(p0), true);
      fr_sum_kernel<F, 1><<<1, kBlock, 0, s>>>(grid, vec(p0), vec(p0), mvec(out0), mvec(out0), false);
    }
  });
  ctx->launches++;
  return post_launch(ctx, "fr_sum");
}
int arkmpc_fr_beaver_recombine_sum(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n, const uint64_t* d_mine,
                                   const uint64_t* e_mine, const uint64_t* d_peer, const uint64_t* e_peer, const uint64_t* a_share,
                                   const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac, const uint64_t* c_share,
                                   const uint64_t* c_mac, uint64_t* out_share, uint64_t* out_mac) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, field == ARKMPC_BN254_FR || field == ARKMPC_CURVE25519_FR, "unknown field id");  // before any scratch is allocated
  ARK_REQUIRE(ctx, out_share && out_mac && aligned32(out_share) && aligned32(out_mac), "null or misaligned pointer");
  if (n == 0) return sum_impl(ctx, field, 0, out_share, out_mac, out_share, out_mac);  // the additive identity, like an empty sum()
  ARK_REQUIRE(ctx, key_host && d_mine && e_mine && d_peer && e_peer && a_share && a_mac && b_share && b_mac && c_share && c_mac, "null pointer");
  const void* ptrs[] = {d_mine, e_mine, d_peer, e_peer, a_share, a_mac, b_share, b_mac, c_share, c_mac};
  for (const void* p : ptrs) ARK_REQUIRE(ctx, aligned32(p), "planes must be 32-byte aligned");
  RecombineArgs g;
  g.d_mine = vec(d_mine); g.e_mine = vec(e_mine); g.d_peer = vec(d_peer); g.e_peer = vec(e_peer);
  g.a_s = vec(a_share); g.a_m = vec(a_mac); g.b_s = vec(b_share); g.b_m = vec(b_mac); g.c_s = vec(c_share); g.c_m = vec(c_mac);
  g.out_s = mvec(nullptr); g.out_m = mvec(nullptr); g.d_open = mvec(nullptr); g.e_open = mvec(nullptr);
  const size_t per_block = (size_t)kBlock * kSumGatesPerThread;
  const size_t need = (n + per_block - 1) / per_block;
  const unsigned grid = (unsigned)(need < (1u << 30) ? need : (1u << 30));
  const size_t warps = (size_t)grid * (kBlock / 32);
  cudaStream_t s = ctx->stream;
  char* part = nullptr;  // one (share, mac) partial per warp, stream-ordered scratch
  ARK_CUDA(ctx, cudaMallocAsync(&part, warps * 64, s));
  char* part_m = part + warps * 32;
  ARK_FIELD_SWITCH(ctx, field, {
    g.key = host_ctab<F>(key_host);
    g.independent = take_hint(ctx);
    if (party_id == 0) launch_pdl(ctx, beaver_recombine_sum_kernel<F, 0>, grid, s, n, g, mvec(part), mvec(part_m));
    else launch_pdl(ctx, beaver_recombine_sum_kernel<F, 1>, grid, s, n, g, mvec(part), mvec(part_m));
  });
  int rc = post_launch(ctx, "beaver_recombine_sum_kernel");
  if (rc == ARKMPC_OK)
    rc = sum_impl(ctx, field, warps, reinterpret_cast<const uint64_t*>(part), reinterpret_cast<const uint64_t*>(part_m), out_share, out_mac);
  cudaFreeAsync(part, s);
  return rc;
}

int arkmpc_fr_share_sum(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac, uint64_t* out_share, uint64_t* out_mac) {
  if (!a_mac) return fail(ctx, ARKMPC_ERR_INVALID, "null pointer");
  return sum_impl(ctx, field, n, a_share, a_mac, out_share, out_mac);
}
int arkmpc_fr_sum(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out) { return sum_impl(ctx, field, n, a, nullptr, out, nullptr); }

int arkmpc_fr_random(arkmpc_ctx* ctx, int field, uint64_t seed, uint64_t first_index, size_t n, uint64_t* out) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, out && aligned32(out), "null or misaligned pointer");
  ARK_FIELD_SWITCH(ctx, field, (fr_random_kernel<F><<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, seed, first_index, mvec(out))));
  return post_launch(ctx, "arkmpc_fr_random");
}

}  // extern "C"

// ================================================================================================
// Host-buffer path: both phases of batch_mul over the reference's AoS host images, chunked so that
// the H2D copy of chunk k+1, the kernel of chunk k and the D2H copy of chunk k-1 overlap.
// ================================================================================================
struct arkmpc_batch_mul {
  arkmpc_ctx* ctx;
  int field, party;
  uint64_t key[4];
  size_t n;
  char* abc = nullptr;    // a | b | c, AoS, n*64 B each (resident between begin and finish)
  char* de = nullptr;     // d_mine | e_mine planes, n*32 B each
  char* stage = nullptr;  // kSlots * chunk staging
  size_t chunk = 0;
  bool xy_zero_copy = false;
};

namespace {
size_t stage_bytes_per_slot(size_t chunk) { return chunk * 64 * 2 + chunk * 32 * 2; }  // begin: x,y AoS ; finish: d_peer,e_peer + out AoS + d/e open

// device-visible address of a pinned + mapped host buffer, or nullptr for pageable memory
const char* mapped_device_pointer(const void* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return static_cast<const char*>(at.devicePointer);
}

// the caller holds the context guard; drains the slot streams before the buffers go back to the pool
int session_free(arkmpc_batch_mul* s) {
  arkmpc_ctx* ctx = s->ctx;
  for (int i = 0; i < kSlots; i++) cudaStreamSynchronize(ctx->slot_stream[i]);
  if (s->abc) cudaFreeAsync(s->abc, ctx->stream);
  if (s->de) cudaFreeAsync(s->de, ctx->stream);
  if (s->stage) cudaFreeAsync(s->stage, ctx->stream);
  delete s;
  return ARKMPC_OK;
}

// xy_stride: 64 = x_host / y_host are AoS ScalarShare images (share || mac), 32 = planes of the share halves only
int begin_host_impl(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n, size_t xy_stride,
                    const uint64_t* x_host, const uint64_t* y_host, const uint64_t* a_host, const uint64_t* b_host,
                    const uint64_t* c_host, uint64_t* de_mine_host, arkmpc_batch_mul** session) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, session, "null session pointer");
  *session = nullptr;
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, field == ARKMPC_BN254_FR || field == ARKMPC_CURVE25519_FR, "unknown field id");
  ARK_REQUIRE(ctx, key_host, "null key");
  if (n > 0) ARK_REQUIRE(ctx, x_host && y_host && a_host && b_host && c_host && de_mine_host, "null pointer");
  arkmpc_batch_mul* s = new (std::nothrow) arkmpc_batch_mul();
  if (!s) return fail(ctx, ARKMPC_ERR_OOM, "session alloc");
  s->ctx = ctx; s->field = field; s->party = party_id; s->n = n;
  memcpy(s->key, key_host, 32);
  s->chunk = n < ctx->chunk_elems ? (n ? n : 1) : ctx->chunk_elems;
  *session = s;
  if (n == 0) return ARKMPC_OK;
  cudaStream_t cs = ctx->stream;
  cudaError_t e = cudaMallocAsync(&s->abc, n * 64 * 3, cs);
  if (e == cudaSuccess) e = cudaMallocAsync(&s->de, n * 32 * 2, cs);
  if (e == cudaSuccess) e = cudaMallocAsync(&s->stage, stage_bytes_per_slot(s->chunk) * kSlots, cs);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
  if (e != cudaSuccess) {
    session_free(s);
    *session = nullptr;
    return fail(ctx, e == cudaErrorMemoryAllocation ? ARKMPC_ERR_OOM : ARKMPC_ERR_CUDA, std::string("session buffers: ") + cudaGetErrorString(e));
  }
  const char* xh = reinterpret_cast<const char*>(x_host);
  const char* yh = reinterpret_cast<const char*>(y_host);
  const char* ah = reinterpret_cast<const char*>(a_host);
  const char* bh = reinterpret_cast<const char*>(b_host);
  const char* ch = reinterpret_cast<const char*>(c_host);
  char* dh = reinterpret_cast<char*>(de_mine_host);
  char* eh = dh + n * 32;
  char* a_dev = s->abc;
  char* b_dev = s->abc + n * 64;
  char* c_dev = s->abc + n * 128;
  char* d_dev = s->de;
  char* e_dev = s->de + n * 32;
  // zero-copy needs both buffers pinned and mapped (cudaHostAlloc / cudaHostRegister / arkmpc_host_alloc); pageable memory
  // falls back to the flat copy
  const char* x_map = ctx->xy_mode == 2 && xy_stride == 64 ? mapped_device_pointer(xh) : nullptr;
  const char* y_map = ctx->xy_mode == 2 && xy_stride == 64 ? mapped_device_pointer(yh) : nullptr;
  const bool xy_zero_copy = x_map && y_map;
  s->xy_zero_copy = xy_zero_copy;
  auto run = [&]() -> int {
    int slot = 0;
    for (size_t off = 0; off < n; off += s->chunk, slot = (slot + 1) % kSlots) {
      const size_t m = (n - off < s->chunk) ? n - off : s->chunk;
      cudaStream_t st = ctx->slot_stream[slot];
      char* xs = s->stage + (size_t)slot * stage_bytes_per_slot(s->chunk);
      char* ys = xs + s->chunk * 64;
      // Only the share halves of x and y are read by the mask kernel (open_batch sends share.share() only,
      // authenticated_scalar.rs:141-145).  xy_mode (ARKMPC_XY): "zc" = the kernel reads them straight from the caller's
      // pinned buffers over PCIe with 256-bit loads at stride 64, the MAC halves never cross the link; "2d" = a strided
      // DMA copy; "flat" = copy the whole AoS image.
      Vec xv, yv;
      if (xy_zero_copy) {
        xv = vec(x_map + off * 64, 64);
        yv = vec(y_map + off * 64, 64);
      } else if (ctx->xy_mode == 1 && xy_stride == 64) {
        ARK_CUDA(ctx, cudaMemcpy2DAsync(xs, 32, xh + off * 64, 64, 32, m, cudaMemcpyHostToDevice, st));
        ARK_CUDA(ctx, cudaMemcpy2DAsync(ys, 32, yh + off * 64, 64, 32, m, cudaMemcpyHostToDevice, st));
        xv = vec(xs);
        yv = vec(ys);
      } else {
        ARK_CUDA(ctx, cudaMemcpyAsync(xs, xh + off * xy_stride, m * xy_stride, cudaMemcpyHostToDevice, st));
        ARK_CUDA(ctx, cudaMemcpyAsync(ys, yh + off * xy_stride, m * xy_stride, cudaMemcpyHostToDevice, st));
        xv = vec(xs, (uint32_t)xy_stride);
        yv = vec(ys, (uint32_t)xy_stride);
      }
      ARK_CUDA(ctx, cudaMemcpyAsync(a_dev + off * 64, ah + off * 64, m * 64, cudaMemcpyHostToDevice, st));
      ARK_CUDA(ctx, cudaMemcpyAsync(b_dev + off * 64, bh + off * 64, m * 64, cudaMemcpyHostToDevice, st));
      int rc = ARKMPC_OK;
      ARK_FIELD_SWITCH(ctx, field, rc = launch_mask<F>(ctx, st, m, xv, yv, vec(a_dev + off * 64, 64), vec(b_dev + off * 64, 64),
                                                       mvec(d_dev + off * 32), mvec(e_dev + off * 32)));
      if (rc != ARKMPC_OK) return rc;
      ARK_CUDA(ctx, cudaMemcpyAsync(dh + off * 32, d_dev + off * 32, m * 32, cudaMemcpyDeviceToHost, st));
      ARK_CUDA(ctx, cudaMemcpyAsync(eh + off * 32, e_dev + off * 32, m * 32, cudaMemcpyDeviceToHost, st));
      // c is only needed by phase 2: queue it behind the latency-critical d/e download
      ARK_CUDA(ctx, cudaMemcpyAsync(c_dev + off * 64, ch + off * 64, m * 64, cudaMemcpyHostToDevice, st));
    }
    for (int i = 0; i < kSlots; i++) ARK_CUDA(ctx, cudaStreamSynchronize(ctx->slot_stream[i]));
    return ARKMPC_OK;
  };
  const int rc = run();
  if (rc != ARKMPC_OK) {  // drain what was queued and give the resident triples back: no session survives a failed begin
    session_free(s);
    *session = nullptr;
  }
  return rc;
}
}  // namespace

extern "C" {

int arkmpc_fr_batch_mul_begin_host(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                   const uint64_t* x_host, const uint64_t* y_host, const uint64_t* a_host, const uint64_t* b_host,
                                   const uint64_t* c_host, uint64_t* de_mine_host, arkmpc_batch_mul** session) {
  return begin_host_impl(ctx, field, party_id, key_host, n, 64, x_host, y_host, a_host, b_host, c_host, de_mine_host, session);
}
int arkmpc_fr_batch_mul_begin_host_shares(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                          const uint64_t* x_share_host, const uint64_t* y_share_host, const uint64_t* a_host,
                                          const uint64_t* b_host, const uint64_t* c_host, uint64_t* de_mine_host, arkmpc_batch_mul** session) {
  return begin_host_impl(ctx, field, party_id, key_host, n, 32, x_share_host, y_share_host, a_host, b_host, c_host, de_mine_host, session);
}

int arkmpc_fr_batch_mul_finish_host(arkmpc_batch_mul* s, const uint64_t* de_peer_host, uint64_t* out_host, uint64_t* de_open_host) {
  if (!s) return ARKMPC_ERR_INVALID;
  arkmpc_ctx* ctx = s->ctx;
  ARK_CHECK_CTX(ctx);
  const size_t n = s->n;
  if (n == 0) return session_free(s);
  if (!(de_peer_host && out_host)) {
    session_free(s);
    return fail(ctx, ARKMPC_ERR_INVALID, "null pointer");
  }
  const char* dph = reinterpret_cast<const char*>(de_peer_host);
  const char* eph = dph + n * 32;
  char* oh = reinterpret_cast<char*>(out_host);
  char* doh = reinterpret_cast<char*>(de_open_host);
  char* eoh = doh ? doh + n * 32 : nullptr;
  char* a_dev = s->abc;
  char* b_dev = s->abc + n * 64;
  char* c_dev = s->abc + n * 128;
  auto run = [&]() -> int {
    int slot = 0;
    for (size_t off = 0; off < n; off += s->chunk, slot = (slot + 1) % kSlots) {
      const size_t m = (n - off < s->chunk) ? n - off : s->chunk;
      cudaStream_t st = ctx->slot_stream[slot];
      char* base = s->stage + (size_t)slot * stage_bytes_per_slot(s->chunk);
      char* dp = base;                      // chunk*32
      char* ep = base + s->chunk * 32;      // chunk*32
      char* out = base + s->chunk * 64;     // chunk*64 AoS
      // the opened d/e overwrite the peer staging planes in place (element-for-element aliasing)
      ARK_CUDA(ctx, cudaMemcpyAsync(dp, dph + off * 32, m * 32, cudaMemcpyHostToDevice, st));
      ARK_CUDA(ctx, cudaMemcpyAsync(ep, eph + off * 32, m * 32, cudaMemcpyHostToDevice, st));
      RecombineArgs g;
      g.d_mine = vec(s->de + off * 32); g.e_mine = vec(s->de + n * 32 + off * 32);
      g.d_peer = vec(dp); g.e_peer = vec(ep);
      g.a_s = vec(a_dev + off * 64, 64); g.a_m = vec(a_dev + off * 64 + 32, 64);
      g.b_s = vec(b_dev + off * 64, 64); g.b_m = vec(b_dev + off * 64 + 32, 64);
      g.c_s = vec(c_dev + off * 64, 64); g.c_m = vec(c_dev + off * 64 + 32, 64);
      g.out_s = mvec(out, 64); g.out_m = mvec(out + 32, 64);
      g.d_open = mvec(dp); g.e_open = mvec(ep);
      int rc = ARKMPC_OK;
      ARK_FIELD_SWITCH(ctx, s->field, { g.key = host_ctab<F>(s->key); rc = launch_recombine<F>(ctx, st, s->party, m, g, doh != nullptr); });
      if (rc != ARKMPC_OK) return rc;
      ARK_CUDA(ctx, cudaMemcpyAsync(oh + off * 64, out, m * 64, cudaMemcpyDeviceToHost, st));
      if (doh) {
        ARK_CUDA(ctx, cudaMemcpyAsync(doh + off * 32, dp, m * 32, cudaMemcpyDeviceToHost, st));
        ARK_CUDA(ctx, cudaMemcpyAsync(eoh + off * 32, ep, m * 32, cudaMemcpyDeviceToHost, st));
      }
    }
    for (int i = 0; i < kSlots; i++) ARK_CUDA(ctx, cudaStreamSynchronize(ctx->slot_stream[i]));
    return ARKMPC_OK;
  };
  const int rc = run();
  session_free(s);  // finish consumes the session whether it succeeded or not (slot streams drained first)
  return rc;
}

int arkmpc_fr_batch_mul_abort(arkmpc_batch_mul* s) {
  if (!s) return ARKMPC_OK;
  arkmpc_ctx* ctx = s->ctx;
  ARK_CHECK_CTX(ctx);
  return session_free(s);
}

/* bytes the host path moves per party for a batch of n gates (bench.py's h2d/d2h accounting comes from here) */
int arkmpc_fr_batch_mul_host_bytes(arkmpc_ctx* ctx, size_t n, int with_open, uint64_t* h2d_bytes, uint64_t* d2h_bytes) {
  if (!ctx || !h2d_bytes || !d2h_bytes) return ARKMPC_ERR_INVALID;
  const uint64_t xy = ctx->xy_mode == 0 ? 128 : 64;       // whole AoS x, y images, or their share halves only
  *h2d_bytes = (uint64_t)n * (xy + 3 * 64 + 64);           // x, y, triple a, b, c, peer d || e
  *d2h_bytes = (uint64_t)n * (64 + 64 + (with_open ? 64 : 0));  // own d || e, result shares, opened d || e on request
  return ARKMPC_OK;
}

}  // extern "C"
