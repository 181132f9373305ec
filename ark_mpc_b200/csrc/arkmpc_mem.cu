// C ABI: device memory for hosts that are not torch (the C++ / Rust mirrors of the reference's fabric), declared in
// include/arkmpc_b200.h.  The reference allocates a fresh Vec for every gate result (fabric/result.rs: one ResultValue per
// handle); its GPU mirror therefore allocates and frees ~60 device buffers per 1024-gate batch_mul + open_authenticated, and with
// plain cudaMalloc / cudaFree that was 45 % of the iteration (14 us + 11 us per pair at 32 KB, 520 us + 450 us at 2 MB:
// profiles/r02o_config0_profile.txt).  So:
//
//  * arkmpc_malloc / arkmpc_free keep freed blocks in a per-device cache, by size class.  A block is handed out again without
//    any host synchronisation: arkmpc_free records an event on the CURRENT stream of every live context of that device (any of
//    them may have work on the block in flight: device buffers travel between the two parties' contexts by reference), and the
//    context that receives the block next makes its stream wait for those events.  Work submitted on other streams is the
//    caller's to synchronise before the free.
//  * arkmpc_memcpy_h2d stages copies of up to 256 KB through a pinned ring (per device, 8 slots): 5 us instead of the 38 us a
//    pageable cudaMemcpyAsync costs at 32 KB, and the source may be released as soon as the call returns.
//
// ARKMPC_ALLOC_CACHE_MB caps the bytes held in free blocks (default 4096; 0 = plain cudaMalloc / cudaFree).
#include <map>
#include <unordered_map>
#include <vector>

#include "ctx.hpp"

using namespace arkctx;

namespace {

constexpr int kMaxDevices = 64;
constexpr size_t kStageSlotBytes = 256u << 10;
constexpr int kStageSlots = 8;

struct FreeBlock {
  void* ptr;
  std::vector<cudaEvent_t> events;  // everything that may still touch the block
};

struct DeviceMem {
  std::unordered_map<void*, size_t> live;                // handed out: ptr -> size class
  std::multimap<size_t, FreeBlock> cache;                // size class -> free blocks
  size_t cached_bytes = 0;
  std::vector<cudaEvent_t> event_pool;
  std::vector<arkmpc_ctx*> contexts;                     // live contexts of this device
  // pinned staging ring
  std::mutex stage_mu;
  char* stage = nullptr;
  cudaEvent_t stage_ev[kStageSlots] = {};
  int stage_next = 0;
};

std::mutex g_mu;  // guards every DeviceMem except its staging ring, and every context's `stream` field
DeviceMem g_dev[kMaxDevices];

size_t cache_cap() {
  static const size_t cap = [] {
    const char* e = getenv("ARKMPC_ALLOC_CACHE_MB");
    return (size_t)(e && *e ? atol(e) : 4096) << 20;
  }();
  return cap;
}

// < 1 MiB: next power of two (>= 512 B); above: next multiple of 2 MiB
size_t size_class(size_t bytes) {
  if (bytes <= ((size_t)1 << 20)) {
    size_t c = 512;
    while (c < bytes) c <<= 1;
    return c;
  }
  const size_t g = (size_t)2 << 20;
  return (bytes + g - 1) / g * g;
}

cudaEvent_t get_event(DeviceMem& d) {
  if (!d.event_pool.empty()) {
    cudaEvent_t e = d.event_pool.back();
    d.event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return e;
}

// cudaFree of every cached block (each synchronises the device); g_mu held
void trim_locked(DeviceMem& d) {
  for (auto& kv : d.cache) {
    cudaFree(kv.second.ptr);
    for (cudaEvent_t e : kv.second.events) d.event_pool.push_back(e);
  }
  d.cache.clear();
  d.cached_bytes = 0;
}

}  // namespace

namespace arkctx {

void mem_register(arkmpc_ctx* ctx) {
  if (ctx->device < 0 || ctx->device >= kMaxDevices) return;
  std::lock_guard<std::mutex> l(g_mu);
  g_dev[ctx->device].contexts.push_back(ctx);
}
void mem_unregister(arkmpc_ctx* ctx) {
  if (ctx->device < 0 || ctx->device >= kMaxDevices) return;
  std::lock_guard<std::mutex> l(g_mu);
  auto& v = g_dev[ctx->device].contexts;
  for (size_t i = 0; i < v.size(); i++)
    if (v[i] == ctx) { v.erase(v.begin() + i); break; }
}
void mem_set_stream(arkmpc_ctx* ctx, cudaStream_t s) {
  std::lock_guard<std::mutex> l(g_mu);
  ctx->stream = s;
}

}  // namespace arkctx

extern "C" {

int arkmpc_malloc(arkmpc_ctx* ctx, size_t bytes, void** dev_ptr) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, dev_ptr, "null out pointer");
  *dev_ptr = nullptr;
  if (bytes == 0) return ARKMPC_OK;
  if (cache_cap() == 0 || ctx->device >= kMaxDevices) {
    ARK_CUDA(ctx, cudaMalloc(dev_ptr, bytes));
    return ARKMPC_OK;
  }
  DeviceMem& d = g_dev[ctx->device];
  const size_t cls = size_class(bytes);
  {
    std::lock_guard<std::mutex> l(g_mu);
    auto it = d.cache.find(cls);
    if (it != d.cache.end()) {
      FreeBlock b = std::move(it->second);
      d.cache.erase(it);
      d.cached_bytes -= cls;
      for (cudaEvent_t e : b.events) {
        cudaStreamWaitEvent(ctx->stream, e, 0);
        d.event_pool.push_back(e);  // a later record on a pooled event does not disturb a wait already enqueued
      }
      d.live[b.ptr] = cls;
      *dev_ptr = b.ptr;
      return ARKMPC_OK;
    }
  }
  cudaError_t e = cudaMalloc(dev_ptr, cls);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back and try once more
    cudaGetLastError();
    { std::lock_guard<std::mutex> l(g_mu); trim_locked(d); }
    e = cudaMalloc(dev_ptr, cls);
  }
  ARK_CUDA(ctx, e);
  std::lock_guard<std::mutex> l(g_mu);
  d.live[*dev_ptr] = cls;
  return ARKMPC_OK;
}

int arkmpc_free(arkmpc_ctx* ctx, void* dev_ptr) {
  ARK_CHECK_CTX(ctx);
  if (!dev_ptr) return ARKMPC_OK;
  if (ctx->device < kMaxDevices) {
    DeviceMem& d = g_dev[ctx->device];
    std::lock_guard<std::mutex> l(g_mu);
    auto it = d.live.find(dev_ptr);
    if (it != d.live.end()) {
      const size_t cls = it->second;
      d.live.erase(it);
      if (d.cached_bytes + cls <= cache_cap()) {
        FreeBlock b;
        b.ptr = dev_ptr;
        bool ok = true;
        for (arkmpc_ctx* c : d.contexts) {
          cudaEvent_t e = get_event(d);
          if (!e || cudaEventRecord(e, c->stream) != cudaSuccess) { cudaGetLastError(); ok = false; if (e) d.event_pool.push_back(e); break; }
          b.events.push_back(e);
        }
        if (ok) {
          d.cache.emplace(cls, std::move(b));
          d.cached_bytes += cls;
          return ARKMPC_OK;
        }
        for (cudaEvent_t e : b.events) d.event_pool.push_back(e);
      }
    }
  }
  ARK_CUDA(ctx, cudaFree(dev_ptr));  // not cached (cache off, over the cap, or not from arkmpc_malloc's cache): synchronising free
  return ARKMPC_OK;
}

int arkmpc_mem_trim(arkmpc_ctx* ctx) {
  ARK_CHECK_CTX(ctx);
  if (ctx->device >= kMaxDevices) return ARKMPC_OK;
  std::lock_guard<std::mutex> l(g_mu);
  trim_locked(g_dev[ctx->device]);
  return ARKMPC_OK;
}

int arkmpc_mem_cached_bytes(arkmpc_ctx* ctx, size_t* bytes) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, bytes, "null out pointer");
  std::lock_guard<std::mutex> l(g_mu);
  *bytes = ctx->device < kMaxDevices ? g_dev[ctx->device].cached_bytes : 0;
  return ARKMPC_OK;
}

int arkmpc_host_alloc(arkmpc_ctx* ctx, size_t bytes, void** pinned_ptr) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, pinned_ptr, "null out pointer");
  *pinned_ptr = nullptr;
  if (bytes == 0) return ARKMPC_OK;
  ARK_CUDA(ctx, cudaMallocHost(pinned_ptr, bytes));
  return ARKMPC_OK;
}
int arkmpc_host_free(arkmpc_ctx* ctx, void* pinned_ptr) {
  ARK_CHECK_CTX(ctx);
  if (pinned_ptr) ARK_CUDA(ctx, cudaFreeHost(pinned_ptr));
  return ARKMPC_OK;
}

int arkmpc_memcpy_h2d(arkmpc_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  ARK_CHECK_CTX(ctx);
  if (bytes == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, dst_dev && src_host, "null pointer");
  if (bytes <= kStageSlotBytes && cache_cap() != 0 && ctx->device < kMaxDevices) {
    DeviceMem& d = g_dev[ctx->device];
    std::lock_guard<std::mutex> l(d.stage_mu);
    if (!d.stage) {
      void* p = nullptr;
      if (cudaMallocHost(&p, kStageSlotBytes * kStageSlots) == cudaSuccess) {
        d.stage = static_cast<char*>(p);
        for (int i = 0; i < kStageSlots; i++) cudaEventCreateWithFlags(&d.stage_ev[i], cudaEventDisableTiming);
      } else {
        cudaGetLastError();
      }
    }
    if (d.stage) {
      const int slot = d.stage_next;
      d.stage_next = (slot + 1) % kStageSlots;
      ARK_CUDA(ctx, cudaEventSynchronize(d.stage_ev[slot]));  // the copy that used this slot eight copies ago
      char* s = d.stage + (size_t)slot * kStageSlotBytes;
      memcpy(s, src_host, bytes);
      ARK_CUDA(ctx, cudaMemcpyAsync(dst_dev, s, bytes, cudaMemcpyHostToDevice, ctx->stream));
      ARK_CUDA(ctx, cudaEventRecord(d.stage_ev[slot], ctx->stream));
      return ARKMPC_OK;
    }
  }
  ARK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ARKMPC_OK;
}

}  // extern "C"
