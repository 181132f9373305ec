// C ABI: device memory for hosts that are not torch (the C++ / Rust mirrors of the reference's fabric), declared in
// include/arkmpc_b200.h.  The reference allocates a fresh Vec for every gate result (fabric/result.rs: one ResultValue per
// handle); its GPU mirror therefore allocates and frees ~60 device buffers per 1024-gate batch_mul + open_authenticated, and with
// plain cudaMalloc / cudaFree that was 45 % of the iteration (14 us + 11 us per pair at 32 KB, 520 us + 450 us at 2 MB:
// profiles/r02o_config0_profile.txt).  So:
//
//  * arkmpc_malloc / arkmpc_free keep freed blocks in a per-device cache, by size class.  A block is handed out again without
//    any host synchronisation: before a context reuses a block, an event is recorded on the CURRENT stream of every live
//    context of that device (any of them may have work on the block in flight: device buffers travel between the two parties'
//    contexts by reference) and the reusing context's stream waits for them.  The recording is shared: it covers every block
//    freed before it, so a burst of frees and allocations costs one recording per context (arkmpc_free itself makes no CUDA
//    call: 0.4 us against 6 us with an event per free and 11 us for cudaFree).  Work submitted on other streams is the caller's
//    to synchronise before the free.
//  * arkmpc_memcpy_h2d stages copies of up to 256 KB through a pinned ring (8 slots, owned by the context, recycled through a
//    per-device pool): the source may be released as soon as the call returns, and a 32 KB copy no longer costs the 38 us of a
//    pageable cudaMemcpyAsync.
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
  uint64_t freed_at;  // value of DeviceMem::free_seq when the block was freed
};

struct Member {  // a live context of the device
  arkmpc_ctx* ctx;
  cudaEvent_t ev;          // this context's entry in the current event set
  uint64_t waited_set = 0; // the last event set this context's stream waits for ...
  cudaStream_t waited_on = nullptr;  // ... and the stream that wait was enqueued on
};

struct StageRing {
  char* base = nullptr;
  cudaEvent_t ev[kStageSlots] = {};
  int next = 0;
};

struct DeviceMem {
  std::unordered_map<void*, size_t> live;  // handed out: ptr -> size class
  std::multimap<size_t, FreeBlock> cache;  // size class -> free blocks
  size_t cached_bytes = 0;
  std::vector<Member> members;
  // The event set: one event per member, recorded on its current stream.  It is renewed lazily, when a block freed after the
  // last recording is about to be reused; a burst of frees followed by a burst of allocations costs one recording per context.
  uint64_t free_seq = 0;   // counts frees
  uint64_t set_seq = 0;    // free_seq when the set was last recorded: it covers every block with freed_at <= set_seq
  uint64_t set_id = 0;     // counts recordings
  std::vector<StageRing*> idle_rings;
};

std::mutex g_mu;  // guards every DeviceMem and every context's `stream` field
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

// cudaFree of every cached block (each synchronises the device); g_mu held
void trim_locked(DeviceMem& d) {
  for (auto& kv : d.cache) cudaFree(kv.second.ptr);
  d.cache.clear();
  d.cached_bytes = 0;
}

// Orders `ctx`'s stream after everything submitted so far on the current stream of every other live context (and of its own
// previous stream, if it was re-pointed), unless an event set at least as recent as `freed_at` is already waited for.  g_mu held.
bool order_after_free(DeviceMem& d, arkmpc_ctx* ctx, uint64_t freed_at) {
  if (d.set_seq < freed_at) {
    for (Member& m : d.members)
      if (cudaEventRecord(m.ev, m.ctx->stream) != cudaSuccess) { cudaGetLastError(); return false; }
    d.set_seq = d.free_seq;
    d.set_id++;
  }
  for (Member& me : d.members) {
    if (me.ctx != ctx) continue;
    if (me.waited_set == d.set_id && me.waited_on == ctx->stream) return true;
    for (Member& m : d.members) {
      if (m.ctx == ctx && me.waited_on == ctx->stream) continue;  // the own stream is in order already
      if (cudaStreamWaitEvent(ctx->stream, m.ev, 0) != cudaSuccess) { cudaGetLastError(); return false; }
    }
    me.waited_set = d.set_id;
    me.waited_on = ctx->stream;
    return true;
  }
  return false;  // not a registered context
}

}  // namespace

namespace arkctx {

void mem_register(arkmpc_ctx* ctx) {
  if (ctx->device < 0 || ctx->device >= kMaxDevices) return;
  Member m;
  m.ctx = ctx;
  m.ev = nullptr;
  if (cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
  std::lock_guard<std::mutex> l(g_mu);
  DeviceMem& d = g_dev[ctx->device];
  // a newcomer has no work on cached blocks, but the current set holds no event of its own: start it at "nothing waited for"
  d.members.push_back(m);
}
void mem_unregister(arkmpc_ctx* ctx) {
  if (ctx->device < 0 || ctx->device >= kMaxDevices) return;
  StageRing* ring = static_cast<StageRing*>(ctx->stage_ring);
  ctx->stage_ring = nullptr;
  std::lock_guard<std::mutex> l(g_mu);
  DeviceMem& d = g_dev[ctx->device];
  if (ring) d.idle_rings.push_back(ring);  // the caller synchronises the context's streams before the ring is used again
  for (size_t i = 0; i < d.members.size(); i++)
    if (d.members[i].ctx == ctx) {
      cudaEventDestroy(d.members[i].ev);
      d.members.erase(d.members.begin() + i);
      break;
    }
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
    if (it != d.cache.end() && order_after_free(d, ctx, it->second.freed_at)) {
      void* p = it->second.ptr;
      d.cache.erase(it);
      d.cached_bytes -= cls;
      d.live[p] = cls;
      *dev_ptr = p;
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
        d.cache.emplace(cls, FreeBlock{dev_ptr, ++d.free_seq});
        d.cached_bytes += cls;
        return ARKMPC_OK;
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
    // the ring belongs to this context until it is destroyed (the context lock is held: no other thread is in here); rings
    // outlive contexts in a per-device pool, so a short-lived context does not pay for pinning memory
    StageRing* ring = static_cast<StageRing*>(ctx->stage_ring);
    if (!ring) {
      {
        std::lock_guard<std::mutex> l(g_mu);
        auto& idle = g_dev[ctx->device].idle_rings;
        if (!idle.empty()) { ring = idle.back(); idle.pop_back(); }
      }
      if (!ring) {
        void* p = nullptr;
        if (cudaMallocHost(&p, kStageSlotBytes * kStageSlots) == cudaSuccess) {
          ring = new StageRing();
          ring->base = static_cast<char*>(p);
          for (int i = 0; i < kStageSlots; i++) cudaEventCreateWithFlags(&ring->ev[i], cudaEventDisableTiming);
        } else {
          cudaGetLastError();
        }
      }
      ctx->stage_ring = ring;
    }
    if (ring) {
      const int slot = ring->next;
      ring->next = (slot + 1) % kStageSlots;
      ARK_CUDA(ctx, cudaEventSynchronize(ring->ev[slot]));  // the copy that used this slot eight copies ago
      char* s = ring->base + (size_t)slot * kStageSlotBytes;
      memcpy(s, src_host, bytes);
      ARK_CUDA(ctx, cudaMemcpyAsync(dst_dev, s, bytes, cudaMemcpyHostToDevice, ctx->stream));
      ARK_CUDA(ctx, cudaEventRecord(ring->ev[slot], ctx->stream));
      return ARKMPC_OK;
    }
  }
  ARK_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return ARKMPC_OK;
}

}  // extern "C"
