// Multi-GPU plumbing behind the C ABI: the one collective of the path is the all-gather of the opened d || e of
// `batch_open` (/root/reference/online-phase/src/algebra/scalar/authenticated_scalar.rs:129-172) when every device needs the
// whole vector.  Three forms:
//   arkmpc_allgather_open                  plain ncclAllGather of this rank's opened rows (NCCL resolved at run time)
//   arkmpc_fr_beaver_recombine_gather      per-peer stores from inside the recombine kernel (CUDA IPC, arkmpc_b200.cu)
//   arkmpc_fr_beaver_recombine_gather_mc   NVSwitch multicast stores from inside the recombine kernel (this file)
// The multicast window is built with the CUDA driver's virtual-memory API (cuMulticast*, cuMem*), resolved through
// cudaGetDriverEntryPoint so that the library itself links against the runtime only and still loads on a box with no driver.
#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "ctx.hpp"

using namespace ark;
using namespace arkctx;

// ================================================================================================
// driver entry points
// ================================================================================================
namespace {

struct Drv {
  bool ok = false;
  std::string why;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};

template <class Fn>
bool resolve(Drv& d, Fn& slot, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    d.why = std::string("driver entry point not found: ") + name;
    return false;
  }
  slot = reinterpret_cast<Fn>(p);
  return true;
}

Drv& drv() {
  static Drv d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.ok = resolve(d, d.GetErrorString, "cuGetErrorString") && resolve(d, d.DeviceGet, "cuDeviceGet") &&
           resolve(d, d.DeviceGetAttribute, "cuDeviceGetAttribute") && resolve(d, d.MulticastCreate, "cuMulticastCreate") &&
           resolve(d, d.MulticastAddDevice, "cuMulticastAddDevice") && resolve(d, d.MulticastBindMem, "cuMulticastBindMem") &&
           resolve(d, d.MulticastUnbind, "cuMulticastUnbind") && resolve(d, d.MulticastGetGranularity, "cuMulticastGetGranularity") &&
           resolve(d, d.MemCreate, "cuMemCreate") && resolve(d, d.MemRelease, "cuMemRelease") &&
           resolve(d, d.MemAddressReserve, "cuMemAddressReserve") && resolve(d, d.MemAddressFree, "cuMemAddressFree") &&
           resolve(d, d.MemMap, "cuMemMap") && resolve(d, d.MemUnmap, "cuMemUnmap") && resolve(d, d.MemSetAccess, "cuMemSetAccess") &&
           resolve(d, d.MemExportToShareableHandle, "cuMemExportToShareableHandle") &&
           resolve(d, d.MemImportFromShareableHandle, "cuMemImportFromShareableHandle") &&
           resolve(d, d.MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
  });
  return d;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (drv().GetErrorString) drv().GetErrorString(r, &s);
  return s ? s : "unknown driver error";
}

#define ARK_CU(ctx, expr)                                                                   \
  do {                                                                                      \
    CUresult _r = (expr);                                                                   \
    if (_r != CUDA_SUCCESS) return fail(ctx, ARKMPC_ERR_CUDA, std::string(#expr) + ": " + cu_err(_r)); \
  } while (0)

}  // namespace

// ================================================================================================
// multicast window
// ================================================================================================
struct arkmpc_mc {
  arkmpc_ctx* ctx = nullptr;
  int world = 0, rank = 0;
  size_t bytes = 0;  // rounded up to the multicast granularity
  CUmemGenericAllocationHandle mc = 0, phys = 0;
  bool have_mc = false, have_phys = false, device_added = false, bound = false;
  int export_fd = -1;  // owner: the shareable handle other ranks duplicate; importer: our duplicate (closed after import)
  CUdeviceptr uc_va = 0, mc_va = 0;
  bool uc_mapped = false, mc_mapped = false;
};

namespace {
int mc_granularity(arkmpc_ctx* ctx, int world, size_t* gran) {
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = (unsigned)world;
  mp.size = 2u << 20;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  ARK_CU(ctx, drv().MulticastGetGranularity(gran, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
  return ARKMPC_OK;
}
}  // namespace

extern "C" {

int arkmpc_mc_supported(arkmpc_ctx* ctx, int* supported) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, supported, "null out pointer");
  *supported = 0;
  if (!drv().ok) return ARKMPC_OK;
  CUdevice dev;
  int mc = 0, fd = 0;
  if (drv().DeviceGet(&dev, ctx->device) != CUDA_SUCCESS) return ARKMPC_OK;
  drv().DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  drv().DeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  *supported = (mc && fd) ? 1 : 0;
  return ARKMPC_OK;
}

/* Step 1.  owner_pid == 0: this rank CREATES the multicast object (one rank per window, by convention rank 0) and
 * arkmpc_mc_export then yields the (pid, fd) the other ranks pass here; owner_pid > 0: duplicate fd `owner_fd` of process
 * `owner_pid` (pidfd_getfd) and import it; owner_pid < 0: `owner_fd` is already a descriptor of THIS process (the host
 * moved it, e.g. with SCM_RIGHTS).  Every rank then adds its device to the object. */
int arkmpc_mc_open(arkmpc_ctx* ctx, size_t bytes, int world, int rank, int owner_pid, int owner_fd, arkmpc_mc** out) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, out, "null out pointer");
  *out = nullptr;
  ARK_REQUIRE(ctx, world >= 2 && world <= kMaxPeers && rank >= 0 && rank < world && bytes > 0, "bad world / rank / size");
  if (!drv().ok) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, drv().why);
  int sup = 0;
  arkmpc_mc_supported(ctx, &sup);
  if (!sup) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, "device does not support multicast objects with POSIX file-descriptor handles");
  size_t gran = 0;
  int rc = mc_granularity(ctx, world, &gran);
  if (rc != ARKMPC_OK) return rc;
  arkmpc_mc* m = new (std::nothrow) arkmpc_mc();
  if (!m) return fail(ctx, ARKMPC_ERR_OOM, "mc window");
  m->ctx = ctx; m->world = world; m->rank = rank;
  m->bytes = (bytes + gran - 1) / gran * gran;
  auto body = [&]() -> int {
    CUdevice dev;
    ARK_CU(ctx, drv().DeviceGet(&dev, ctx->device));
    if (owner_pid == 0) {
      CUmulticastObjectProp mp;
      memset(&mp, 0, sizeof mp);
      mp.numDevices = (unsigned)world;
      mp.size = m->bytes;
      mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      ARK_CU(ctx, drv().MulticastCreate(&m->mc, &mp));
      m->have_mc = true;
      ARK_CU(ctx, drv().MemExportToShareableHandle(&m->export_fd, m->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    } else {
      int fd = owner_fd;
      if (owner_pid > 0) {
        const int pidfd = (int)syscall(SYS_pidfd_open, owner_pid, 0);
        if (pidfd < 0) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, std::string("pidfd_open: ") + strerror(errno));
        fd = (int)syscall(SYS_pidfd_getfd, pidfd, owner_fd, 0);
        const int e = errno;
        close(pidfd);
        if (fd < 0) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, std::string("pidfd_getfd: ") + strerror(e));
      }
      CUresult r = drv().MemImportFromShareableHandle(&m->mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      if (owner_pid > 0) close(fd);
      if (r != CUDA_SUCCESS) return fail(ctx, ARKMPC_ERR_CUDA, "cuMemImportFromShareableHandle: " + cu_err(r));
      m->have_mc = true;
    }
    ARK_CU(ctx, drv().MulticastAddDevice(m->mc, dev));
    m->device_added = true;
    return ARKMPC_OK;
  };
  rc = body();
  if (rc != ARKMPC_OK) {
    arkmpc_mc_close(m);
    return rc;
  }
  *out = m;
  return ARKMPC_OK;
}

int arkmpc_mc_export(arkmpc_mc* m, int* pid, int* fd) {
  if (!m || !pid || !fd) return ARKMPC_ERR_INVALID;
  if (m->export_fd < 0) return fail(m->ctx, ARKMPC_ERR_INVALID, "only the rank that created the multicast object can export it");
  *pid = (int)getpid();
  *fd = m->export_fd;
  return ARKMPC_OK;
}

/* Step 2, after EVERY rank has returned from arkmpc_mc_open (host barrier): allocate this rank's physical memory, bind it
 * to the multicast object and map both views.  A second host barrier must follow before the first multicast store. */
int arkmpc_mc_bind(arkmpc_mc* m, void** local_ptr, void** multicast_ptr) {
  if (!m) return ARKMPC_ERR_INVALID;
  arkmpc_ctx* ctx = m->ctx;
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, local_ptr && multicast_ptr, "null out pointer");
  ARK_REQUIRE(ctx, !m->bound, "window already bound");
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof ap);
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = ctx->device;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  ARK_CU(ctx, drv().MemCreate(&m->phys, m->bytes, &ap, 0));
  m->have_phys = true;
  ARK_CU(ctx, drv().MulticastBindMem(m->mc, 0, m->phys, 0, m->bytes, 0));
  m->bound = true;
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = ctx->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  size_t gran = 0;
  int rc = mc_granularity(ctx, m->world, &gran);
  if (rc != ARKMPC_OK) return rc;
  ARK_CU(ctx, drv().MemAddressReserve(&m->uc_va, m->bytes, gran, 0, 0));
  ARK_CU(ctx, drv().MemMap(m->uc_va, m->bytes, 0, m->phys, 0));
  m->uc_mapped = true;
  ARK_CU(ctx, drv().MemSetAccess(m->uc_va, m->bytes, &acc, 1));
  ARK_CU(ctx, drv().MemAddressReserve(&m->mc_va, m->bytes, gran, 0, 0));
  ARK_CU(ctx, drv().MemMap(m->mc_va, m->bytes, 0, m->mc, 0));
  m->mc_mapped = true;
  ARK_CU(ctx, drv().MemSetAccess(m->mc_va, m->bytes, &acc, 1));
  *local_ptr = reinterpret_cast<void*>(m->uc_va);
  *multicast_ptr = reinterpret_cast<void*>(m->mc_va);
  return ARKMPC_OK;
}

/* The caller synchronises its stream and barriers with the other ranks first. */
int arkmpc_mc_close(arkmpc_mc* m) {
  if (!m) return ARKMPC_OK;
  arkmpc_ctx* ctx = m->ctx;
  ARK_CHECK_CTX(ctx);
  Drv& d = drv();
  if (m->mc_mapped) d.MemUnmap(m->mc_va, m->bytes);
  if (m->mc_va) d.MemAddressFree(m->mc_va, m->bytes);
  if (m->uc_mapped) d.MemUnmap(m->uc_va, m->bytes);
  if (m->uc_va) d.MemAddressFree(m->uc_va, m->bytes);
  if (m->bound) {
    CUdevice dev;
    if (d.DeviceGet(&dev, ctx->device) == CUDA_SUCCESS) d.MulticastUnbind(m->mc, dev, 0, m->bytes);
  }
  if (m->have_phys) d.MemRelease(m->phys);
  if (m->have_mc) d.MemRelease(m->mc);
  if (m->export_fd >= 0) close(m->export_fd);
  delete m;
  return ARKMPC_OK;
}

int arkmpc_fr_beaver_recombine_gather_mc(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                         const uint64_t* d_mine, const uint64_t* e_mine, const uint64_t* d_peer, const uint64_t* e_peer,
                                         const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac,
                                         const uint64_t* c_share, const uint64_t* c_mac, uint64_t* out_share, uint64_t* out_mac,
                                         int world, int rank, uint64_t* gather_d_multicast, uint64_t* gather_e_multicast) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "bad world / rank");
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, key_host && gather_d_multicast && gather_e_multicast, "null pointer");
  const void* ptrs[] = {d_mine, e_mine, d_peer, e_peer, a_share, a_mac, b_share, b_mac, c_share, c_mac, out_share, out_mac, gather_d_multicast, gather_e_multicast};
  for (const void* p : ptrs) ARK_REQUIRE(ctx, p && aligned32(p), "null or misaligned plane");
  RecombineArgs g;
  g.d_mine = vec(d_mine); g.e_mine = vec(e_mine); g.d_peer = vec(d_peer); g.e_peer = vec(e_peer);
  g.a_s = vec(a_share); g.a_m = vec(a_mac); g.b_s = vec(b_share); g.b_m = vec(b_mac); g.c_s = vec(c_share); g.c_m = vec(c_mac);
  g.out_s = mvec(out_share); g.out_m = mvec(out_mac); g.d_open = mvec(nullptr); g.e_open = mvec(nullptr);
  McGatherArgs q;
  q.d = reinterpret_cast<char*>(gather_d_multicast) + (size_t)rank * n * 32;
  q.e = reinterpret_cast<char*>(gather_e_multicast) + (size_t)rank * n * 32;
  const size_t need = (n + kBlock - 1) / kBlock;
  const unsigned grid = (unsigned)(need < (1u << 30) ? need : (1u << 30));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kBlock);
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  ARK_FIELD_SWITCH(ctx, field, {
    g.key = host_ctab<F>(key_host);
    g.independent = take_hint(ctx);
    if (party_id == 0) cudaLaunchKernelEx(&cfg, beaver_recombine_gather_mc_kernel<F, 0>, n, g, q);
    else cudaLaunchKernelEx(&cfg, beaver_recombine_gather_mc_kernel<F, 1>, n, g, q);
  });
  return post_launch(ctx, "beaver_recombine_gather_mc_kernel");
}

/* rows [rank*n, (rank+1)*n) of every rank's gathered plane <- this rank's local rows, by multicast stores (no arithmetic) */
int arkmpc_mc_allgather_rows(arkmpc_ctx* ctx, size_t n, int rank, const uint64_t* local_rows, uint64_t* gathered_multicast) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, rank >= 0 && local_rows && gathered_multicast && aligned32(local_rows) && aligned32(gathered_multicast), "null or misaligned plane");
  multicast_rows_kernel<<<grid_stream(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(local_rows), reinterpret_cast<char*>(gathered_multicast) + (size_t)rank * n * 32);
  return post_launch(ctx, "multicast_rows_kernel");
}

}  // extern "C"

// ================================================================================================
// NCCL all-gather behind the C ABI (resolved at run time: libnccl.so.2, the copy a host process has already loaded if any)
// ================================================================================================
namespace {
struct Nccl {
  bool ok = false;
  std::string why;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { n.why = std::string("dlopen libnccl.so.2: ") + dlerror(); return; }
    auto sym = [&](auto& slot, const char* name) {
      slot = reinterpret_cast<std::remove_reference_t<decltype(slot)>>(dlsym(h, name));
      if (!slot) n.why = std::string("missing NCCL symbol ") + name;
      return slot != nullptr;
    };
    n.ok = sym(n.GetUniqueId, "ncclGetUniqueId") && sym(n.CommInitRank, "ncclCommInitRank") && sym(n.CommDestroy, "ncclCommDestroy") &&
           sym(n.AllGather, "ncclAllGather") && sym(n.GroupStart, "ncclGroupStart") && sym(n.GroupEnd, "ncclGroupEnd") &&
           sym(n.GetErrorString, "ncclGetErrorString");
  });
  return n;
}

#define ARK_NCCL(ctx, expr)                                                                                       \
  do {                                                                                                            \
    ncclResult_t _r = (expr);                                                                                     \
    if (_r != ncclSuccess) return fail(ctx, ARKMPC_ERR_NCCL, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
  } while (0)
}  // namespace

extern "C" {

int arkmpc_nccl_unique_id(uint8_t* id_out) {
  if (!id_out) return ARKMPC_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == ARKMPC_NCCL_ID_BYTES, "ncclUniqueId size");
  if (!nccl().ok) return fail(nullptr, ARKMPC_ERR_UNSUPPORTED, nccl().why);
  ncclUniqueId id;
  ARK_NCCL(nullptr, nccl().GetUniqueId(&id));
  memcpy(id_out, &id, sizeof id);
  return ARKMPC_OK;
}

/* Collective over the `world` ranks: every rank passes the id rank 0 obtained from arkmpc_nccl_unique_id. */
int arkmpc_nccl_init(arkmpc_ctx* ctx, int world, int rank, const uint8_t* id_bytes) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, id_bytes && world >= 1 && rank >= 0 && rank < world, "bad world / rank / id");
  ARK_REQUIRE(ctx, !ctx->nccl, "the context already has a communicator");
  if (!nccl().ok) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, nccl().why);
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  ncclComm_t comm = nullptr;
  ARK_NCCL(ctx, nccl().CommInitRank(&comm, world, id, rank));
  ctx->nccl = comm;
  ctx->nccl_world = world;
  ctx->nccl_rank = rank;
  return ARKMPC_OK;
}

int arkmpc_nccl_destroy(arkmpc_ctx* ctx) {
  if (!ctx) return ARKMPC_ERR_INVALID;
  if (!ctx->nccl) return ARKMPC_OK;
  ARK_CHECK_CTX(ctx);
  cudaStreamSynchronize(ctx->stream);
  nccl().CommDestroy(static_cast<ncclComm_t>(ctx->nccl));
  ctx->nccl = nullptr;
  ctx->nccl_world = 0;
  ctx->nccl_rank = -1;
  return ARKMPC_OK;
}

/* batch_open on a sharded batch: this rank's n_local opened rows of d and of e -> rows
 * [rank*n_local, (rank+1)*n_local) of the world*n_local-row gathered planes on every rank; two ncclAllGather in one group
 * on the context's stream.  d_local / e_local may point into d_all / e_all at this rank's block (in place). */
int arkmpc_allgather_open(arkmpc_ctx* ctx, size_t n_local, const uint64_t* d_local, const uint64_t* e_local, uint64_t* d_all, uint64_t* e_all) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, ctx->nccl, "no communicator: call arkmpc_nccl_init first");
  if (n_local == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, d_local && d_all && (e_local == nullptr) == (e_all == nullptr), "null pointer");
  ncclComm_t comm = static_cast<ncclComm_t>(ctx->nccl);
  ARK_NCCL(ctx, nccl().GroupStart());
  ncclResult_t r1 = nccl().AllGather(d_local, d_all, n_local * 32, ncclUint8, comm, ctx->stream);
  ncclResult_t r2 = e_local ? nccl().AllGather(e_local, e_all, n_local * 32, ncclUint8, comm, ctx->stream) : ncclSuccess;
  ncclResult_t r3 = nccl().GroupEnd();
  ARK_NCCL(ctx, r1);
  ARK_NCCL(ctx, r2);
  ARK_NCCL(ctx, r3);
  return ARKMPC_OK;
}

}  // extern "C"
