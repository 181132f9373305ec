// NVSwitch multicast probe (single process, all visible devices): can a multicast object be created, bound to one
// physical allocation per device and written with multimem.st so that one store lands in every device's memory?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/mc_probe tools/mc_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#define CK(x)                                                        \
  do {                                                               \
    CUresult _r = (x);                                               \
    if (_r != CUDA_SUCCESS) {                                        \
      const char* s = nullptr;                                       \
      cuGetErrorString(_r, &s);                                      \
      printf("FAIL %s -> %d (%s)\n", #x, (int)_r, s ? s : "?");      \
      return 1;                                                      \
    }                                                                \
  } while (0)

__global__ void mc_store(uint32_t* mc, size_t words, uint32_t tag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 4 + 3 < words) {
    uint32_t a = tag + (uint32_t)i * 4;
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i * 4), "f"(__uint_as_float(a)), "f"(__uint_as_float(a + 1)),
                 "f"(__uint_as_float(a + 2)), "f"(__uint_as_float(a + 3))
                 : "memory");
  }
}

int main() {
  CK(cuInit(0));
  int ndev = 0;
  CK(cuDeviceGetCount(&ndev));
  printf("devices: %d\n", ndev);
  std::vector<CUdevice> dev(ndev);
  std::vector<CUcontext> ctxs(ndev);
  for (int d = 0; d < ndev; d++) {
    CK(cuDeviceGet(&dev[d], d));
    int mc = 0, fd = 0, fab = 0, vmm = 0;
    cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev[d]);
    cuDeviceGetAttribute(&fd, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev[d]);
    cuDeviceGetAttribute(&fab, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev[d]);
    cuDeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev[d]);
    printf("dev %d: multicast=%d posix_fd=%d fabric=%d vmm=%d\n", d, mc, fd, fab, vmm);
    CK(cuDevicePrimaryCtxRetain(&ctxs[d], dev[d]));
  }
  for (int a = 0; a < ndev; a++)
    for (int b = 0; b < ndev; b++)
      if (a != b) {
        int can = 0;
        cuDeviceCanAccessPeer(&can, dev[a], dev[b]);
        if (!can) printf("no peer access %d -> %d\n", a, b);
      }
  const size_t want = 4u << 20;
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = ndev;
  mp.size = want;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0, gran_min = 0;
  CK(cuMulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  CK(cuMulticastGetGranularity(&gran_min, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
  printf("multicast granularity: recommended %zu, minimum %zu\n", gran, gran_min);
  const size_t size = (want + gran - 1) / gran * gran;
  mp.size = size;
  CUmemGenericAllocationHandle mch;
  CK(cuMulticastCreate(&mch, &mp));
  int fdh = -1;
  CK(cuMemExportToShareableHandle(&fdh, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  printf("multicast object created (%zu B), exported as fd %d\n", size, fdh);
  for (int d = 0; d < ndev; d++) CK(cuMulticastAddDevice(mch, dev[d]));
  std::vector<CUmemGenericAllocationHandle> phys(ndev);
  std::vector<CUdeviceptr> uc(ndev), mcva(ndev);
  for (int d = 0; d < ndev; d++) {
    CK(cuCtxSetCurrent(ctxs[d]));
    CUmemAllocationProp ap;
    memset(&ap, 0, sizeof ap);
    ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    ap.location.id = d;
    ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t ag = 0;
    CK(cuMemGetAllocationGranularity(&ag, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
    if (d == 0) printf("allocation granularity %zu\n", ag);
    CK(cuMemCreate(&phys[d], size, &ap, 0));
    CK(cuMulticastBindMem(mch, 0, phys[d], 0, size, 0));
    CUmemAccessDesc acc;
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = d;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    CK(cuMemAddressReserve(&uc[d], size, gran, 0, 0));
    CK(cuMemMap(uc[d], size, 0, phys[d], 0));
    CK(cuMemSetAccess(uc[d], size, &acc, 1));
    CK(cuMemAddressReserve(&mcva[d], size, gran, 0, 0));
    CK(cuMemMap(mcva[d], size, 0, mch, 0));
    CK(cuMemSetAccess(mcva[d], size, &acc, 1));
    cudaMemset((void*)uc[d], 0, size);
    cudaDeviceSynchronize();
  }
  // device 0 stores through its multicast mapping; every device's unicast view must show the data
  CK(cuCtxSetCurrent(ctxs[0]));
  const size_t words = size / 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  mc_store<<<(unsigned)((words / 4 + 255) / 256), 256>>>((uint32_t*)mcva[0], words, 0xA0000000u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 20; r++) mc_store<<<(unsigned)((words / 4 + 255) / 256), 256>>>((uint32_t*)mcva[0], words, 0xA0000000u);
  cudaEventRecord(e1);
  cudaError_t ce = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("multimem.st kernel: %s, %.2f us per %zu-byte broadcast (%.1f GB/s payload)\n", cudaGetErrorString(ce), 1e3 * ms / 20, size,
         size / (ms / 20 * 1e-3) / 1e9);
  int bad = 0;
  std::vector<uint32_t> h(words);
  for (int d = 0; d < ndev; d++) {
    CK(cuCtxSetCurrent(ctxs[d]));
    cudaMemcpy(h.data(), (void*)uc[d], size, cudaMemcpyDeviceToHost);
    size_t wrong = 0;
    for (size_t i = 0; i < words; i++) wrong += h[i] != 0xA0000000u + (uint32_t)i;
    printf("dev %d: %zu wrong words of %zu\n", d, wrong, words);
    bad += wrong != 0;
  }
  printf(bad ? "MULTICAST PROBE: FAIL\n" : "MULTICAST PROBE: OK\n");
  return bad;
}
