// Scratch: how should the 32-byte share halves of an AoS ScalarShare vector (64-byte stride, pinned host memory) reach the GPU?
//  flat   cudaMemcpyAsync of the whole 64-byte image (ships the unused MAC halves)
//  2d     cudaMemcpy2DAsync, 32 of every 64 bytes
//  zc     a kernel reads the halves straight from the mapped pinned buffer (256-bit loads at stride 64)
// each alone, and with a concurrent flat H2D copy on another stream (the host path keeps a, b, c as bulk copies).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_zc tools/_zc.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void zc_read(size_t n, const char* host, char* dev, int stride) {
  const size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    uint32_t v[8];
    asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(host + i * stride) : "memory");
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "l"(dev + i * 32) : "memory");
  }
}

int main() {
  const size_t n = 1u << 20;
  char *h, *h2, *d, *d2;
  cudaMallocHost(&h, n * 64); cudaMallocHost(&h2, n * 64 * 3);
  cudaMalloc(&d, n * 64); cudaMalloc(&d2, n * 64 * 3);
  for (size_t i = 0; i < n * 64; i++) h[i] = (char)(i * 7);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t e0, e1, f0, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f0); cudaEventCreate(&f1);
  float ms, ms2;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0, s1); cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("flat 64 MB                   : %7.3f ms  %5.1f GB/s on the link, %5.1f GB/s useful\n", ms, n * 64 / ms / 1e6, n * 32 / ms / 1e6);
    cudaEventRecord(e0, s1); cudaMemcpy2DAsync(d, 32, h, 64, 32, n, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("2D 32-of-64 x 2^20           : %7.3f ms  %5.1f GB/s useful\n", ms, n * 32 / ms / 1e6);
    for (int blocks : {32, 148, 592, 4096}) {
      cudaEventRecord(e0, s1); zc_read<<<blocks, 256, 0, s1>>>(n, h, d, 64); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); printf("zero-copy stride 64, %4d blk : %7.3f ms  %5.1f GB/s useful  %s\n", blocks, ms, n * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    cudaEventRecord(e0, s1); zc_read<<<592, 256, 0, s1>>>(2 * n, h, d, 32); cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("zero-copy dense (stride 32)  : %7.3f ms  %5.1f GB/s\n", ms, n * 64 / ms / 1e6);
    // concurrent: bulk copy of 192 MB (a, b, c) on s2 while x.share is fetched on s1
    cudaDeviceSynchronize();
    cudaEventRecord(f0, s2); cudaMemcpyAsync(d2, h2, n * 192, cudaMemcpyHostToDevice, s2); cudaEventRecord(f1, s2);
    cudaEventRecord(e0, s1); zc_read<<<592, 256, 0, s1>>>(n, h, d, 64); zc_read<<<592, 256, 0, s1>>>(n, h, d, 64); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); cudaEventSynchronize(f1);
    cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
    printf("concurrent: bulk 192 MB %7.3f ms (%5.1f GB/s) + 2 x zero-copy 32 MB %7.3f ms (%5.1f GB/s useful)\n", ms2, n * 192 / ms2 / 1e6, ms, n * 64 / ms / 1e6);
    cudaEventRecord(f0, s2); cudaMemcpyAsync(d2, h2, n * 192, cudaMemcpyHostToDevice, s2); cudaEventRecord(f1, s2);
    cudaEventRecord(e0, s1); cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1); cudaEventSynchronize(f1);
    cudaEventElapsedTime(&ms, e0, e1); cudaEventElapsedTime(&ms2, f0, f1);
    printf("concurrent: bulk 192 MB %7.3f ms (%5.1f GB/s) + 2 x flat 64 MB      %7.3f ms (%5.1f GB/s useful)\n", ms2, n * 192 / ms2 / 1e6, ms, n * 64 / ms / 1e6);
  }
  return 0;
}
