// Scratch: is a strided (2D) H2D copy of the 32-byte share halves of an AoS ScalarShare vector cheaper than copying all 64 bytes?
#include <cuda_runtime.h>
#include <cstdio>
int main() {
  const size_t n = 1u << 20;
  char *h, *d;
  cudaMallocHost(&h, n * 64);
  cudaMalloc(&d, n * 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0); cudaMemcpyAsync(d, h, n * 64, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("flat 64 MB        : %.3f ms  %.1f GB/s\n", ms, n * 64 / ms / 1e6);
    cudaEventRecord(e0); cudaMemcpy2DAsync(d, 32, h, 64, 32, n, cudaMemcpyHostToDevice); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("2D 32-of-64 x 2^20: %.3f ms  %.1f GB/s useful\n", ms, n * 32 / ms / 1e6);
    cudaEventRecord(e0); cudaMemcpyAsync(h, d, n * 64, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); printf("flat D2H 64 MB    : %.3f ms  %.1f GB/s\n", ms, n * 64 / ms / 1e6);
  }
  return 0;
}
