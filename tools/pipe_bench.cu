// Pipe-throughput microbenchmarks for sm_100a: how fast do the integer multiply, wide multiply-add,
// add and FP64 pipes issue, alone and mixed, and what does that make a 256-bit Montgomery product cost?
// Gives the INT32 "peak" the modular-multiplication kernels are set against.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/pipe_bench tools/pipe_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../ark_mpc_b200/csrc/fp256.cuh"
using namespace ark;

constexpr int ITERS = 512;
constexpr int THREADS = 256;

__global__ void k_imad(uint32_t* out, uint32_t c) {
  uint32_t x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(c), "r"(x[(k + 1) & 7]));
  }
  uint32_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x[k];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void k_imad_wide(uint32_t* out, uint32_t c) {
  uint64_t x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"((uint32_t)x[(k + 1) & 7]), "r"(c));
  }
  uint64_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x[k];
  if (s == 0x12345678u) out[0] = (uint32_t)s;
}
// the exact row pattern of the Montgomery code: two carry chains of four IMAD.WIDE.U32.X
__global__ void k_row(uint32_t* out, uint32_t c) {
  MontAcc t;
  acc_zero(t);
  uint32_t a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 2654435761u + k + c;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) acc_row(t, a, c + r);
  }
  uint32_t s = 0;
  for (int k = 0; k < 8; k++) s ^= t.E[k] ^ t.O[k];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void k_iadd3(uint32_t* out, uint32_t c) {
  uint32_t x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(x[(k + 3) & 7]));
  }
  uint32_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x[k];
  if (s == 0x12345678u) out[0] = s;
}
__global__ void k_dfma(uint32_t* out, double c) {
  double x[8];
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(x[k]) : "d"(c));
  }
  double s = 0;
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == 0.12345) out[0] = 1;
}
__global__ void k_mix_wide_dfma(uint32_t* out, uint32_t c, double cd) {
  uint64_t x[8];
  double y[8];
  for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x + k; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"((uint32_t)x[(k + 1) & 7]), "r"(c));
        asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(y[k]) : "d"(cd));
      }
  }
  uint64_t s = 0;
  double sd = 0;
  for (int k = 0; k < 8; k++) { s ^= x[k]; sd += y[k]; }
  if (s == 0x12345678u || sd == 0.12345) out[0] = (uint32_t)s;
}
__global__ void k_mix_wide_iadd(uint32_t* out, uint32_t c) {
  uint64_t x[8];
  uint32_t y[8];
  for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x + k; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"((uint32_t)x[(k + 1) & 7]), "r"(c));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(y[(k + 3) & 7]));
      }
  }
  uint64_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x[k] ^ y[k];
  if (s == 0x12345678u) out[0] = (uint32_t)s;
}
template <class F>
__global__ void k_montmul(uint32_t* out, uint32_t c) {
  fe8 x, y;
  for (int k = 0; k < 8; k++) { x.v[k] = threadIdx.x * 2654435761u + k; y.v[k] = c + k; }
  x.v[7] &= 0x0fffffffu; y.v[7] &= 0x0fffffffu;
  for (int it = 0; it < ITERS; it++) {
    fe8 r;
    Fp<F>::mul(r, x, y);
    x = r;
  }
  uint32_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x.v[k];
  if (s == 0x12345678u) out[0] = s;
}
template <class F>
__global__ void k_montmul2(uint32_t* out, uint32_t c) {
  fe8 x, y, z;
  for (int k = 0; k < 8; k++) { x.v[k] = threadIdx.x * 2654435761u + k; y.v[k] = c + k; z.v[k] = c * 3 + k; }
  x.v[7] &= 0x0fffffffu; y.v[7] &= 0x0fffffffu; z.v[7] &= 0x0fffffffu;
  for (int it = 0; it < ITERS; it++) {
    fe8 r;
    Fp<F>::mul2_lazy(r, x, y, z, x);
    Fp<F>::csub_p(r);
    x = r;
  }
  uint32_t s = 0;
  for (int k = 0; k < 8; k++) s ^= x.v[k];
  if (s == 0x12345678u) out[0] = s;
}

template <class K, class... A>
double run(const char* name, double ops_per_thread, int blocks_per_sm, int sms, double clk_ghz, K kern, A... args) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int grid = sms * blocks_per_sm;
  kern<<<grid, THREADS>>>(args...);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    kern<<<grid, THREADS>>>(args...);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double total = ops_per_thread * (double)grid * THREADS;
  double per_s = total / (best * 1e-3);
  printf("%-22s blocks/SM=%d  %8.3f ms  %10.2f Gop/s  %7.2f op/clk/SM (at %.3f GHz)\n", name, blocks_per_sm, best, per_s * 1e-9,
         per_s / sms / (clk_ghz * 1e9), clk_ghz);
  return per_s;
}

__global__ void k_clock(long long* out) {
  long long c0 = clock64();
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 20000000ull);
  long long c1 = clock64();
  out[0] = c1 - c0;
  out[1] = (long long)(t1 - t0);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  uint32_t* out;
  cudaMalloc(&out, 64);
  long long* ck;
  cudaMallocManaged(&ck, 16);
  // warm the clocks with some work, then measure SM clock
  for (int i = 0; i < 20; i++) k_imad<<<sms * 4, THREADS>>>(out, 3);
  k_clock<<<1, 1>>>(ck);
  cudaDeviceSynchronize();
  double ghz = (double)ck[0] / (double)ck[1];
  printf("device %s, %d SMs, measured SM clock %.3f GHz (idle-ish)\n", p.name, sms, ghz);
  const double per = (double)ITERS * 64;
  for (int b : {2, 4, 8}) {
    run("IMAD", per, b, sms, ghz, k_imad, out, 3u);
    run("IMAD.WIDE", per, b, sms, ghz, k_imad_wide, out, 3u);
    run("IMAD.WIDE.X row", per, b, sms, ghz, k_row, out, 3u);
    run("IADD3", per, b, sms, ghz, k_iadd3, out, 3u);
    run("DFMA", per, b, sms, ghz, k_dfma, out, 1.000001);
    run("IMAD.WIDE+DFMA (ops)", per, b, sms, ghz, k_mix_wide_dfma, out, 3u, 1.000001);
    run("IMAD.WIDE+IADD (ops)", per, b, sms, ghz, k_mix_wide_iadd, out, 3u);
    run("montmul bn254_fr", (double)ITERS, b, sms, ghz, k_montmul<Bn254Fr>, out, 3u);
    run("montmul c25519_fr", (double)ITERS, b, sms, ghz, k_montmul<Curve25519Fr>, out, 3u);
    run("montmul2 bn254_fr", (double)ITERS, b, sms, ghz, k_montmul2<Bn254Fr>, out, 3u);
  }
  k_clock<<<1, 1>>>(ck);
  cudaDeviceSynchronize();
  printf("SM clock after: %.3f GHz\n", (double)ck[0] / (double)ck[1]);
  return 0;
}
