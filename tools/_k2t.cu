// Scratch: where does the fused recombine kernel's time go?  Same kernel body as beaver_recombine_kernel with per-block
// start/end timestamps (globaltimer) and SM ids; prints the ramp-up (first wave), the steady state and the drain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/_k2t tools/_k2t.cu
#include <algorithm>
#include <cstdio>
#include <vector>
#include "../ark_mpc_b200/csrc/ctab.hpp"
#include "../ark_mpc_b200/csrc/fr_kernels.cuh"
using namespace ark;

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <class F, int BLK, int MINB, int MODE>
__global__ void __launch_bounds__(BLK, MINB) k2t(size_t n, const __grid_constant__ RecombineArgs g, unsigned long long* stamps) {
  unsigned long long t0 = 0;
  if (stamps && threadIdx.x == 0) t0 = gtime();
  const size_t i = (size_t)blockIdx.x * BLK + threadIdx.x;
  if (i < n) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i); ld_fe(dp, g.d_peer, i); ld_fe(em, g.e_mine, i); ld_fe(ep, g.e_peer, i);
    ld_fe(bs, g.b_s, i); ld_fe(as, g.a_s, i);
    if (MODE == 0) { ld_fe(bm, g.b_m, i); ld_fe(am, g.a_m, i); ld_fe(cs, g.c_s, i); ld_fe(cm, g.c_m, i); }
    fe8 os, om, d, e;
    if (MODE == 0) {
      beaver_recombine_elem<F>(os, om, d, e, 0, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
      st_fe(g.out_s, i, os); st_fe(g.out_m, i, om);
    } else {
      // MODE 1: share half first (7 operands in flight), the three mac operands are requested after d and e are known
      Fp<F>::add(d, dm, dp);
      Fp<F>::add(e, em, ep);
      ld_fe(cs, g.c_s, i);
      ld_fe(bm, g.b_m, i); ld_fe(am, g.a_m, i); ld_fe(cm, g.c_m, i);
      fe8 x, s;
      Fp<F>::add_raw(x, bs, e);
      Fp<F>::mul2_lazy(s, d, x, e, as);
      Fp<F>::csub_p(s);
      Fp<F>::add(os, s, cs);
      st_fe(g.out_s, i, os);
      fe8 ke, y, m;
      Fp<F>::mul_ctab_lazy(ke, g.key, e);
      Fp<F>::add_raw(y, bm, ke);
      Fp<F>::mul2_lazy(m, d, y, e, am);
      Fp<F>::csub_p(m);
      Fp<F>::add(om, m, cm);
      st_fe(g.out_m, i, om);
    }
  }
  if (stamps) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      stamps[3 * blockIdx.x] = t0; stamps[3 * blockIdx.x + 1] = gtime(); stamps[3 * blockIdx.x + 2] = smid;
    }
  }
}

template <class F, int BLK, int MINB, int MODE>
void run(const char* name, size_t n, RecombineArgs g, unsigned long long* stamps_dev) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  unsigned grid = (unsigned)((n + BLK - 1) / BLK);
  for (int i = 0; i < 3; i++) k2t<F, BLK, MINB, MODE><<<grid, BLK>>>(n, g, nullptr);
  cudaDeviceSynchronize();
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k2t<F, BLK, MINB, MODE><<<grid, BLK>>>(n, g, nullptr);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<uint64_t> h(n * 4);
  cudaMemcpy(h.data(), g.out_m.p, n * 32, cudaMemcpyDeviceToHost);
  uint64_t s = 0; for (size_t i = 0; i < n * 4; i++) s = s * 1000003u + h[i];
  printf("%-34s blk=%d minb=%d grid=%u  %8.2f us/launch  frac=%.3f chk=%016llx %s\n", name, BLK, MINB, grid, 1e3 * ms / reps,
         384.0 * n / (1e-3 * ms / reps) / 1e9 / 6549.8, (unsigned long long)s, cudaGetErrorString(cudaGetLastError()));
  // one instrumented launch
  cudaMemset(stamps_dev, 0, (size_t)grid * 24);
  k2t<F, BLK, MINB, MODE><<<grid, BLK>>>(n, g, stamps_dev);
  cudaDeviceSynchronize();
  std::vector<unsigned long long> st((size_t)grid * 3);
  cudaMemcpy(st.data(), stamps_dev, (size_t)grid * 24, cudaMemcpyDeviceToHost);
  unsigned long long tmin = ~0ull, tmax = 0;
  for (unsigned b = 0; b < grid; b++) { tmin = std::min(tmin, st[3 * b]); tmax = std::max(tmax, st[3 * b + 1]); }
  const double total = (tmax - tmin) * 1e-3;
  // blocks resident over time in 2 us buckets; mean duration of blocks started in each bucket
  const int nb = (int)(total / 2.0) + 1;
  std::vector<double> resid(nb, 0), dur(nb, 0); std::vector<int> started(nb, 0), ended(nb, 0);
  for (unsigned b = 0; b < grid; b++) {
    double s0 = (st[3 * b] - tmin) * 1e-3, s1 = (st[3 * b + 1] - tmin) * 1e-3;
    int b0 = std::min(nb - 1, (int)(s0 / 2.0)), b1 = std::min(nb - 1, (int)(s1 / 2.0));
    started[b0]++; ended[b1]++; dur[b0] += s1 - s0;
    for (int k = b0; k <= b1; k++) { double lo = std::max(s0, 2.0 * k), hi = std::min(s1, 2.0 * (k + 1)); if (hi > lo) resid[k] += (hi - lo) / 2.0; }
  }
  printf("  instrumented launch: first block start -> last block end %.2f us; per 2-us bucket: [t] resident blocks | started | ended | mean duration of blocks started\n", total);
  for (int k = 0; k < nb; k++) printf("   [%5.1f] %6.1f | %4d | %4d | %6.2f\n", 2.0 * k, resid[k], started[k], ended[k], started[k] ? dur[k] / started[k] : 0.0);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  const size_t n = 1u << 20;
  char* buf[12];
  for (int k = 0; k < 12; k++) cudaMalloc(&buf[k], n * 32);
  for (int k = 0; k < 10; k++) fr_random_kernel<Bn254Fr><<<sms * 8, kBlock>>>(n, 1000 + k, 0, MVec{buf[k], 32});
  RecombineArgs g;
  g.d_mine = Vec{buf[0], 32}; g.e_mine = Vec{buf[1], 32}; g.d_peer = Vec{buf[2], 32}; g.e_peer = Vec{buf[3], 32};
  g.a_s = Vec{buf[4], 32}; g.a_m = Vec{buf[5], 32}; g.b_s = Vec{buf[6], 32}; g.b_m = Vec{buf[7], 32}; g.c_s = Vec{buf[8], 32}; g.c_m = Vec{buf[9], 32};
  g.out_s = MVec{buf[10], 32}; g.out_m = MVec{buf[11], 32}; g.d_open = MVec{nullptr, 32}; g.e_open = MVec{nullptr, 32}; g.independent = 0;
  const uint64_t key[4] = {0x123456789abcdef1ull, 0x0fedcba987654321ull, 0x1122334455667788ull, 0x0123456789abcdefull};
  ctab_build<Bn254Fr>(g.key, key);
  unsigned long long* stamps; cudaMalloc(&stamps, (size_t)(n / 64 + 1) * 24);
  run<Bn254Fr, 256, 3, 0>("all loads first 256x3", n, g, stamps);
  run<Bn254Fr, 128, 6, 0>("all loads first 128x6", n, g, stamps);
  run<Bn254Fr, 256, 3, 1>("mac operands late 256x3", n, g, stamps);
  run<Bn254Fr, 256, 4, 1>("mac operands late 256x4 (64 regs)", n, g, stamps);
  run<Bn254Fr, 128, 8, 1>("mac operands late 128x8 (64 regs)", n, g, stamps);
  return 0;
}
