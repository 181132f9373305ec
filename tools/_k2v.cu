// Scratch A/B harness: occupancy variants of the fused recombine kernel (block size x min blocks/SM), same arithmetic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/_k2v tools/_k2v.cu
#include <cstdio>
#include <vector>
#include "../ark_mpc_b200/csrc/ctab.hpp"
#include "../ark_mpc_b200/csrc/fr_kernels.cuh"
using namespace ark;
template <class F, int PARTY, int BLK, int MINB>
__global__ void __launch_bounds__(BLK, MINB) k2v(size_t n, const __grid_constant__ RecombineArgs g) {
  const size_t step = (size_t)gridDim.x * BLK;
  for (size_t i = (size_t)blockIdx.x * BLK + threadIdx.x; i < n; i += step) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i); ld_fe(dp, g.d_peer, i); ld_fe(em, g.e_mine, i); ld_fe(ep, g.e_peer, i);
    ld_fe(bs, g.b_s, i); ld_fe(as, g.a_s, i); ld_fe(bm, g.b_m, i); ld_fe(am, g.a_m, i); ld_fe(cs, g.c_s, i); ld_fe(cm, g.c_m, i);
    fe8 os, om, d, e;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    st_fe(g.out_s, i, os); st_fe(g.out_m, i, om);
  }
}
template <class F, int BLK, int MINB>
float run(const char* name, size_t n, RecombineArgs g, int sms, int blocks_per_sm, uint64_t* sum_out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  size_t need = (n + BLK - 1) / BLK; size_t cap = (size_t)sms * blocks_per_sm; unsigned grid = (unsigned)(need < cap ? need : cap);
  for (int i = 0; i < 3; i++) k2v<F, 0, BLK, MINB><<<grid, BLK>>>(n, g);
  cudaDeviceSynchronize();
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) k2v<F, 0, BLK, MINB><<<grid, BLK>>>(n, g);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  std::vector<uint64_t> h(n * 4);
  cudaMemcpy(h.data(), g.out_m.p, n * 32, cudaMemcpyDeviceToHost);
  uint64_t s = 0; for (size_t i = 0; i < n * 4; i++) s = s * 1000003u + h[i];
  *sum_out = s;
  printf("%-28s blk=%d minb=%d grid=%u  %8.2f us/launch  frac=%.3f  chk=%016llx %s\n", name, BLK, MINB, grid, 1e3 * ms / reps,
         384.0 * n / (1e-3 * ms / reps) / 1e9 / 6549.8, (unsigned long long)s, err == cudaSuccess ? "" : cudaGetErrorString(err));
  return ms / reps;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
  const size_t n = 1u << 20;
  char* buf[12];
  for (int k = 0; k < 12; k++) { cudaMalloc(&buf[k], n * 32); }
  for (int k = 0; k < 10; k++) fr_random_kernel<Bn254Fr><<<sms * 8, kBlock>>>(n, 1000 + k, 0, MVec{buf[k], 32});
  RecombineArgs g;
  g.d_mine = Vec{buf[0], 32}; g.e_mine = Vec{buf[1], 32}; g.d_peer = Vec{buf[2], 32}; g.e_peer = Vec{buf[3], 32};
  g.a_s = Vec{buf[4], 32}; g.a_m = Vec{buf[5], 32}; g.b_s = Vec{buf[6], 32}; g.b_m = Vec{buf[7], 32}; g.c_s = Vec{buf[8], 32}; g.c_m = Vec{buf[9], 32};
  g.out_s = MVec{buf[10], 32}; g.out_m = MVec{buf[11], 32}; g.d_open = MVec{nullptr, 32}; g.e_open = MVec{nullptr, 32}; g.independent = 0;
  const uint64_t key[4] = {0x123456789abcdef1ull, 0x0fedcba987654321ull, 0x1122334455667788ull, 0x0123456789abcdefull};
  ctab_build<Bn254Fr>(g.key, key);
  uint64_t s;
  for (int rep = 0; rep < 2; rep++) {
    run<Bn254Fr, 256, 3>("occ3 full grid (current)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 160, 5>("b160x5 (800 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 416, 2>("b416x2 (832 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 224, 3>("b224x3 (672 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 128, 6>("b128x6 (768 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 96, 8>("b96x8 (768 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Bn254Fr, 192, 4>("b192x4 (768 thr/SM)", n, g, sms, 1 << 20, &s);
    run<Curve25519Fr, 256, 3>("c25519 occ3 (current)", n, g, sms, 1 << 20, &s);
    run<Curve25519Fr, 160, 5>("c25519 b160x5", n, g, sms, 1 << 20, &s);
  }
  return 0;
}
