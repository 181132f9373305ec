// Scratch: does programmatic dependent launch (PDL) trim the gaps between the four kernels of a two-party Beaver step?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o tools/_pdl tools/_pdl.cu
#include <cstdio>
#include <vector>
#include "../ark_mpc_b200/csrc/fr_kernels.cuh"
using namespace ark;

template <class F, bool PDL>
__global__ void __launch_bounds__(kBlock) k1(size_t n, Vec x, Vec y, Vec a, Vec b, MVec d, MVec e) {
  if (PDL) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 xs, ys, as, bs, dm, em;
    ld_fe(xs, x, i); ld_fe(ys, y, i); ld_fe(as, a, i); ld_fe(bs, b, i);
    beaver_mask_elem<F>(dm, em, xs, ys, as, bs);
    st_fe(d, i, dm); st_fe(e, i, em);
  }
}
template <class F, int PARTY, bool PDL>
__global__ void __launch_bounds__(kBlock, 3) k2(size_t n, const __grid_constant__ RecombineArgs g) {
  if (PDL) { asm volatile("griddepcontrol.launch_dependents;"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 dm, em, dp, ep, as, am, bs, bm, cs, cm;
    ld_fe(dm, g.d_mine, i); ld_fe(dp, g.d_peer, i); ld_fe(em, g.e_mine, i); ld_fe(ep, g.e_peer, i);
    ld_fe(bs, g.b_s, i); ld_fe(as, g.a_s, i); ld_fe(bm, g.b_m, i); ld_fe(am, g.a_m, i); ld_fe(cs, g.c_s, i); ld_fe(cm, g.c_m, i);
    fe8 os, om, d, e;
    beaver_recombine_elem<F>(os, om, d, e, PARTY, g.key, dm, em, dp, ep, as, am, bs, bm, cs, cm);
    st_fe(g.out_s, i, os); st_fe(g.out_m, i, om);
  }
}

template <class... Args>
void launch(void (*kern)(Args...), unsigned grid, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kBlock); cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);
}

int main() {
  const size_t n = 1u << 20;
  using F = Bn254Fr;
  char* buf[2][16];
  for (int p = 0; p < 2; p++) for (int k = 0; k < 16; k++) { cudaMalloc(&buf[p][k], n * 32); if (k < 8) fr_random_kernel<F><<<1184, kBlock>>>(n, 100 * p + k, 0, MVec{buf[p][k], 32}); }
  // per party: 0 x 1 y 2 a_s 3 a_m 4 b_s 5 b_m 6 c_s 7 c_m 8 d 9 e 10 out_s 11 out_m
  RecombineArgs g[2];
  for (int p = 0; p < 2; p++) {
    g[p].d_mine = Vec{buf[p][8], 32}; g[p].e_mine = Vec{buf[p][9], 32}; g[p].d_peer = Vec{buf[1 - p][8], 32}; g[p].e_peer = Vec{buf[1 - p][9], 32};
    g[p].a_s = Vec{buf[p][2], 32}; g[p].a_m = Vec{buf[p][3], 32}; g[p].b_s = Vec{buf[p][4], 32}; g[p].b_m = Vec{buf[p][5], 32};
    g[p].c_s = Vec{buf[p][6], 32}; g[p].c_m = Vec{buf[p][7], 32}; g[p].out_s = MVec{buf[p][10], 32}; g[p].out_m = MVec{buf[p][11], 32};
    g[p].d_open = MVec{nullptr, 32}; g[p].e_open = MVec{nullptr, 32};
    for (int j = 0; j < 8; j++) g[p].key.v[j] = 0x1234567u * (j + 1 + p);
    g[p].key.v[7] &= 0x0fffffffu;
  }
  const unsigned grid = (unsigned)(n / kBlock);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; rep++)
    for (int pdl = 0; pdl < 2; pdl++) {
      auto step = [&]() {
        for (int p = 0; p < 2; p++) {
          if (pdl) launch(k1<F, true>, grid, true, n, Vec{buf[p][0], 32}, Vec{buf[p][1], 32}, Vec{buf[p][2], 32}, Vec{buf[p][4], 32}, MVec{buf[p][8], 32}, MVec{buf[p][9], 32});
          else launch(k1<F, false>, grid, false, n, Vec{buf[p][0], 32}, Vec{buf[p][1], 32}, Vec{buf[p][2], 32}, Vec{buf[p][4], 32}, MVec{buf[p][8], 32}, MVec{buf[p][9], 32});
        }
        if (pdl) { launch(k2<F, 0, true>, grid, true, n, g[0]); launch(k2<F, 1, true>, grid, true, n, g[1]); }
        else { launch(k2<F, 0, false>, grid, false, n, g[0]); launch(k2<F, 1, false>, grid, false, n, g[1]); }
      };
      for (int i = 0; i < 20; i++) step();
      cudaDeviceSynchronize();
      const int steps = 500;
      cudaEventRecord(e0);
      for (int i = 0; i < steps; i++) step();
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<uint64_t> h(n * 4);
      cudaMemcpy(h.data(), buf[0][11], n * 32, cudaMemcpyDeviceToHost);
      uint64_t s = 0; for (size_t i = 0; i < n * 4; i++) s = s * 1000003u + h[i];
      printf("%s  %8.2f us/step  %.3f G mults/s  chk=%016llx  %s\n", pdl ? "PDL      " : "plain    ", 1e3 * ms / steps, n / (ms / steps) / 1e6, (unsigned long long)s, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
