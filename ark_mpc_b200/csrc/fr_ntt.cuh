// Batch inversion and radix-2 NTT over the scalar field: the other element-parallel callers of the hot path
// (SURVEY §8f rank 3).  Reference behaviour restated (paths under /root/reference/online-phase/src/algebra/scalar):
//   Scalar::batch_inverse            scalar.rs:93-100  -> ark_ff::batch_inversion (zeros stay zero)
//   ScalarShare::fft_helper          share.rs:162-192  -> ark-poly `EvaluationDomain::fft / ifft` on the share and mac planes
//   AuthenticatedScalarResult::fft   authenticated_scalar.rs:1011-1070
// ark-poly's Radix2EvaluationDomain of size n = 2^k evaluates X_j = sum_i x_i w^(ij) with w = TWO_ADIC_ROOT_OF_UNITY^(2^(s-k))
// (s = 28, generator 5 for BN254 Fr) and returns the result in natural order; ifft is its inverse (w^-1, scaled by n^-1).
// The transform is a function of w alone, so any correct algorithm is bit-identical on canonical residues.
#pragma once
#include "curve.cuh"  // Fq<>::inv: Fermat inversion for any of the Montgomery fields
#include "fr_kernels.cuh"

namespace ark {

// K selects the Karatsuba product (Fp::mul_kara: 112 wide multiply-adds) over the interleaved one (Fp::mul: 128)
template <class F, bool K>
__device__ __forceinline__ void fmul(fe8& r, const fe8& a, const fe8& b) {
  if (K) Fp<F>::mul_kara(r, a, b); else Fp<F>::mul(r, a, b);
}

// ---------------------------------------------------------------------------------------------
// Batch inversion (scalar.rs:93-100 -> ark_ff::batch_inversion: zeros stay zero): Montgomery's trick as a product TREE.
//   up    every thread multiplies a group of kInvGroup = 8 elements (element j of group g is x[g + j*groups]: coalesced) as a
//         balanced binary tree — 4 + 2 + 1 multiplications, depth 3 — keeps the six inner products and hands the group product
//         to the next level; levels shrink by 8 until <= kInvTop remain.  Zeros (and the padding of a ragged last group) enter
//         the product as one;
//   top   one safegcd inversion per remaining element (Fp::inv_mont): up to 16 k of them cost the latency of one, a warp per
//         SM sub-partition;
//   down  every thread walks its group's tree down: the inverse of a node times its sibling's product is the inverse of the
//         other child — 2 + 4 + 8 independent multiplications.
// 21 multiplications per 8 elements with dependent depth 3 + 3 (the serial prefix form: 24, depth 7 + 8, and at 12 % occupancy
// nothing hides that chain), and ONE inversion latency on the critical path — round 1 ran a 380-multiplication Fermat chain per
// 32 elements on n/32 threads: 320 us at n = 2^16 and at 2^20.
// ---------------------------------------------------------------------------------------------
constexpr int kInvGroup = 8;
constexpr size_t kInvTop = 16384;
constexpr int kInvBlock = 128;  // upper bound (launch bounds); the sweeps are launched with 64-thread blocks: at n = 2^20 the down sweep's 2^17
                                 // threads at 134 registers are 1.98 waves of 7 blocks per SM (128-thread blocks: 2.31 waves of 3, a third of the last one idle)
constexpr int kInvLaunchBlock = 64;

// loads the eight elements of group g (all loads issued before the first use), replaces zeros and out-of-range slots by one and
// returns the mask of the slots that hold a non-zero input
template <class F>
__device__ __forceinline__ uint32_t inv_load_group(fe8 (&z)[kInvGroup], size_t n, size_t groups, size_t g, Vec x) {
#pragma unroll
  for (int j = 0; j < kInvGroup; j++) {
    const size_t i = g + (size_t)j * groups;
    if (i < n) ld_fe(z[j], x, i); else Fp<F>::set_zero(z[j]);
  }
  uint32_t live = 0;
#pragma unroll
  for (int j = 0; j < kInvGroup; j++) {
    if (Fp<F>::is_zero(z[j])) Fp<F>::set_one(z[j]); else live |= 1u << j;
  }
  return live;
}

template <class F, bool K>
__global__ void __launch_bounds__(kInvBlock) fr_inv_up_kernel(size_t n, size_t groups, Vec x, MVec tree, MVec prod) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  fe8 z[kInvGroup];
  (void)inv_load_group<F>(z, n, groups, g, x);
  // tree[k * groups + g]: k = 0..3 the pair products, 4..5 the products of four
#pragma unroll
  for (int k = 0; k < 4; k++) {
    fmul<F, K>(z[2 * k], z[2 * k], z[2 * k + 1]);
    st_fe(tree, (size_t)k * groups + g, z[2 * k]);
  }
  fmul<F, K>(z[0], z[0], z[2]);
  fmul<F, K>(z[4], z[4], z[6]);
  st_fe(tree, 4 * groups + g, z[0]);
  st_fe(tree, 5 * groups + g, z[4]);
  fmul<F, K>(z[0], z[0], z[4]);
  st_fe(prod, g, z[0]);
}

// Cooperative sweeps for the SMALL levels of the tree (a few thousand groups): with one thread per group such a launch is a
// lone warp per scheduler running 7 (up) or 14 (down) multiplications back to back at ~1 us each — 8.5 + 13.3 us at 2^20 for
// an eighth of the data.  Here four threads share a group: thread (k, g) of a 256-thread block multiplies pair k of group g, the
// levels of the tree meet in shared memory, and the dependent depth is 3 multiplications up and 1 + 2 down: 7.5 + 10.5 us
// (profiles/r02z4_summary.txt; what remains is load latency and the ~1 us a lone warp needs per multiplication).  Doing the
// last up sweep, the inversions and the first down sweep in ONE kernel, a thread per group, was tried and is slower above 2^16
// elements (105 against 98 us at 2^20, profiles/r02z2_summary.txt): the 21 multiplications then run on a lone warp too.
constexpr int kInvCoopGroups = 64;               // groups per block
constexpr int kInvCoopBlock = 4 * kInvCoopGroups;
constexpr size_t kInvCoopMaxGroups = 32768;      // above this the launch is throughput-bound and one thread per group is cheaper

template <class F>
__device__ __forceinline__ uint32_t inv_load_pair(fe8& z0, fe8& z1, size_t n, size_t groups, size_t g, int k, Vec x) {
  const size_t i0 = g + (size_t)(2 * k) * groups, i1 = i0 + groups;
  if (i0 < n) ld_fe(z0, x, i0); else Fp<F>::set_zero(z0);
  if (i1 < n) ld_fe(z1, x, i1); else Fp<F>::set_zero(z1);
  uint32_t live = 0;
  if (Fp<F>::is_zero(z0)) Fp<F>::set_one(z0); else live |= 1u;
  if (Fp<F>::is_zero(z1)) Fp<F>::set_one(z1); else live |= 2u;
  return live;
}

template <class F, bool K>
__global__ void __launch_bounds__(kInvCoopBlock) fr_inv_up_coop_kernel(size_t n, size_t groups, Vec x, MVec tree, MVec prod) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  __shared__ __align__(32) fe8 sp[4][kInvCoopGroups];
  __shared__ __align__(32) fe8 sq[2][kInvCoopGroups];
  const int k = threadIdx.x / kInvCoopGroups, gl = threadIdx.x % kInvCoopGroups;
  const size_t g = (size_t)blockIdx.x * kInvCoopGroups + gl;
  const bool act = g < groups;
  if (act) {
    fe8 z0, z1, p;
    (void)inv_load_pair<F>(z0, z1, n, groups, g, k, x);
    fmul<F, K>(p, z0, z1);
    st_fe(tree, (size_t)k * groups + g, p);
    sp[k][gl] = p;
  }
  __syncthreads();
  if (act && k < 2) {
    fe8 q;
    fmul<F, K>(q, sp[2 * k][gl], sp[2 * k + 1][gl]);
    st_fe(tree, (size_t)(4 + k) * groups + g, q);
    sq[k][gl] = q;
  }
  __syncthreads();
  if (act && k == 0) {
    fe8 t;
    fmul<F, K>(t, sq[0][gl], sq[1][gl]);
    st_fe(prod, g, t);
  }
}

template <class F, bool K>
__global__ void __launch_bounds__(kInvCoopBlock) fr_inv_down_coop_kernel(size_t n, size_t groups, Vec x, Vec tree, Vec ginv, MVec out) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  __shared__ __align__(32) fe8 siq[2][kInvCoopGroups];
  const int k = threadIdx.x / kInvCoopGroups, gl = threadIdx.x % kInvCoopGroups;
  const size_t g = (size_t)blockIdx.x * kInvCoopGroups + gl;
  const bool act = g < groups;
  fe8 z0, z1, ps;
  uint32_t live = 0;
  if (act) {  // everything this thread needs from memory is requested before the first multiplication
    live = inv_load_pair<F>(z0, z1, n, groups, g, k, x);
    ld_fe(ps, tree, (size_t)(k ^ 1) * groups + g);  // the sibling pair's product
  }
  if (act && k < 2) {
    fe8 inv, qo, iq;
    ld_fe(inv, ginv, g);
    ld_fe(qo, tree, (size_t)(4 + (k ^ 1)) * groups + g);
    fmul<F, K>(iq, inv, qo);
    siq[k][gl] = iq;
  }
  __syncthreads();
  if (act) {
    fe8 ip, r;
    fmul<F, K>(ip, siq[k >> 1][gl], ps);
    const size_t i0 = g + (size_t)(2 * k) * groups, i1 = i0 + groups;
    if (i0 < n) {
      fmul<F, K>(r, ip, z1);
      if (!(live & 1u)) Fp<F>::set_zero(r);  // zeros stay zero
      st_fe(out, i0, r);
    }
    if (i1 < n) {
      fmul<F, K>(r, ip, z0);
      if (!(live & 2u)) Fp<F>::set_zero(r);
      st_fe(out, i1, r);
    }
  }
}

// One warp per block: the <= 16384 inversions at the top are latency-bound, so they are spread one warp per SM sub-partition
// (512 blocks over 148 SMs x 4 schedulers) and cost the latency of a single inversion.
constexpr int kInvTopBlock = 32;
template <class F>
__global__ void __launch_bounds__(kInvTopBlock) fr_inv_top_kernel(size_t n, Vec x, MVec out) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  const size_t i = (size_t)blockIdx.x * kInvTopBlock + threadIdx.x;
  if (i >= n) return;
  fe8 v, r;
  ld_fe(v, x, i);
  if (Fp<F>::is_zero(v)) r = v; else Fp<F>::inv_mont(r, v);
  st_fe(out, i, r);
}

template <class F, bool K>
__global__ void __launch_bounds__(kInvBlock) fr_inv_down_kernel(size_t n, size_t groups, Vec x, Vec tree, Vec ginv, MVec out) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  fe8 z[kInvGroup], p[4], q[2], inv;
  ld_fe(inv, ginv, g);
#pragma unroll
  for (int k = 0; k < 4; k++) ld_fe(p[k], tree, (size_t)k * groups + g);
  ld_fe(q[0], tree, 4 * groups + g);
  ld_fe(q[1], tree, 5 * groups + g);
  const uint32_t live = inv_load_group<F>(z, n, groups, g, x);
  fe8 iq[2], ip[4];
  fmul<F, K>(iq[0], inv, q[1]);
  fmul<F, K>(iq[1], inv, q[0]);
#pragma unroll
  for (int k = 0; k < 4; k++) fmul<F, K>(ip[k], iq[k >> 1], p[k ^ 1]);
#pragma unroll
  for (int j = 0; j < kInvGroup; j++) {
    const size_t i = g + (size_t)j * groups;
    if (i >= n) continue;
    fe8 r;
    fmul<F, K>(r, ip[j >> 1], z[j ^ 1]);
    if (!((live >> j) & 1u)) Fp<F>::set_zero(r);  // zeros stay zero
    st_fe(out, i, r);
  }
}

// ---------------------------------------------------------------------------------------------
// NTT
// ---------------------------------------------------------------------------------------------
// The 2^28-th root of unity of BN254 Fr that ark-ff fixes: TWO_ADIC_ROOT_OF_UNITY = 5^((p-1)/2^28)
// = 19103219067921713944291392827692070036145651957329286315305642004821462161904 (Montgomery image below).
struct NttRoot {
  static constexpr int kTwoAdicity = 28;
  __device__ __forceinline__ static void set(fe8& r) {
    r.v[0] = 0x80d13d9cu; r.v[1] = 0x636e7355u; r.v[2] = 0x2445ffd6u; r.v[3] = 0xa22bf374u;
    r.v[4] = 0x1eb203d8u; r.v[5] = 0x56452ac0u; r.v[6] = 0x2963f9e7u; r.v[7] = 0x1860ef94u;
  }
};

// consts[0] = w = root^(2^(28 - log2n)) (inverted for the inverse transform), consts[1] = (2^log2n)^-1.  One thread.
template <class F>
__global__ void fr_ntt_setup_kernel(int log2n, int inverse, fe8* consts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  fe8 w;
  NttRoot::set(w);
  for (int i = log2n; i < NttRoot::kTwoAdicity; i++) Fp<F>::mul(w, w, w);
  if (inverse) Fq<F>::inv(w, w);
  fe8 two, nn;
  Fp<F>::set_one(two);
  Fp<F>::add(two, two, two);
  Fp<F>::set_one(nn);
  for (int i = 0; i < log2n; i++) Fp<F>::mul(nn, nn, two);
  Fq<F>::inv(nn, nn);
  consts[0] = w;
  consts[1] = nn;
}

// out = a * (*s): the n^-1 scaling of the inverse transform, scalar taken from device memory
template <class F>
__global__ void __launch_bounds__(kBlock) fr_scale_dev_kernel(size_t n, Vec a, const fe8* s, MVec out) {
  const size_t step = (size_t)gridDim.x * kBlock;
  const fe8 k = *s;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x, r;
    ld_fe(x, a, i);
    Fp<F>::mul(r, x, k);
    st_fe(out, i, r);
  }
}

// tw[k] = w^k for k < half (left-to-right square-and-multiply on the bits of k)
template <class F>
__global__ void __launch_bounds__(kBlock) fr_ntt_twiddle_kernel(size_t half, const fe8* wp, MVec tw) {
  const size_t step = (size_t)gridDim.x * kBlock;
  const fe8 w = *wp;
  for (size_t k = (size_t)blockIdx.x * kBlock + threadIdx.x; k < half; k += step) {
    fe8 acc;
    Fp<F>::set_one(acc);
    for (int b = 63 - __clzll((unsigned long long)(k | 1)); b >= 0; b--) {
      Fp<F>::mul(acc, acc, acc);
      if ((k >> b) & 1) Fp<F>::mul(acc, acc, w);
    }
    st_fe(tw, k, acc);
  }
}

constexpr int kNttTileLog = 10;               // 1024 elements (32 KiB) per block in shared memory
constexpr int kNttTile = 1 << kNttTileLog;
constexpr int kNttThreads = kNttTile / 4;     // one radix-4 group (two stages of two butterflies) per thread per double stage

template <class F, bool K>
__device__ __forceinline__ void ntt_butterfly(fe8& lo, fe8& hi, const fe8& w) {
  fe8 v, s, d;
  fmul<F, K>(v, hi, w);
  Fp<F>::add(s, lo, v);
  Fp<F>::sub(d, lo, v);
  lo = s;
  hi = d;
}

// Stages 1..min(log2n, 10): a bit-reversed gather of one tile into shared memory, the butterflies in place, a coalesced store.
// Twiddle of butterfly j in stage s: w^(j * n / 2^s) = tw[j << (log2n - s)]; the tile's 2^tile_log-th roots of unity are staged in
// shared memory once.  Stages are taken two at a time: a thread owns the four elements {b, b + h, b + 2h, b + 3h} (h = 2^(s-1)),
// runs the two butterflies of stage s and the two of stage s + 1 in registers, and the block synchronises once per PAIR of
// stages — half the barriers and half the shared-memory round trips of one butterfly per thread per stage.
// The FIRST double stage never reads shared memory: a thread gathers its four bit-reversed inputs straight into registers, and
// three of its four twiddles are one (stage 1: w^0 twice; stage 2: w^0 and w^(n/4)), so it costs ONE multiplication instead of
// four — 7.5 % of the transform's multiplications.  The inverse transform folds its n^-1 in here (`scale`: three more
// multiplications per four elements) instead of a separate pass over the data.
template <class F, bool K>
__global__ void __launch_bounds__(kNttThreads) fr_ntt_tile_kernel(int log2n, Vec in, Vec tw, const fe8* __restrict__ scale, MVec out) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  extern __shared__ __align__(32) unsigned char ntt_smem[];
  __shared__ __align__(32) fe8 stw[kNttTile / 2];  // stw[k] = w^(k n / 2^tile_log)
  fe8* x = reinterpret_cast<fe8*>(ntt_smem);
  const int tile_log = log2n < kNttTileLog ? log2n : kNttTileLog;
  const size_t tile = (size_t)1 << tile_log;
  const size_t base = (size_t)blockIdx.x * tile;
  const int t = threadIdx.x;
  for (size_t k = t; k < tile / 2; k += blockDim.x) ld_fe(stw[k], tw, k << (log2n - tile_log));
  int s = 1;
  if (tile_log >= 2) {
    fe8 wi, c;
    ld_fe(wi, tw, (size_t)1 << (log2n - 2));  // w^(n/4)
    if (scale) {
      c = *scale;
      fmul<F, K>(wi, wi, c);
    }
    for (size_t q = t; q < tile / 4; q += blockDim.x) {
      fe8 e[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {  // out position p <- in[bitrev(p)]
        const size_t p = base + 4 * q + r;
        ld_fe(e[r], in, (size_t)(__brevll((unsigned long long)p) >> (64 - log2n)));
      }
      fe8 a, b, cc, d;
      Fp<F>::add(a, e[0], e[1]);
      Fp<F>::sub(b, e[0], e[1]);
      Fp<F>::add(cc, e[2], e[3]);
      Fp<F>::sub(d, e[2], e[3]);
      if (scale) {
        fmul<F, K>(a, a, c);
        fmul<F, K>(b, b, c);
        fmul<F, K>(cc, cc, c);
      }
      fmul<F, K>(d, d, wi);
      Fp<F>::add(e[0], a, cc);
      Fp<F>::sub(e[2], a, cc);
      Fp<F>::add(e[1], b, d);
      Fp<F>::sub(e[3], b, d);
#pragma unroll
      for (int r = 0; r < 4; r++) x[4 * q + r] = e[r];
    }
    s = 3;
  } else {
    // gather: out position p <- in[bitrev(p)]
    for (size_t e = t; e < tile; e += blockDim.x) {
      const size_t p = base + e;
      const size_t src = log2n ? (size_t)(__brevll((unsigned long long)p) >> (64 - log2n)) : 0;
      ld_fe(x[e], in, src);
    }
  }
  __syncthreads();
  for (; s + 1 <= tile_log; s += 2) {
    const size_t h = (size_t)1 << (s - 1);
    for (size_t q = t; q < tile / 4; q += blockDim.x) {
      const size_t j = q & (h - 1);
      const size_t b = ((q >> (s - 1)) << (s + 1)) + j;
      fe8 e0 = x[b], e1 = x[b + h], e2 = x[b + 2 * h], e3 = x[b + 3 * h];
      const fe8 w1 = stw[j << (tile_log - s)];                 // stage s: both butterflies use w^(j n / 2^s)
      ntt_butterfly<F, K>(e0, e1, w1);
      ntt_butterfly<F, K>(e2, e3, w1);
      ntt_butterfly<F, K>(e0, e2, stw[j << (tile_log - s - 1)]);         // stage s + 1: positions j and j + h
      ntt_butterfly<F, K>(e1, e3, stw[(j + h) << (tile_log - s - 1)]);
      x[b] = e0; x[b + h] = e1; x[b + 2 * h] = e2; x[b + 3 * h] = e3;
    }
    __syncthreads();
  }
  if (s <= tile_log) {  // an odd number of stages: the last one on its own
    const size_t h = (size_t)1 << (s - 1);
    for (size_t b2 = t; b2 < tile / 2; b2 += blockDim.x) {
      const size_t j = b2 & (h - 1);
      const size_t i0 = ((b2 >> (s - 1)) << s) + j;
      ntt_butterfly<F, K>(x[i0], x[i0 + h], stw[j << (tile_log - s)]);
    }
    __syncthreads();
  }
  for (size_t e = t; e < tile; e += blockDim.x) st_fe(out, base + e, x[e]);
}

// Stages s0+1 .. s0+T (s0 >= 10, T <= 5) in one pass: a block owns 32 consecutive low indices x 2^T strided positions
// {hi*2^(s0+T) + k*2^s0 + lo : k < 2^T, lo in a 32-wide window}, i.e. 2^T rows of 1 KiB that are each contiguous in memory, so
// loads and stores are fully coalesced; the T butterfly levels run in shared memory (32 KiB), one butterfly per thread per level.
constexpr int kNttStrideLog = 5;                         // up to 5 stages per strided pass
constexpr int kNttStrideThreads = 32 << (kNttStrideLog - 1);  // 512

template <class F, bool K>
__global__ void __launch_bounds__(kNttStrideThreads) fr_ntt_strided_kernel(int log2n, int s0, int T, Vec tw, MVec x) {
  pdl_prologue();  // launched with programmatic stream serialization: wait for the previous kernel of the chain, release the next
  __shared__ __align__(32) fe8 sm[32 << kNttStrideLog];  // [k][lo]
  const int lane = threadIdx.x & 31;
  const int row = threadIdx.x >> 5;                       // 0 .. 15
  const size_t lo_blocks = (size_t)1 << (s0 - 5);
  const size_t hi = blockIdx.x / lo_blocks;
  const size_t lo = (blockIdx.x % lo_blocks) * 32 + lane;
  const size_t base = (hi << (s0 + T)) + lo;
  const int rows = 1 << T;
  const Vec xr{x.p, x.stride};
  for (int k = row; k < rows; k += (blockDim.x >> 5)) ld_fe(sm[k * 32 + lane], xr, base + ((size_t)k << s0));
  __syncthreads();
  // One butterfly per thread per level when T = 5; the twiddle of the NEXT level is requested before this level's arithmetic so
  // that its L2 round trip overlaps the multiplication and the barrier (long-scoreboard was the top stall, profiles/r02i_*).
  auto tw_index = [&](int t, int b) -> size_t {
    const int kl = b & ((1 << (t - 1)) - 1);
    return (((size_t)kl << s0) + lo) << (log2n - (s0 + t));  // j << (log2n - s), j = i0 mod 2^(s-1)
  };
  fe8 w_next;
  if (row < rows / 2) ld_fe(w_next, tw, tw_index(1, row));
  for (int t = 1; t <= T; t++) {
    const int halfk = 1 << (t - 1);
    for (int b = row; b < rows / 2; b += (blockDim.x >> 5)) {
      fe8 w;
      if (b == row) w = w_next; else ld_fe(w, tw, tw_index(t, b));
      if (b == row && t < T) ld_fe(w_next, tw, tw_index(t + 1, row));
      const int kl = b & (halfk - 1);
      const int k = ((b >> (t - 1)) << t) + kl;
      ntt_butterfly<F, K>(sm[k * 32 + lane], sm[(k + halfk) * 32 + lane], w);
    }
    __syncthreads();
  }
  for (int k = row; k < rows; k += (blockDim.x >> 5)) st_fe(x, base + ((size_t)k << s0), sm[k * 32 + lane]);
}

}  // namespace ark
