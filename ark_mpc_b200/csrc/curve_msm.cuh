// Public multiscalar multiplication sum_i s_i * P_i by the bucket method (Pippenger), replacing the reference's call into
// ark-ec's `VariableBaseMSM` (`CurvePoint::msm`, /root/reference/online-phase/src/algebra/curve/curve.rs:549-560, used by
// msm_results / msm_authenticated :587-642).  The value is a group element, so any evaluation order gives the same affine result.
//
// With windows of c bits (W = ceil(256 / c) windows):
//   1. msm_count    histogram of the non-zero window digits of every scalar            (n threads, atomics on W * 2^c counters)
//   2. msm_scan     exclusive prefix sum of the histogram                               (one block)
//   3. msm_scatter  counting sort: point indices grouped by (window, digit)             (n threads)
//   4. msm_bucket   one thread per bucket adds its points; over-full buckets go to msm_bigbucket (one block per bucket)
//   5. msm_chunk    one thread per run of 32 buckets: sum_b b*B_b by running sums plus one small scalar multiple
//   6. msm_window   one block per window folds its chunks and scales by 2^(c*w); a final point sum adds the W windows
// ~n*W + 3*W*2^c point additions instead of n full scalar multiplications (~2600 field multiplications each).
#pragma once
#include "curve_kernels.cuh"

namespace ark {

constexpr int kMsmChunk = 32;       // buckets per chunk thread
constexpr int kMsmScanThreads = 1024;

__device__ __forceinline__ uint32_t msm_digit(const uint32_t* k, int w, int c) {
  const int bit = w * c;
  const int limb = bit >> 5, sh = bit & 31;
  uint64_t v = k[limb];
  if (limb + 1 < 8) v |= (uint64_t)k[limb + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << c) - 1u);  // c <= 16 and sh <= 31: the 64-bit word always covers the window
}

template <class C>
__global__ void __launch_bounds__(kBlock) msm_count_kernel(size_t n, Vec s, int c, int W, uint32_t* counts) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x;
    uint32_t k[8];
    ld_fe(x, s, i);
    scalar_to_plain<typename C::R>(k, x);
    for (int w = 0; w < W; w++) {
      const uint32_t d = msm_digit(k, w, c);
      if (d) atomicAdd(&counts[((size_t)w << c) | d], 1u);
    }
  }
}

// offsets[j] = sum_{i<j} counts[i]; cursors cleared.  One block.
static __global__ void __launch_bounds__(kMsmScanThreads) msm_scan_kernel(size_t m, const uint32_t* counts, uint32_t* offsets, uint32_t* cursors) {
  __shared__ uint32_t part[kMsmScanThreads];
  const size_t per = (m + kMsmScanThreads - 1) / kMsmScanThreads;
  const size_t lo = (size_t)threadIdx.x * per, hi = lo + per < m ? lo + per : m;
  uint32_t sum = 0;
  for (size_t j = lo; j < hi; j++) sum += counts[j];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < kMsmScanThreads; off <<= 1) {  // inclusive Hillis-Steele scan of the per-thread totals
    uint32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = threadIdx.x ? part[threadIdx.x - 1] : 0;
  for (size_t j = lo; j < hi; j++) {
    offsets[j] = run;
    cursors[j] = 0;
    run += counts[j];
  }
}

template <class C>
__global__ void __launch_bounds__(kBlock) msm_scatter_kernel(size_t n, Vec s, int c, int W, const uint32_t* offsets, uint32_t* cursors, uint32_t* idx) {
  const size_t step = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += step) {
    fe8 x;
    uint32_t k[8];
    ld_fe(x, s, i);
    scalar_to_plain<typename C::R>(k, x);
    for (int w = 0; w < W; w++) {
      const uint32_t d = msm_digit(k, w, c);
      if (d) {
        const size_t key = ((size_t)w << c) | d;
        idx[offsets[key] + atomicAdd(&cursors[key], 1u)] = (uint32_t)i;
      }
    }
  }
}

// bucket[key] = sum of the points whose digit in window (key >> c) is (key & mask).  One thread per bucket; buckets with more
// than kMsmBigBucket points (the top window of a scalar field whose bit length is not a multiple of c has few, very full
// buckets: a third of all BN254 scalars share bit 253) are queued for msm_bigbucket_kernel, where a whole block sums each.
constexpr uint32_t kMsmBigBucket = 256;

template <class C>
__global__ void __launch_bounds__(kPtBlock) msm_bucket_kernel(size_t m, const uint32_t* offsets, const uint32_t* counts, const uint32_t* idx, PVec pts,
                                                             PMVec buckets, uint32_t* biglist /* [0] = count, then keys */) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t key = (size_t)blockIdx.x * kPtBlock + threadIdx.x; key < m; key += step) {
    const uint32_t off = offsets[key], cnt = counts[key];
    if (cnt > kMsmBigBucket) {
      biglist[1 + atomicAdd(&biglist[0], 1u)] = (uint32_t)key;
      continue;
    }
    typename C::Pt acc;
    C::set_identity(acc);
#pragma unroll 1
    for (uint32_t j = 0; j < cnt; j++) {
      typename C::Pt x;
      ld_pt<C>(x, pts, idx[off + j]);
      C::add(acc, x);
    }
    st_pt<C>(buckets, key, acc);
  }
}

template <class C>
__global__ void __launch_bounds__(kPtBlock) msm_bigbucket_kernel(const uint32_t* offsets, const uint32_t* counts, const uint32_t* idx, PVec pts, PMVec buckets,
                                                                const uint32_t* biglist) {
  __shared__ typename C::Pt part[kPtBlock / 32];
  const uint32_t nbig = biglist[0];
  for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
    const uint32_t key = biglist[1 + b];
    const uint32_t off = offsets[key], cnt = counts[key];
    typename C::Pt acc;
    C::set_identity(acc);
#pragma unroll 1
    for (uint32_t j = threadIdx.x; j < cnt; j += kPtBlock) {
      typename C::Pt x;
      ld_pt<C>(x, pts, idx[off + j]);
      C::add(acc, x);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
      typename C::Pt t;
      pt_shfl_down<C>(t, acc, o);
      C::add(acc, t);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < kPtBlock / 32; k++) C::add(acc, part[k]);
      st_pt<C>(buckets, key, acc);
    }
    __syncthreads();
  }
}

// partial[w][j] = sum_{b in chunk j} b * B_b.  With lo = j*32 (chunk covers digits lo+1 .. lo+32, digit 0 unused) and the running
// sums T = sum B_b, S = sum (b - lo) B_b accumulated from the top digit down, the chunk value is S + lo * T.
template <class C>
__global__ void __launch_bounds__(kPtBlock) msm_chunk_kernel(int c, int W, PVec buckets, PMVec partials) {
  const size_t chunks = ((size_t)1 << c) / kMsmChunk ? ((size_t)1 << c) / kMsmChunk : 1;
  const size_t total = chunks * W;
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t t = (size_t)blockIdx.x * kPtBlock + threadIdx.x; t < total; t += step) {
    const size_t w = t / chunks, j = t % chunks;
    const uint32_t lo = (uint32_t)j * kMsmChunk;
    const uint32_t width = ((uint32_t)1 << c) < (uint32_t)kMsmChunk ? ((uint32_t)1 << c) : (uint32_t)kMsmChunk;
    typename C::Pt T, S;
    C::set_identity(T);
    C::set_identity(S);
#pragma unroll 1
    for (int b = (int)width; b >= 1; b--) {
      const uint32_t digit = lo + (uint32_t)b;
      if (digit < ((uint32_t)1 << c)) {
        typename C::Pt B;
        ld_pt<C>(B, buckets, (w << c) | digit);
        C::add(T, B);
      }
      C::add(S, T);
    }
    // lo * T by double-and-add (lo < 2^c <= 2^16)
    typename C::Pt L;
    C::set_identity(L);
#pragma unroll 1
    for (int bit = 15; bit >= 0; bit--) {
      C::dbl(L);
      if ((lo >> bit) & 1u) C::add(L, T);
    }
    C::add(S, L);
    st_pt<C>(partials, t, S);
  }
}

// out[w] = 2^(c*w) * sum_j partial[w][j]; one block per window
template <class C>
__global__ void __launch_bounds__(kPtBlock) msm_window_kernel(int c, size_t chunks, PVec partials, PMVec out) {
  __shared__ typename C::Pt part[kPtBlock / 32];
  const size_t w = blockIdx.x;
  typename C::Pt acc;
  C::set_identity(acc);
  for (size_t j = threadIdx.x; j < chunks; j += kPtBlock) {
    typename C::Pt x;
    ld_pt<C>(x, partials, w * chunks + j);
    C::add(acc, x);
  }
#pragma unroll 1
  for (int off = 16; off > 0; off >>= 1) {
    typename C::Pt o;
    pt_shfl_down<C>(o, acc, off);
    C::add(acc, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < kPtBlock / 32; k++) C::add(acc, part[k]);
    for (size_t i = 0; i < (size_t)c * w; i++) C::dbl(acc);
    st_pt<C>(out, w, acc);
  }
}

}  // namespace ark
