// Kernel skeletons for the point gates.  Points are read and written in the reference's AoS memory image
// (BN254 G1Projective x,y,z = 96 B; Curve25519 EdwardsProjective x,y,t,z = 128 B; a PointShare is two consecutive
// points {share, mac}, curve/share.rs:25-30): every coordinate is 32 bytes and 32-byte aligned, so each is one
// LDG.256 / STG.256.  These kernels are compute-bound by three to four orders of magnitude (a scalar multiplication is
// ~3000 base-field multiplications per ~200 bytes), so the layout is chosen for a zero-copy boundary, not for
// coalescing.  Scalars stay planar as in fr_kernels.cuh.
#pragma once
#include "curve_gates.cuh"
#include "fr_kernels.cuh"

namespace ark {

constexpr int kPtBlock = 128;

struct PVec {
  const char* p;
  uint32_t stride;  // bytes between consecutive points (kPointBytes for point vectors, 2*kPointBytes inside PointShares)
};
struct PMVec {
  char* p;
  uint32_t stride;
};

template <class C>
__device__ __forceinline__ void ld_pt(typename C::Pt& r, const PVec& v, size_t i) {
  const Vec c{v.p + i * (size_t)v.stride, 32};
  fe8* f = reinterpret_cast<fe8*>(&r);
#pragma unroll
  for (int k = 0; k < C::kCoords; k++) ld_fe(f[k], c, k);
  C::from_image(r);
}
template <class C>
__device__ __forceinline__ void st_pt(const PMVec& v, size_t i, const typename C::Pt& r_in) {
  const MVec c{v.p + i * (size_t)v.stride, 32};
  typename C::Pt r = r_in;
  C::to_image(r);
  const fe8* f = reinterpret_cast<const fe8*>(&r);
#pragma unroll
  for (int k = 0; k < C::kCoords; k++) st_fe(c, k, f[k]);
}

// ---------------------------------------------------------------------------------------------
// Window tables of the variable-base multiplications: L2-resident scratch records instead of per-thread local arrays.
// Every resident 128-thread block claims one of kTabSlotsPerSm slots of its SM (an atomic bit mask per SM id), and each of its
// threads owns one contiguous kTabRecordBytes record in that slot: entries are written and read with 256-bit accesses, one
// 32-byte sector per lane per instruction, no over-fetch.  The scratch is (SM ids) x 6 x 128 x 4 KiB, is reused by every
// block that ever runs on the SM, and therefore stays in (or near) L2.
// ---------------------------------------------------------------------------------------------
constexpr int kTabSlotsPerSm = 6;
constexpr int kTabHalfBytes = kTabEntries * 128;     // one table: 8 entries of up to 128 B (Edwards cached form); BN254 uses 96 B of each
constexpr int kTabRecordBytes = kMaxSplitParts * kTabHalfBytes;  // the table of P and, for the two-pass gates, of the 2^(s p) P behind it

struct TabScratch {
  char* base;
  unsigned int* masks;  // one claim mask per SM id
};

// Table reads go through L1 (default caching): at 2 resident blocks per SM (BN254, 230 registers) nothing else hides an L2
// round trip in the middle of a dependent chain of point additions, and `prefetch` lets the variable-base loops request the
// entry of the NEXT window before the four doublings that precede its use.  A thread only ever reads records it wrote itself,
// so L1 residency needs no coherence beyond program order.
template <class C, bool AFF = C::kAffineTables>
struct GlobalTab {
  static constexpr bool kAffine = AFF;  // entries are affine once the tables are finished (curve.cuh: finish_tables, glv_add)
  char* rec;
  template <class T>
  __device__ __forceinline__ void put(int idx, const T& c) const {
    static_assert(sizeof(T) % 32 == 0 && sizeof(T) <= 128, "entries are whole 32-byte words of a 128-byte slot");
    const fe8* f = reinterpret_cast<const fe8*>(&c);
    char* a = rec + idx * 128;
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 32); k++)
      asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(f[k].v[0]), "r"(f[k].v[1]), "r"(f[k].v[2]), "r"(f[k].v[3]),
                   "r"(f[k].v[4]), "r"(f[k].v[5]), "r"(f[k].v[6]), "r"(f[k].v[7]), "l"(a + 32 * k)
                   : "memory");
  }
  template <class T>
  __device__ __forceinline__ void get(T& c, int idx) const {
    fe8* f = reinterpret_cast<fe8*>(&c);
    const char* a = rec + idx * 128;
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 32); k++)
      asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(f[k].v[0]), "=r"(f[k].v[1]), "=r"(f[k].v[2]), "=r"(f[k].v[3]), "=r"(f[k].v[4]), "=r"(f[k].v[5]), "=r"(f[k].v[6]),
                     "=r"(f[k].v[7])
                   : "l"(a + 32 * k)
                   : "memory");
  }
  // the last 32 bytes of a slot: free while the entry is a 96-byte Jacobian point (normalize_tables parks a prefix product there)
  __device__ __forceinline__ void put_aux(int idx, const fe8& v) const {
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v.v[0]), "r"(v.v[1]), "r"(v.v[2]), "r"(v.v[3]), "r"(v.v[4]),
                 "r"(v.v[5]), "r"(v.v[6]), "r"(v.v[7]), "l"(rec + idx * 128 + 96)
                 : "memory");
  }
  __device__ __forceinline__ void get_aux(fe8& v, int idx) const {
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.v[0]), "=r"(v.v[1]), "=r"(v.v[2]), "=r"(v.v[3]), "=r"(v.v[4]), "=r"(v.v[5]), "=r"(v.v[6]), "=r"(v.v[7])
                 : "l"(rec + idx * 128 + 96)
                 : "memory");
  }
  __device__ __forceinline__ void prefetch(int idx) const { asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + idx * 128)); }
  __device__ __forceinline__ GlobalTab sub(int p) const { return GlobalTab{rec + p * kTabHalfBytes}; }
};

// All threads of the block call both; `token` identifies the claim between the two.  A block of B = 128 u threads claims u
// consecutive units (aligned to u) of its SM's kTabSlotsPerSm units.
template <class C>
__device__ __forceinline__ GlobalTab<C> tab_claim(const TabScratch& ts, unsigned int& token) {
  __shared__ unsigned int s_token;
  const unsigned int units = blockDim.x / kPtBlock;
  if (threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int want = (1u << units) - 1u;
    unsigned int b = 0;
    for (;;) {  // more resident blocks than units simply wait for one to be released
      const unsigned int bits = want << b;
      const unsigned int old = atomicOr(ts.masks + smid, bits);
      if (!(old & bits)) break;
      atomicAnd(ts.masks + smid, ~(bits & ~old));  // give back the units that were free before our attempt
      b += units;
      if (b + units > (unsigned)kTabSlotsPerSm) b = 0;
    }
    s_token = smid * kTabSlotsPerSm + b;
  }
  __syncthreads();
  token = s_token;
  return GlobalTab<C>{ts.base + ((size_t)token * kPtBlock + threadIdx.x) * kTabRecordBytes};
}
__device__ __forceinline__ void tab_release(const TabScratch& ts, unsigned int token) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int units = blockDim.x / kPtBlock;
    atomicAnd(ts.masks + token / kTabSlotsPerSm, ~(((1u << units) - 1u) << (token % kTabSlotsPerSm)));
  }
}

static __global__ void nsmid_kernel(unsigned int* out) {
  unsigned int n;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n));
  *out = n;
}

// fixed-base table: kFixWindows x kFixEntries entries, one thread per entry (entry 0 of each window is unused)
constexpr int kGtabBlock = 256;
template <class C>
__global__ void __launch_bounds__(kGtabBlock) pt_gtab_kernel(typename C::Aff* gtab) {
  const unsigned e = blockIdx.x * kGtabBlock + threadIdx.x;
  const int j = (int)(e / kFixEntries);
  const uint32_t w = e % kFixEntries;
  if (j < kFixWindows && w != 0) build_gtab_entry<C>(gtab[e], j, w);
}

// strided point copy: PointShare vector <-> separate share / mac point vectors
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_copy_kernel(size_t n, PVec in, PMVec out) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    const Vec ci{in.p + i * (size_t)in.stride, 32};
    const MVec co{out.p + i * (size_t)out.stride, 32};
#pragma unroll
    for (int k = 0; k < C::kCoords; k++) {
      fe8 x;
      ld_fe(x, ci, k);
      st_fe(co, k, x);
    }
  }
}

enum class PtBin { Add, Sub };

// out[i] = a[i] (+|-) b[i]   (open-add :98-108, batch_add/batch_sub on the 2n points of n PointShares :396-465)
template <class C, PtBin OP>
__global__ void __launch_bounds__(kPtBlock) pt_binary_kernel(size_t n, PVec a, PVec b, PMVec out) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x, y;
    ld_pt<C>(x, a, i);
    ld_pt<C>(y, b, i);
    if (OP == PtBin::Sub) C::neg(y);
    C::add(x, y);
    st_pt<C>(out, i, x);
  }
}

template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_neg_kernel(size_t n, PVec a, PMVec out) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x;
    ld_pt<C>(x, a, i);
    C::neg(x);
    st_pt<C>(out, i, x);
  }
}

// out[i] = s[i >> sshift] * P[i]   (CurvePointResult::batch_mul curve.rs:459-479; batch_mul_public :718-751 with
// sshift = 1 over the 2n points of n PointShares)
template <class C>
__global__ void __launch_bounds__(kPtBlock, C::kMinBlocks) pt_mul_kernel(size_t n, Vec s, int sshift, PVec P, PMVec out, TabScratch ts) {
  unsigned int token;
  GlobalTab<C> tab = tab_claim<C>(ts, token);
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x, r;
    fe8 k;
    ld_pt<C>(x, P, i);
    ld_fe(k, s, i >> sshift);
    pt_mul_elem<C>(tab, r, k, x);
    st_pt<C>(out, i, r);
  }
  tab_release(ts, token);
}

// out[i] = (s_share[i] * P[i], s_mac[i] * P[i])   (batch_mul_authenticated curve.rs:483-517)
template <class C>
__global__ void __launch_bounds__(C::kTwoPassBlock, 1) pt_mul_auth_kernel(size_t n, Vec s_share, Vec s_mac, PVec P, PMVec out_s, PMVec out_m, TabScratch ts) {
  unsigned int token;
  GlobalTab<C> tab = tab_claim<C>(ts, token);
  const size_t step = (size_t)gridDim.x * blockDim.x;
  // block-uniform trip count (the two-pass loops synchronise the block, ARK_PHASE_SYNC): idle lanes recompute element n-1
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += step) {
    const bool live = base + threadIdx.x < n;
    const size_t i = live ? base + threadIdx.x : n - 1;
    typename C::Pt x, r0, r1;
    fe8 k0, k1;
    ld_pt<C>(x, P, i);
    ld_fe(k0, s_share, i);
    ld_fe(k1, s_mac, i);
    pt_mul2_elem<C>(tab, r0, r1, k0, k1, x);
    if (live) {
      st_pt<C>(out_s, i, r0);
      st_pt<C>(out_m, i, r1);
    }
  }
  tab_release(ts, token);
}

// out[i] = s[i] * G   (batch_mul_generator :754-780, one launch per plane)
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_mul_gen_kernel(size_t n, Vec s, const typename C::Aff* __restrict__ gtab, PMVec out) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt r;
    fe8 k;
    ld_fe(k, s, i);
    pt_mul_gen_elem<C>(r, k, gtab);
    st_pt<C>(out, i, r);
  }
}

template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_share_add_public_kernel(size_t n, int party, int sub, fe8 key, PVec a_s, PVec a_m, PVec pub,
                                                                     PMVec out_s, PMVec out_m, TabScratch ts) {
  unsigned int token;
  GlobalTab<C> tab = tab_claim<C>(ts, token);
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt s, m, P, os, om;
    ld_pt<C>(s, a_s, i);
    ld_pt<C>(m, a_m, i);
    ld_pt<C>(P, pub, i);
    pt_share_add_public_elem<C>(tab, os, om, party, sub != 0, key, s, m, P);
    st_pt<C>(out_s, i, os);
    st_pt<C>(out_m, i, om);
  }
  tab_release(ts, token);
}

template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_mac_check_kernel(size_t n, fe8 key, PVec opened, PVec mac, PMVec out, TabScratch ts) {
  unsigned int token;
  GlobalTab<C> tab = tab_claim<C>(ts, token);
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt o, m, r;
    ld_pt<C>(o, opened, i);
    ld_pt<C>(m, mac, i);
    pt_mac_check_elem<C>(tab, r, key, o, m);
    st_pt<C>(out, i, r);
  }
  tab_release(ts, token);
}

// flag (initialised to 1) is cleared if any mine[i] + peer[i] is not the identity (:128-131)
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_sum_is_identity_kernel(size_t n, PVec mine, PVec peer, int* flag) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  bool ok = true;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x, y;
    ld_pt<C>(x, mine, i);
    ld_pt<C>(y, peer, i);
    C::add(x, y);
    ok = ok && C::is_identity(x);
  }
  if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}

// flag (initialised to 1) is cleared if any point is off the curve or outside the prime-order subgroup
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_validate_kernel(size_t n, PVec a, int* flag, TabScratch ts) {
  unsigned int token;
  GlobalTab<C> tab = tab_claim<C>(ts, token);
  const size_t step = (size_t)gridDim.x * kPtBlock;
  bool ok = true;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x;
    ld_pt<C>(x, a, i);
    ok = ok && pt_valid_elem<C>(tab, x);
  }
  if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
  tab_release(ts, token);
}

// affine (x, y), 64 B per point; parity with the reference is defined on this form
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_normalize_kernel(size_t n, PVec a, MVec out_x, MVec out_y) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt p;
    fe8 x, y;
    ld_pt<C>(p, a, i);
    C::normalize(x, y, p);
    st_fe(out_x, i, x);
    st_fe(out_y, i, y);
  }
}

// inverse of pt_normalize_kernel: affine (x, y) images -> points in the reference's projective image (Z = 1)
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_from_affine_kernel(size_t n, Vec in_x, Vec in_y, PMVec out) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt p;
    fe8 x, y;
    ld_fe(x, in_x, i);
    ld_fe(y, in_y, i);
    C::from_affine(p, x, y);
    st_pt<C>(out, i, p);
  }
}

// Point Beaver phase 1
template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_beaver_mask_kernel(size_t n, Vec x_s, Vec a_s, Vec b_s, PVec P_s,
                                                                 const typename C::Aff* __restrict__ gtab, MVec d_mine, PMVec E_mine) {
  const size_t step = (size_t)gridDim.x * kPtBlock;
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    fe8 xs, as, bs, dm;
    typename C::Pt P, E;
    ld_fe(xs, x_s, i);
    ld_fe(as, a_s, i);
    ld_fe(bs, b_s, i);
    ld_pt<C>(P, P_s, i);
    pt_beaver_mask_elem<C>(dm, E, xs, as, bs, P, gtab);
    st_fe(d_mine, i, dm);
    st_pt<C>(E_mine, i, E);
  }
}

// Point Beaver phase 2 (open-add of d and E fused with the recombination)
struct PtRecombineArgs {
  Vec d_mine, d_peer;
  PVec E_mine, E_peer;
  Vec a_s, a_m, b_s, b_m, c_s, c_m;
  PMVec out_s, out_m;
  MVec d_open;
  PMVec E_open;
  fe8 key;
  int party;
  int open;
};

template <class C, int B = C::kTwoPassBlock>
__global__ void __launch_bounds__(B, 1) pt_beaver_recombine_kernel(size_t n, const __grid_constant__ PtRecombineArgs g,
                                                                      const typename C::Aff* __restrict__ gtab, TabScratch ts) {
  unsigned int token;
  // BN254: affine tables (mixed additions) pay in the 256-thread build (11.8 -> 11.3 ms at 2^17) and cost in the 384-thread one
  // (78.9 -> 81.3 ms at 2^20, profiles/r02z7_summary.txt), so the deep-grid variant keeps Jacobian entries
  constexpr bool kAff = C::kAffineTables && B == C::kTwoPassBlock;
  GlobalTab<C, kAff> tab{tab_claim<C>(ts, token).rec};
  const size_t step = (size_t)gridDim.x * blockDim.x;
  // block-uniform trip count (the two-pass loops synchronise the block, ARK_PHASE_SYNC): idle lanes recompute element n-1
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += step) {
    const bool live = base + threadIdx.x < n;
    const size_t i = live ? base + threadIdx.x : n - 1;
    fe8 dm, dp, as, am, bs, bm, cs, cm, d;
    typename C::Pt Em, Ep, E;
    ld_fe(dm, g.d_mine, i);
    ld_fe(dp, g.d_peer, i);
    ld_pt<C>(Em, g.E_mine, i);
    ld_pt<C>(Ep, g.E_peer, i);
    ld_fe(as, g.a_s, i);
    ld_fe(am, g.a_m, i);
    ld_fe(bs, g.b_s, i);
    ld_fe(bm, g.b_m, i);
    ld_fe(cs, g.c_s, i);
    ld_fe(cm, g.c_m, i);
    pt_beaver_recombine_elem<C, C::kDualChain>(tab, d, E, g.party, g.key, dm, dp, Em, Ep, as, am, bs, bm, cs, cm, gtab,
                                [&](int which, const typename C::Pt& r) { if (live) st_pt<C>(which ? g.out_m : g.out_s, i, r); });
    if (g.open && live) {
      st_fe(g.d_open, i, d);
      st_pt<C>(g.E_open, i, E);
    }
  }
  tab_release(ts, token);
}

// ---------------------------------------------------------------------------------------------
// Sum of points (the fold of AuthenticatedPointResult::msm, authenticated_curve.rs:798-803, and of CurvePoint sums): every
// thread adds a grid-stride slice, a warp-shuffle tree and a shared-memory step fold the block, one partial per block; a
// second single-block launch folds the partials.  `stride`-separated inputs let the share and mac halves of a PointShare
// vector be summed in place.
// ---------------------------------------------------------------------------------------------
template <class C>
__device__ __forceinline__ void pt_shfl_down(typename C::Pt& r, const typename C::Pt& p, int off) {
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&p);
  uint32_t* dst = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < C::kCoords * 8; j++) dst[j] = __shfl_down_sync(0xffffffffu, src[j], off);
}

template <class C>
__global__ void __launch_bounds__(kPtBlock) pt_sum_kernel(size_t n, PVec a, PMVec out, int per_block) {
  __shared__ typename C::Pt part[kPtBlock / 32];
  const size_t step = (size_t)gridDim.x * kPtBlock;
  typename C::Pt acc;
  C::set_identity(acc);
  for (size_t i = (size_t)blockIdx.x * kPtBlock + threadIdx.x; i < n; i += step) {
    typename C::Pt x;
    ld_pt<C>(x, a, i);
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
    for (int w = 1; w < kPtBlock / 32; w++) C::add(acc, part[w]);
    st_pt<C>(out, per_block ? blockIdx.x : 0, acc);
  }
}

}  // namespace ark
