// Internal: templated launch wrappers of the point kernels and the per-curve dispatch table.  Each curve is
// instantiated in four translation units (curve_<name>.cu + _beaver / _msm / _shares) so the slow-to-compile scalar
// multiplication kernels build in parallel: a unit defines ARK_CURVE_PART and gets the bodies of that part only (the other
// members stay declarations and resolve at link time against the unit that explicitly instantiates them).
// arkmpc_curve.cu holds the C ABI and dispatches through `CurveOps`.
#pragma once
#include "ctx.hpp"

namespace arkctx {

struct CurveOps {
  uint32_t point_bytes;
  int (*binary)(arkmpc_ctx*, size_t n, const void* a, const void* b, void* out, int sub);
  int (*neg)(arkmpc_ctx*, size_t n, const void* a, void* out);
  int (*share_add_public)(arkmpc_ctx*, int party, int sub, const ark::fe8& key, size_t n, const void* a_ps, const void* pub, void* out_ps);
  int (*mul)(arkmpc_ctx*, size_t n_points, const void* scalars, int sshift, const void* pts, void* out);
  int (*mul_auth)(arkmpc_ctx*, size_t n, const void* s_share, const void* s_mac, const void* pts, void* out_ps);
  int (*mul_gen)(arkmpc_ctx*, size_t n, const void* scalars, void* out, uint32_t out_stride);
  int (*beaver_mask)(arkmpc_ctx*, size_t n, const void* x_share, const void* P_ps, const void* a_share, const void* b_share, void* d_mine, void* E_mine);
  int (*beaver_recombine)(arkmpc_ctx*, int party, const ark::fe8& key, size_t n, const void* d_mine, const void* d_peer, const void* E_mine,
                          const void* E_peer, const void* a_s, const void* a_m, const void* b_s, const void* b_m, const void* c_s,
                          const void* c_m, void* out_ps, void* d_open, void* E_open);
  int (*mac_check)(arkmpc_ctx*, const ark::fe8& key, size_t n, const void* opened, const void* a_ps, void* check);
  int (*sum_is_identity)(arkmpc_ctx*, size_t n, const void* mine, const void* peer, int* flag_dev);
  int (*validate)(arkmpc_ctx*, size_t n, const void* pts, int* flag_dev);
  int (*from_affine)(arkmpc_ctx*, size_t n, const void* xy, void* out_pts);
  int (*normalize)(arkmpc_ctx*, size_t n, const void* pts, void* out_xy);
  int (*copy)(arkmpc_ctx*, size_t n, const void* in, uint32_t in_stride, void* out, uint32_t out_stride);
  int (*sum)(arkmpc_ctx*, size_t n, const void* in, uint32_t in_stride, void* out_point);
  int (*msm)(arkmpc_ctx*, size_t n, const void* scalars, const void* pts, void* out_point);
};

const CurveOps* curve_ops_bn254();
const CurveOps* curve_ops_ed25519();

}  // namespace arkctx

#ifdef ARK_CURVE_IMPL
// parts: 0 = dispatch table, linear gates, mul, mul_gen, sum; 1 = point Beaver mask / recombine; 2 = msm; 3 = share-side scalar multiplications
#ifndef ARK_CURVE_PART
#error "define ARK_CURVE_PART (0..3) before including curve_launch.cuh with ARK_CURVE_IMPL"
#endif
#define ARK_IN_PART(k) (ARK_CURVE_PART == (k))
#include "curve_kernels.cuh"
#include "curve_msm.cuh"

namespace arkctx {
using namespace ark;

inline PVec pvec(const void* p, uint32_t stride) { return PVec{static_cast<const char*>(p), stride}; }
inline PMVec pmvec(void* p, uint32_t stride) { return PMVec{static_cast<char*>(p), stride}; }
// One element per thread (hardware block scheduling): measured 6-8 % faster than persistent grids capped at 2-4 resident
// blocks per SM for the scalar-multiplication kernels (profiles/r01e_pt_grid_ab.txt).  ARKMPC_PT_BLOCKS=k (k > 0) selects a
// persistent grid of k blocks per SM instead.
inline unsigned pt_grid(const arkmpc_ctx* ctx, size_t n, int /*blocks_per_sm_hint*/, int block = kPtBlock) {
  static const int persistent_blocks = [] { const char* v = getenv("ARKMPC_PT_BLOCKS"); return v ? atoi(v) : 0; }();
  if (persistent_blocks > 0) return grid_for(ctx, n, persistent_blocks, block);
  size_t need = (n + block - 1) / block;
  return (unsigned)(need < (1u << 30) ? (need ? need : 1) : (1u << 30));
}

template <class C>
struct CurveLaunch {
  using Aff = typename C::Aff;
  static constexpr uint32_t PB = C::kPointBytes;

  // Fixed-base table of the generator, built on first use and kept for the life of the context.
  static int gtab(arkmpc_ctx* ctx, const Aff** out) {
    std::lock_guard<std::mutex> lock(ctx->gtab_mutex);
    void*& slot = ctx->gtab[C::kId];
    if (!slot) {
      void* mem = nullptr;
      ARK_CUDA(ctx, cudaMalloc(&mem, sizeof(Aff) * kFixWindows * kFixEntries));
      cudaError_t e = cudaMemsetAsync(mem, 0, sizeof(Aff) * kFixWindows * kFixEntries, ctx->stream);
      if (e == cudaSuccess) {
        pt_gtab_kernel<C><<<kFixWindows * kFixEntries / kGtabBlock, kGtabBlock, 0, ctx->stream>>>(static_cast<Aff*>(mem));
        ctx->launches++;
        e = cudaGetLastError();
      }
      // other streams may read the table later: finish it before publishing the pointer
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) {
        cudaFree(mem);
        return fail(ctx, ARKMPC_ERR_CUDA, std::string("fixed-base table: ") + cudaGetErrorString(e));
      }
      slot = mem;
    }
    *out = static_cast<const Aff*>(slot);
    return ARKMPC_OK;
  }

  // Scratch records of the variable-base window tables (curve_kernels.cuh), shared by both curves, created on first use.
  static int tab_scratch(arkmpc_ctx* ctx, TabScratch* out) {
    std::lock_guard<std::mutex> lock(ctx->gtab_mutex);
    if (!ctx->tab_scratch) {
      unsigned int* probe = nullptr;
      ARK_CUDA(ctx, cudaMalloc(&probe, sizeof(unsigned int)));
      nsmid_kernel<<<1, 1, 0, ctx->stream>>>(probe);
      unsigned int nsmid = 0;
      cudaError_t e = cudaMemcpyAsync(&nsmid, probe, sizeof nsmid, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      cudaFree(probe);
      if (e != cudaSuccess || nsmid == 0 || nsmid > 4096) return fail(ctx, ARKMPC_ERR_CUDA, "could not read the SM id range");
      const size_t bytes = (size_t)nsmid * kTabSlotsPerSm * kPtBlock * kTabRecordBytes;
      void *mem = nullptr, *masks = nullptr;
      ARK_CUDA(ctx, cudaMalloc(&mem, bytes));
      e = cudaMalloc(&masks, nsmid * sizeof(unsigned int));
      if (e == cudaSuccess) e = cudaMemsetAsync(masks, 0, nsmid * sizeof(unsigned int), ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) {
        cudaFree(mem);
        if (masks) cudaFree(masks);
        return fail(ctx, ARKMPC_ERR_CUDA, std::string("window-table scratch: ") + cudaGetErrorString(e));
      }
      ctx->tab_scratch = mem;
      ctx->tab_masks = masks;
    }
    out->base = static_cast<char*>(ctx->tab_scratch);
    out->masks = static_cast<unsigned int*>(ctx->tab_masks);
    return ARKMPC_OK;
  }

  static int binary(arkmpc_ctx* ctx, size_t n, const void* a, const void* b, void* out, int sub)
#if ARK_IN_PART(0)
  {
    if (sub) pt_binary_kernel<C, PtBin::Sub><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, pvec(a, PB), pvec(b, PB), pmvec(out, PB));
    else pt_binary_kernel<C, PtBin::Add><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, pvec(a, PB), pvec(b, PB), pmvec(out, PB));
    return post_launch(ctx, "pt_binary_kernel");
  }
#else
  ;
#endif
  static int neg(arkmpc_ctx* ctx, size_t n, const void* a, void* out)
#if ARK_IN_PART(0)
  {
    pt_neg_kernel<C><<<pt_grid(ctx, n, 8), kPtBlock, 0, ctx->stream>>>(n, pvec(a, PB), pmvec(out, PB));
    return post_launch(ctx, "pt_neg_kernel");
  }
#else
  ;
#endif
  static int share_add_public(arkmpc_ctx* ctx, int party, int sub, const fe8& key, size_t n, const void* a_ps, const void* pub, void* out_ps)
#if ARK_IN_PART(3)
  {
    const char* a = static_cast<const char*>(a_ps);
    char* o = static_cast<char*>(out_ps);
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    pt_share_add_public_kernel<C><<<pt_grid(ctx, n, 2), kPtBlock, 0, ctx->stream>>>(n, party, sub, key, pvec(a, 2 * PB), pvec(a + PB, 2 * PB), pvec(pub, PB),
                                                                                  pmvec(o, 2 * PB), pmvec(o + PB, 2 * PB), ts);
    return post_launch(ctx, "pt_share_add_public_kernel");
  }
#else
  ;
#endif
  static int mul(arkmpc_ctx* ctx, size_t n_points, const void* scalars, int sshift, const void* pts, void* out)
#if ARK_IN_PART(0)
  {
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    pt_mul_kernel<C><<<pt_grid(ctx, n_points, 2), kPtBlock, 0, ctx->stream>>>(n_points, vec(scalars), sshift, pvec(pts, PB), pmvec(out, PB), ts);
    return post_launch(ctx, "pt_mul_kernel");
  }
#else
  ;
#endif
  static int mul_auth(arkmpc_ctx* ctx, size_t n, const void* s_share, const void* s_mac, const void* pts, void* out_ps)
#if ARK_IN_PART(3)
  {
    char* o = static_cast<char*>(out_ps);
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    pt_mul_auth_kernel<C><<<pt_grid(ctx, n, 1, C::kTwoPassBlock), C::kTwoPassBlock, 0, ctx->stream>>>(n, vec(s_share), vec(s_mac), pvec(pts, PB), pmvec(o, 2 * PB), pmvec(o + PB, 2 * PB), ts);
    return post_launch(ctx, "pt_mul_auth_kernel");
  }
#else
  ;
#endif
  static int mul_gen(arkmpc_ctx* ctx, size_t n, const void* scalars, void* out, uint32_t out_stride)
#if ARK_IN_PART(0)
  {
    const Aff* g;
    int rc = gtab(ctx, &g);
    if (rc != ARKMPC_OK) return rc;
    pt_mul_gen_kernel<C><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, vec(scalars), g, pmvec(out, out_stride));
    return post_launch(ctx, "pt_mul_gen_kernel");
  }
#else
  ;
#endif
  static int beaver_mask(arkmpc_ctx* ctx, size_t n, const void* x_share, const void* P_ps, const void* a_share, const void* b_share, void* d_mine,
                         void* E_mine)
#if ARK_IN_PART(1)
  {
    const Aff* g;
    int rc = gtab(ctx, &g);
    if (rc != ARKMPC_OK) return rc;
    pt_beaver_mask_kernel<C><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, vec(x_share), vec(a_share), vec(b_share), pvec(P_ps, 2 * PB), g,
                                                                             mvec(d_mine), pmvec(E_mine, PB));
    return post_launch(ctx, "pt_beaver_mask_kernel");
  }
#else
  ;
#endif
  static int beaver_recombine(arkmpc_ctx* ctx, int party, const fe8& key, size_t n, const void* d_mine, const void* d_peer, const void* E_mine,
                              const void* E_peer, const void* a_s, const void* a_m, const void* b_s, const void* b_m, const void* c_s,
                              const void* c_m, void* out_ps, void* d_open, void* E_open)
#if ARK_IN_PART(1)
  {
    const Aff* gt;
    int rc = gtab(ctx, &gt);
    if (rc != ARKMPC_OK) return rc;
    char* o = static_cast<char*>(out_ps);
    PtRecombineArgs g;
    g.d_mine = vec(d_mine); g.d_peer = vec(d_peer);
    g.E_mine = pvec(E_mine, PB); g.E_peer = pvec(E_peer, PB);
    g.a_s = vec(a_s); g.a_m = vec(a_m); g.b_s = vec(b_s); g.b_m = vec(b_m); g.c_s = vec(c_s); g.c_m = vec(c_m);
    g.out_s = pmvec(o, 2 * PB); g.out_m = pmvec(o + PB, 2 * PB);
    g.d_open = mvec(d_open); g.E_open = pmvec(E_open, PB);
    g.key = key;
    g.party = party;
    g.open = d_open != nullptr;
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    if constexpr (C::kTwoPassBlock == 256) {
      // BN254: 12 warps per SM at 166 registers (some spills) beat 8 warps at 224 once the grid is many waves deep — 79.0 against
      // 82.6 ms at 2^20, but 12.5 against 11.8 ms at 2^17, where the coarser blocks leave a longer tail (profiles/r02t_*).
      // ARKMPC_PT_BN_BLOCK=256|384 forces one.
      static const int forced = [] { const char* v = getenv("ARKMPC_PT_BN_BLOCK"); return v ? atoi(v) : 0; }();
      if (forced == 384 || (forced != 256 && n >= ((size_t)1 << 19))) {
        pt_beaver_recombine_kernel<C, 384><<<pt_grid(ctx, n, 1, 384), 384, 0, ctx->stream>>>(n, g, gt, ts);
        return post_launch(ctx, "pt_beaver_recombine_kernel");
      }
    }
    // (Curve25519 stays at 512 threads / 128 registers: 384 / 168 measured 48.3 against 46.7 ms at 2^20, profiles/r02_point_kernels_history.txt)
    pt_beaver_recombine_kernel<C><<<pt_grid(ctx, n, 1, C::kTwoPassBlock), C::kTwoPassBlock, 0, ctx->stream>>>(n, g, gt, ts);
    return post_launch(ctx, "pt_beaver_recombine_kernel");
  }
#else
  ;
#endif
  static int mac_check(arkmpc_ctx* ctx, const fe8& key, size_t n, const void* opened, const void* a_ps, void* check)
#if ARK_IN_PART(3)
  {
    const char* a = static_cast<const char*>(a_ps);
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    pt_mac_check_kernel<C><<<pt_grid(ctx, n, 2), kPtBlock, 0, ctx->stream>>>(n, key, pvec(opened, PB), pvec(a + PB, 2 * PB), pmvec(check, PB), ts);
    return post_launch(ctx, "pt_mac_check_kernel");
  }
#else
  ;
#endif
  static int sum_is_identity(arkmpc_ctx* ctx, size_t n, const void* mine, const void* peer, int* flag_dev)
#if ARK_IN_PART(0)
  {
    pt_sum_is_identity_kernel<C><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, pvec(mine, PB), pvec(peer, PB), flag_dev);
    return post_launch(ctx, "pt_sum_is_identity_kernel");
  }
#else
  ;
#endif
  static int validate(arkmpc_ctx* ctx, size_t n, const void* pts, int* flag_dev)
#if ARK_IN_PART(0)
  {
    TabScratch ts;
    int trc = tab_scratch(ctx, &ts);
    if (trc != ARKMPC_OK) return trc;
    pt_validate_kernel<C><<<pt_grid(ctx, n, 2), kPtBlock, 0, ctx->stream>>>(n, pvec(pts, PB), flag_dev, ts);
    return post_launch(ctx, "pt_validate_kernel");
  }
#else
  ;
#endif
  static int normalize(arkmpc_ctx* ctx, size_t n, const void* pts, void* out_xy)
#if ARK_IN_PART(0)
  {
    char* o = static_cast<char*>(out_xy);
    pt_normalize_kernel<C><<<pt_grid(ctx, n, 4), kPtBlock, 0, ctx->stream>>>(n, pvec(pts, PB), mvec(o, 64), mvec(o + 32, 64));
    return post_launch(ctx, "pt_normalize_kernel");
  }
#else
  ;
#endif

  static int from_affine(arkmpc_ctx* ctx, size_t n, const void* xy, void* out_pts)
#if ARK_IN_PART(0)
  {
    const char* i = static_cast<const char*>(xy);
    pt_from_affine_kernel<C><<<pt_grid(ctx, n, 8), kPtBlock, 0, ctx->stream>>>(n, vec(i, 64), vec(i + 32, 64), pmvec(out_pts, PB));
    return post_launch(ctx, "pt_from_affine_kernel");
  }
#else
  ;
#endif

  static int copy(arkmpc_ctx* ctx, size_t n, const void* in, uint32_t in_stride, void* out, uint32_t out_stride)
#if ARK_IN_PART(0)
  {
    pt_copy_kernel<C><<<pt_grid(ctx, n, 8), kPtBlock, 0, ctx->stream>>>(n, pvec(in, in_stride), pmvec(out, out_stride));
    return post_launch(ctx, "pt_copy_kernel");
  }
#else
  ;
#endif

  // scratch: ctx->partials holds 2 * kMaxPartialBlocks field elements = 64 KiB = 512 BN254 / 512 Edwards points at most
  static int sum(arkmpc_ctx* ctx, size_t n, const void* in, uint32_t in_stride, void* out_point)
#if ARK_IN_PART(0)
  {
    const size_t cap = (size_t)2 * kMaxPartialBlocks * 32 / PB;
    size_t blocks = (n + kPtBlock - 1) / kPtBlock;
    if (blocks > cap) blocks = cap;
    if (blocks > (size_t)ctx->sm_count * 2) blocks = (size_t)ctx->sm_count * 2;
    if (blocks == 0) blocks = 1;
    pt_sum_kernel<C><<<(unsigned)blocks, kPtBlock, 0, ctx->stream>>>(n, pvec(in, in_stride), pmvec(ctx->partials, PB), 1);
    ctx->launches++;
    pt_sum_kernel<C><<<1, kPtBlock, 0, ctx->stream>>>(blocks, pvec(ctx->partials, PB), pmvec(out_point, PB), 0);
    return post_launch(ctx, "pt_sum_kernel");
  }
#else
  ;
#endif

  // Public MSM: n parallel scalar multiplications + a sum below kMsmNaiveBelow points (the reference switches to Pippenger at
  // MSM_SIZE_THRESHOLD = 10, curve.rs:34; on the GPU the bucket method's serial tail — up to 253 dependent doublings to weight
  // the top window, ~0.5-0.8 ms — only pays off from ~2^15 points: measured 2^12: 0.63 ms naive vs 1.09 ms buckets, 2^16: 1.82 vs
  // 1.48 ms, 2^20: 25.5 vs 7.1 ms on Curve25519), the bucket method of curve_msm.cuh above.  Scratch is stream-ordered.
  static constexpr size_t kMsmNaiveBelow = (size_t)1 << 15;
  static int msm(arkmpc_ctx* ctx, size_t n, const void* scalars, const void* pts, void* out_point)
#if ARK_IN_PART(2)
  {
    cudaStream_t st = ctx->stream;
    if (n < kMsmNaiveBelow) {
      void* tmp = nullptr;
      ARK_CUDA(ctx, cudaMallocAsync(&tmp, (n ? n : 1) * PB, st));
      int rc = n ? mul(ctx, n, scalars, 0, pts, tmp) : ARKMPC_OK;
      if (rc == ARKMPC_OK) rc = sum(ctx, n, tmp, PB, out_point);
      cudaFreeAsync(tmp, st);
      return rc;
    }
    int lg = 0;
    while (((size_t)2 << lg) <= n) lg++;
    int c = lg - 5;
    c = c < 4 ? 4 : (c > 15 ? 15 : c);
    const int bits = C::R::kBits;
    const int W = (bits + c - 1) / c;
    const size_t M = (size_t)W << c;
    const size_t chunks = ((size_t)1 << c) / kMsmChunk ? ((size_t)1 << c) / kMsmChunk : 1;
    uint32_t* counts = nullptr;
    uint32_t* idx = nullptr;
    char* ptmem = nullptr;
    ARK_CUDA(ctx, cudaMallocAsync(&counts, (4 * M + 1) * sizeof(uint32_t), st));
    uint32_t* offsets = counts + M;
    uint32_t* cursors = counts + 2 * M;
    uint32_t* biglist = counts + 3 * M;  // [0] = number of over-full buckets, then their keys
    cudaError_t e = cudaMallocAsync(&idx, n * (size_t)W * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMallocAsync(&ptmem, (M + chunks * W + W) * (size_t)PB, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, M * sizeof(uint32_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(biglist, 0, sizeof(uint32_t), st);
    if (e != cudaSuccess) {
      cudaFreeAsync(counts, st);
      if (idx) cudaFreeAsync(idx, st);
      if (ptmem) cudaFreeAsync(ptmem, st);
      return fail(ctx, e == cudaErrorMemoryAllocation ? ARKMPC_ERR_OOM : ARKMPC_ERR_CUDA, std::string("msm scratch: ") + cudaGetErrorString(e));
    }
    char* buckets = ptmem;
    char* partials = ptmem + M * PB;
    char* wsum = partials + chunks * W * PB;
    msm_count_kernel<C><<<grid_for(ctx, n, 8), kBlock, 0, st>>>(n, vec(scalars), c, W, counts);
    msm_scan_kernel<<<1, kMsmScanThreads, 0, st>>>(M, counts, offsets, cursors);
    msm_scatter_kernel<C><<<grid_for(ctx, n, 8), kBlock, 0, st>>>(n, vec(scalars), c, W, offsets, cursors, idx);
    msm_bucket_kernel<C><<<(unsigned)((M + kPtBlock - 1) / kPtBlock), kPtBlock, 0, st>>>(M, offsets, counts, idx, pvec(pts, PB), pmvec(buckets, PB), biglist);
    msm_bigbucket_kernel<C><<<(unsigned)(ctx->sm_count * 2), kPtBlock, 0, st>>>(offsets, counts, idx, pvec(pts, PB), pmvec(buckets, PB), biglist);
    msm_chunk_kernel<C><<<(unsigned)((chunks * W + kPtBlock - 1) / kPtBlock), kPtBlock, 0, st>>>(c, W, pvec(buckets, PB), pmvec(partials, PB));
    msm_window_kernel<C><<<W, kPtBlock, 0, st>>>(c, chunks, pvec(partials, PB), pmvec(wsum, PB));
    ctx->launches += 6;
    int rc = post_launch(ctx, "msm kernels");
    if (rc == ARKMPC_OK) rc = sum(ctx, (size_t)W, wsum, PB, out_point);
    cudaFreeAsync(counts, st);
    cudaFreeAsync(idx, st);
    cudaFreeAsync(ptmem, st);
    return rc;
  }
#else
  ;
#endif

  static const CurveOps* ops()
#if ARK_IN_PART(0)
  {
    static const CurveOps t = {PB, binary, neg, share_add_public, mul, mul_auth, mul_gen, beaver_mask, beaver_recombine, mac_check, sum_is_identity, validate, from_affine, normalize, copy, sum, msm};
    return &t;
  }
#else
  ;
#endif
};

}  // namespace arkctx
#endif  // ARK_CURVE_IMPL
