// C ABI: batch inversion and NTT over the scalar field (declared in include/arkmpc_b200.h); kernels in fr_ntt.cuh.
#include "ctx.hpp"
#include "fr_ntt.cuh"

#include <type_traits>

using namespace ark;
using namespace arkctx;

namespace {

// Launch a kernel of a dependent chain (inversion tree, NTT passes) with the programmatic-stream-serialization attribute: its
// blocks may become resident while the previous kernel of the chain drains; every such kernel starts with pdl_prologue(), i.e.
// waits for its predecessor's completion and memory flush before it touches anything (fr_ntt.cuh).
template <class... KArgs, class... Args>
void launch_chain(const arkmpc_ctx* ctx, void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, args...);  // errors surface through cudaGetLastError in post_launch
}

template <class F> struct FieldTag { using type = F; };
constexpr bool kMulKaraDefault = false;

// Twiddles w^k (k < n/2) followed by the two constants {w, n^-1}; one cached table per direction, keyed by (field, log2n), allocated
// and freed in stream order (no host synchronisation when the size changes).
template <class F>
int ntt_table(arkmpc_ctx* ctx, int field, int log2n, int inverse, const char** tw, const fe8** consts) {
  const size_t half = log2n ? (size_t)1 << (log2n - 1) : 1;
  const int dir = inverse ? 1 : 0;
  const long key = ((long)field << 16) | (long)log2n;
  if (ctx->ntt_key[dir] != key) {
    if (ctx->ntt_tw[dir]) {
      cudaFreeAsync(ctx->ntt_tw[dir], ctx->stream);  // after every transform already enqueued on this stream
      ctx->ntt_tw[dir] = nullptr;
      ctx->ntt_key[dir] = -1;
    }
    void* mem = nullptr;
    ARK_CUDA(ctx, cudaMallocAsync(&mem, (half + 2) * 32, ctx->stream));
    fe8* c = reinterpret_cast<fe8*>(static_cast<char*>(mem) + half * 32);
    fr_ntt_setup_kernel<F><<<1, 1, 0, ctx->stream>>>(log2n, inverse, c);
    ctx->launches++;
    fr_ntt_twiddle_kernel<F><<<grid_for(ctx, half, 4), kBlock, 0, ctx->stream>>>(half, c, mvec(mem));
    int rc = post_launch(ctx, "fr_ntt_twiddle_kernel");
    if (rc != ARKMPC_OK) { cudaFreeAsync(mem, ctx->stream); return rc; }
    ctx->ntt_tw[dir] = mem;
    ctx->ntt_key[dir] = key;
    ctx->ntt_stream[dir] = ctx->stream;
  } else if (ctx->ntt_stream[dir] != ctx->stream) {
    // the context was re-pointed at another stream since the table was built: order this stream after the builder
    cudaEvent_t ev;
    ARK_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaEventRecord(ev, ctx->ntt_stream[dir]);
    cudaStreamWaitEvent(ctx->stream, ev, 0);
    cudaEventDestroy(ev);
    ctx->ntt_stream[dir] = ctx->stream;
  }
  *tw = static_cast<const char*>(ctx->ntt_tw[dir]);
  *consts = reinterpret_cast<const fe8*>(*tw + half * 32);
  return ARKMPC_OK;
}

// ARKMPC_MUL=cios|kara: the field product of the NTT butterflies and the inversion sweeps
inline bool use_kara() {
  static const bool k = [] { const char* e = getenv("ARKMPC_MUL"); return e ? strcmp(e, "kara") == 0 : kMulKaraDefault; }();
  return k;
}

// ARKMPC_INV_COOP=0: the small levels of the inversion tree with one thread per group, like the large ones
inline bool use_coop() {
  static const bool c = [] { const char* e = getenv("ARKMPC_INV_COOP"); return !(e && strcmp(e, "0") == 0); }();
  return c;
}

template <class F, bool K>
int ntt_plane(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in, uint64_t* out) {
  const char* tw;
  const fe8* consts;
  int rc = ntt_table<F>(ctx, field, log2n, inverse, &tw, &consts);
  if (rc != ARKMPC_OK) return rc;
  const size_t n = (size_t)1 << log2n;
  const int tile_log = log2n < kNttTileLog ? log2n : kNttTileLog;
  const size_t tiles = n >> tile_log;
  static const int tile_threads = [] { const char* e = getenv("ARKMPC_NTT_TILE_THREADS"); const int v = e ? atoi(e) : 0; return v == 128 || v == 256 ? v : kNttThreads; }();
  static const int stride_threads = [] { const char* e = getenv("ARKMPC_NTT_STRIDE_THREADS"); const int v = e ? atoi(e) : 0; return v == 128 || v == 256 || v == 512 ? v : 0; }();
  const bool fold_scale = inverse && log2n >= 2;  // n^-1 rides on the first double stage of the tile kernel
  launch_chain(ctx, fr_ntt_tile_kernel<F, K>, (unsigned)tiles, (unsigned)tile_threads, ((size_t)32 << tile_log), ctx->stream, log2n, vec(in), vec(tw), fold_scale ? consts + 1 : (const fe8*)nullptr, mvec(out));
  rc = post_launch(ctx, "fr_ntt_tile_kernel");
  const int rest = log2n > kNttTileLog ? log2n - kNttTileLog : 0;
  int passes = (rest + kNttStrideLog - 1) / kNttStrideLog;
  for (int s0 = kNttTileLog; rc == ARKMPC_OK && s0 < log2n; passes--) {  // the remaining stages, spread evenly over <= 5-stage passes
    const int T = (log2n - s0 + passes - 1) / passes;
    const size_t blocks = n >> (5 + T);
    // large transforms: two butterflies per thread per level — blocks half the size of the tile's butterfly count stagger their
    // load, barrier and store phases (2^20: 266 -> 254 us, 2^22: 1219 -> 1039 us, 2^24: 5008 -> 4282 us for the inverse transform,
    // profiles/r02u_summary.txt); small ones keep one butterfly per thread, where the latency of a level is what counts
    const int one_each = 32 << (T - 1);  // one butterfly per thread per level
    const int threads = stride_threads ? stride_threads : (log2n >= 19 && one_each >= 64 ? one_each / 2 : one_each);
    launch_chain(ctx, fr_ntt_strided_kernel<F, K>, (unsigned)blocks, (unsigned)threads, 0, ctx->stream, log2n, s0, T, vec(tw), mvec(out));
    rc = post_launch(ctx, "fr_ntt_strided_kernel");
    s0 += T;
  }
  if (rc == ARKMPC_OK && inverse && !fold_scale) {
    fr_scale_dev_kernel<F><<<grid_for(ctx, n, 8), kBlock, 0, ctx->stream>>>(n, vec(out), consts + 1, mvec(out));
    rc = post_launch(ctx, "fr_scale_dev_kernel");
  }
  return rc;
}

int fft_impl(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in0, const uint64_t* in1, uint64_t* out0, uint64_t* out1) {
  ARK_CHECK_CTX(ctx);
  ARK_REQUIRE(ctx, field == ARKMPC_BN254_FR || field == ARKMPC_CURVE25519_FR, "unknown field id");
  // ark-ff: Curve25519 Fr has two-adicity 2 and no reference test instantiates it; only BN254 Fr (two-adicity 28) is supported
  if (field != ARKMPC_BN254_FR) return fail(ctx, ARKMPC_ERR_UNSUPPORTED, "FFT domains exist for BN254 Fr only");
  ARK_REQUIRE(ctx, log2n >= 0 && log2n <= NttRoot::kTwoAdicity, "domain size must be 2^0 .. 2^28");
  ARK_REQUIRE(ctx, in0 && out0 && aligned32(in0) && aligned32(out0) && in0 != out0, "null, misaligned or aliased plane (the transform is out of place)");
  if (in1 || out1) ARK_REQUIRE(ctx, in1 && out1 && aligned32(in1) && aligned32(out1) && in1 != out1, "null, misaligned or aliased plane");
  auto plane = use_kara() ? ntt_plane<Bn254Fr, true> : ntt_plane<Bn254Fr, false>;
  int rc = plane(ctx, field, log2n, inverse, in0, out0);
  if (rc == ARKMPC_OK && in1) rc = plane(ctx, field, log2n, inverse, in1, out1);
  return rc;
}

}  // namespace

extern "C" {

int arkmpc_fr_batch_inverse(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out) {
  ARK_CHECK_CTX(ctx);
  if (n == 0) return ARKMPC_OK;
  ARK_REQUIRE(ctx, a && out && aligned32(a) && aligned32(out), "null or misaligned plane");
  ARK_REQUIRE(ctx, field == ARKMPC_BN254_FR || field == ARKMPC_CURVE25519_FR, "unknown field id");
  // level sizes n_0 = n, n_(l+1) = ceil(n_l / kInvGroup) until <= kInvTop
  size_t sizes[16];
  int levels = 0;
  sizes[0] = n;
  while (sizes[levels] > kInvTop && levels < 14) {
    sizes[levels + 1] = (sizes[levels] + kInvGroup - 1) / kInvGroup;
    levels++;
  }
  // scratch per level above the input: the six inner products of every group (6 n_(l+1)), the group products x_(l+1) and their inverses
  size_t elems = 0;
  for (int l = 0; l < levels; l++) elems += 8 * sizes[l + 1];
  cudaStream_t s = ctx->stream;
  char* scratch = nullptr;
  if (elems) ARK_CUDA(ctx, cudaMallocAsync(&scratch, elems * 32, s));
  const char* xs[16];
  char *inv[16], *tree[16];
  xs[0] = reinterpret_cast<const char*>(a);
  inv[0] = reinterpret_cast<char*>(out);
  char* p = scratch;
  for (int l = 0; l < levels; l++) {
    tree[l] = p; p += 6 * sizes[l + 1] * 32;
    xs[l + 1] = p; p += sizes[l + 1] * 32;
    inv[l + 1] = p; p += sizes[l + 1] * 32;
  }
  static const int inv_block = [] { const char* e = getenv("ARKMPC_INV_BLOCK"); const int v = e ? atoi(e) : 0; return v == 32 || v == 64 || v == 128 ? v : kInvLaunchBlock; }();
  auto blocks = [](size_t work) { return (unsigned)((work + inv_block - 1) / inv_block); };
  int rc = ARKMPC_OK;
  auto sweeps = [&](auto fld, auto kara) -> int {
    using F = typename decltype(fld)::type;
    constexpr bool K = decltype(kara)::value;
    int rc = ARKMPC_OK;
    auto coop = [](size_t groups) { return use_coop() && groups <= kInvCoopMaxGroups; };
    auto coop_blocks = [](size_t groups) { return (unsigned)((groups + kInvCoopGroups - 1) / kInvCoopGroups); };
    for (int l = 0; l < levels && rc == ARKMPC_OK; l++) {
      if (coop(sizes[l + 1]))
        launch_chain(ctx, fr_inv_up_coop_kernel<F, K>, coop_blocks(sizes[l + 1]), (unsigned)kInvCoopBlock, 0, s, sizes[l], sizes[l + 1], vec(xs[l]), mvec(tree[l]), mvec(const_cast<char*>(xs[l + 1])));
      else
        launch_chain(ctx, fr_inv_up_kernel<F, K>, blocks(sizes[l + 1]), (unsigned)inv_block, 0, s, sizes[l], sizes[l + 1], vec(xs[l]), mvec(tree[l]), mvec(const_cast<char*>(xs[l + 1])));
      rc = post_launch(ctx, "fr_inv_up_kernel");
    }
    if (rc == ARKMPC_OK) {
      // a plain launch: placed while its predecessor still holds the SMs, the single-warp blocks land unevenly and the inversions,
      // which want one warp per scheduler, take longer (2^20: 105 against 95 us for the whole inversion, profiles/r02z20_summary.txt)
      fr_inv_top_kernel<F><<<(unsigned)((sizes[levels] + kInvTopBlock - 1) / kInvTopBlock), kInvTopBlock, 0, s>>>(sizes[levels], vec(xs[levels]), mvec(inv[levels]));
      rc = post_launch(ctx, "fr_inv_top_kernel");
    }
    for (int l = levels - 1; l >= 0 && rc == ARKMPC_OK; l--) {
      if (coop(sizes[l + 1]))
        launch_chain(ctx, fr_inv_down_coop_kernel<F, K>, coop_blocks(sizes[l + 1]), (unsigned)kInvCoopBlock, 0, s, sizes[l], sizes[l + 1], vec(xs[l]), vec(tree[l]), vec(inv[l + 1]), mvec(inv[l]));
      else
        launch_chain(ctx, fr_inv_down_kernel<F, K>, blocks(sizes[l + 1]), (unsigned)inv_block, 0, s, sizes[l], sizes[l + 1], vec(xs[l]), vec(tree[l]), vec(inv[l + 1]), mvec(inv[l]));
      rc = post_launch(ctx, "fr_inv_down_kernel");
    }
    return rc;
  };
  ARK_FIELD_SWITCH(ctx, field, {
    rc = use_kara() ? sweeps(FieldTag<F>{}, std::true_type{}) : sweeps(FieldTag<F>{}, std::false_type{});
  });
  if (scratch) cudaFreeAsync(scratch, s);
  return rc;
}

int arkmpc_fr_fft(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in, uint64_t* out) {
  return fft_impl(ctx, field, log2n, inverse, in, nullptr, out, nullptr);
}

int arkmpc_fr_share_fft(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in_share, const uint64_t* in_mac, uint64_t* out_share,
                        uint64_t* out_mac) {
  if (!in_mac || !out_mac) return fail(ctx, ARKMPC_ERR_INVALID, "null pointer");
  return fft_impl(ctx, field, log2n, inverse, in_share, in_mac, out_share, out_mac);
}

}  // extern "C"
