// C ABI, point gates (declared in include/arkmpc_b200.h): argument checks and dispatch to the per-curve launchers.
#include "curve_launch.cuh"

using namespace ark;
using namespace arkctx;

namespace {

const CurveOps* ops_for(int curve) {
  switch (curve) {
    case ARKMPC_BN254_G1: return curve_ops_bn254();
    case ARKMPC_CURVE25519_EDWARDS: return curve_ops_ed25519();
    default: return nullptr;
  }
}

int check_ptrs(arkmpc_ctx* ctx, std::initializer_list<const void*> ptrs) {
  for (const void* p : ptrs) {
    if (!p) return fail(ctx, ARKMPC_ERR_INVALID, "null pointer");
    if (!aligned32(p)) return fail(ctx, ARKMPC_ERR_INVALID, "arrays must be 32-byte aligned");
  }
  return ARKMPC_OK;
}

// common prologue: context, curve id, empty batch, pointer checks
#define ARK_PT_PROLOGUE(...)                                        \
  ARK_CHECK_CTX(ctx);                                               \
  const CurveOps* ops = ops_for(curve);                             \
  ARK_REQUIRE(ctx, ops != nullptr, "unknown curve id");             \
  if (n == 0) return ARKMPC_OK;                                     \
  {                                                                 \
    int _rc = check_ptrs(ctx, {__VA_ARGS__});                       \
    if (_rc != ARKMPC_OK) return _rc;                               \
  }

}  // namespace

extern "C" {

size_t arkmpc_point_bytes(int curve) {
  const CurveOps* ops = ops_for(curve);
  return ops ? ops->point_bytes : 0;
}

int arkmpc_pt_add(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  ARK_PT_PROLOGUE(a, b, out);
  return ops->binary(ctx, n, a, b, out, 0);
}
int arkmpc_pt_sub(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out) {
  ARK_PT_PROLOGUE(a, b, out);
  return ops->binary(ctx, n, a, b, out, 1);
}
int arkmpc_pt_neg(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a, uint64_t* out) {
  ARK_PT_PROLOGUE(a, out);
  return ops->neg(ctx, n, a, out);
}

int arkmpc_pt_share_add_public(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n, const uint64_t* a_ps,
                               const uint64_t* pub_pts, uint64_t* out_ps) {
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, key_host, "null key");
  ARK_PT_PROLOGUE(a_ps, pub_pts, out_ps);
  return ops->share_add_public(ctx, party_id, 0, load_host_fe(key_host), n, a_ps, pub_pts, out_ps);
}
int arkmpc_pt_share_sub_public(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n, const uint64_t* a_ps,
                               const uint64_t* pub_pts, uint64_t* out_ps) {
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, key_host, "null key");
  ARK_PT_PROLOGUE(a_ps, pub_pts, out_ps);
  return ops->share_add_public(ctx, party_id, 1, load_host_fe(key_host), n, a_ps, pub_pts, out_ps);
}

int arkmpc_pt_mul(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out_pts) {
  ARK_PT_PROLOGUE(scalars, pts, out_pts);
  return ops->mul(ctx, n, scalars, 0, pts, out_pts);
}
int arkmpc_pt_share_mul_public(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* a_ps, uint64_t* out_ps) {
  ARK_PT_PROLOGUE(scalars, a_ps, out_ps);
  return ops->mul(ctx, 2 * n, scalars, 1, a_ps, out_ps);  // the 2n points of n PointShares; point i uses scalar i >> 1
}
int arkmpc_pt_mul_authenticated(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac, const uint64_t* pts,
                                uint64_t* out_ps) {
  ARK_PT_PROLOGUE(s_share, s_mac, pts, out_ps);
  return ops->mul_auth(ctx, n, s_share, s_mac, pts, out_ps);
}
int arkmpc_pt_mul_generator_public(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, uint64_t* out_pts) {
  ARK_PT_PROLOGUE(scalars, out_pts);
  return ops->mul_gen(ctx, n, scalars, out_pts, ops->point_bytes);
}
int arkmpc_pt_mul_generator(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac, uint64_t* out_ps) {
  ARK_PT_PROLOGUE(s_share, s_mac, out_ps);
  int rc = ops->mul_gen(ctx, n, s_share, out_ps, 2 * ops->point_bytes);
  if (rc != ARKMPC_OK) return rc;
  return ops->mul_gen(ctx, n, s_mac, reinterpret_cast<char*>(out_ps) + ops->point_bytes, 2 * ops->point_bytes);
}

int arkmpc_pt_beaver_mask(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* x_share, const uint64_t* P_ps, const uint64_t* a_share,
                          const uint64_t* b_share, uint64_t* d_mine, uint64_t* E_mine_pts) {
  ARK_PT_PROLOGUE(x_share, P_ps, a_share, b_share, d_mine, E_mine_pts);
  return ops->beaver_mask(ctx, n, x_share, P_ps, a_share, b_share, d_mine, E_mine_pts);
}

int arkmpc_pt_beaver_recombine(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n, const uint64_t* d_mine,
                               const uint64_t* d_peer, const uint64_t* E_mine_pts, const uint64_t* E_peer_pts, const uint64_t* a_share,
                               const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac, const uint64_t* c_share,
                               const uint64_t* c_mac, uint64_t* out_ps, uint64_t* d_open, uint64_t* E_open_pts) {
  ARK_REQUIRE(ctx, party_id == 0 || party_id == 1, "party_id must be 0 or 1");
  ARK_REQUIRE(ctx, (d_open == nullptr) == (E_open_pts == nullptr), "d_open and E_open_pts must both be given or both be NULL");
  ARK_REQUIRE(ctx, key_host, "null key");
  ARK_PT_PROLOGUE(d_mine, d_peer, E_mine_pts, E_peer_pts, a_share, a_mac, b_share, b_mac, c_share, c_mac, out_ps);
  ARK_REQUIRE(ctx, aligned32(d_open) && aligned32(E_open_pts), "arrays must be 32-byte aligned");
  return ops->beaver_recombine(ctx, party_id, load_host_fe(key_host), n, d_mine, d_peer, E_mine_pts, E_peer_pts, a_share, a_mac, b_share, b_mac,
                               c_share, c_mac, out_ps, d_open, E_open_pts);
}

int arkmpc_pt_mac_check(arkmpc_ctx* ctx, int curve, const uint64_t* key_host, size_t n, const uint64_t* opened_pts, const uint64_t* a_ps,
                        uint64_t* check_pts) {
  ARK_REQUIRE(ctx, key_host, "null key");
  ARK_PT_PROLOGUE(opened_pts, a_ps, check_pts);
  return ops->mac_check(ctx, load_host_fe(key_host), n, opened_pts, a_ps, check_pts);
}

int arkmpc_pt_sum_is_identity(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* mine_pts, const uint64_t* peer_pts, int* all_identity_host) {
  if (!ctx || !all_identity_host) return ARKMPC_ERR_INVALID;
  *all_identity_host = 1;
  ARK_PT_PROLOGUE(mine_pts, peer_pts);
  *ctx->flag_host = 1;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_dev, ctx->flag_host, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  int rc = ops->sum_is_identity(ctx, n, mine_pts, peer_pts, ctx->flag_dev);
  if (rc != ARKMPC_OK) return rc;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ARK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *all_identity_host = *ctx->flag_host;
  return ARKMPC_OK;
}

int arkmpc_pt_validate(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, int* all_valid_host) {
  if (!ctx || !all_valid_host) return ARKMPC_ERR_INVALID;
  *all_valid_host = 1;
  ARK_PT_PROLOGUE(pts);
  *ctx->flag_host = 1;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_dev, ctx->flag_host, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  int rc = ops->validate(ctx, n, pts, ctx->flag_dev);
  if (rc != ARKMPC_OK) return rc;
  ARK_CUDA(ctx, cudaMemcpyAsync(ctx->flag_host, ctx->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ARK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *all_valid_host = *ctx->flag_host;
  return ARKMPC_OK;
}

int arkmpc_pt_share_split(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_ps, uint64_t* share_pts, uint64_t* mac_pts) {
  ARK_PT_PROLOGUE(a_ps);
  ARK_REQUIRE(ctx, aligned32(share_pts) && aligned32(mac_pts), "arrays must be 32-byte aligned");
  const uint32_t pb = ops->point_bytes;
  if (share_pts) {
    int rc = ops->copy(ctx, n, a_ps, 2 * pb, share_pts, pb);
    if (rc != ARKMPC_OK) return rc;
  }
  if (mac_pts) return ops->copy(ctx, n, reinterpret_cast<const char*>(a_ps) + pb, 2 * pb, mac_pts, pb);
  return ARKMPC_OK;
}

int arkmpc_pt_share_join(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* share_pts, const uint64_t* mac_pts, uint64_t* out_ps) {
  ARK_PT_PROLOGUE(share_pts, mac_pts, out_ps);
  const uint32_t pb = ops->point_bytes;
  int rc = ops->copy(ctx, n, share_pts, pb, out_ps, 2 * pb);
  if (rc != ARKMPC_OK) return rc;
  return ops->copy(ctx, n, mac_pts, pb, reinterpret_cast<char*>(out_ps) + pb, 2 * pb);
}

int arkmpc_pt_sum(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, uint64_t* out_pt) {
  ARK_CHECK_CTX(ctx);
  const CurveOps* ops = ops_for(curve);
  ARK_REQUIRE(ctx, ops != nullptr, "unknown curve id");
  ARK_REQUIRE(ctx, out_pt && aligned32(out_pt) && (n == 0 || (pts && aligned32(pts))), "null or misaligned array");
  return ops->sum(ctx, n, n ? pts : out_pt, ops->point_bytes, out_pt);  // n == 0 gives the identity, like an empty Rust sum()
}

int arkmpc_pt_share_sum(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_ps, uint64_t* out_ps) {
  ARK_CHECK_CTX(ctx);
  const CurveOps* ops = ops_for(curve);
  ARK_REQUIRE(ctx, ops != nullptr, "unknown curve id");
  ARK_REQUIRE(ctx, out_ps && aligned32(out_ps) && (n == 0 || (a_ps && aligned32(a_ps))), "null or misaligned array");
  const uint32_t pb = ops->point_bytes;
  const char* a = reinterpret_cast<const char*>(n ? a_ps : out_ps);
  int rc = ops->sum(ctx, n, a, 2 * pb, out_ps);
  if (rc != ARKMPC_OK) return rc;
  return ops->sum(ctx, n, a + pb, 2 * pb, reinterpret_cast<char*>(out_ps) + pb);
}

int arkmpc_pt_msm(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out_pt) {
  ARK_CHECK_CTX(ctx);
  const CurveOps* ops = ops_for(curve);
  ARK_REQUIRE(ctx, ops != nullptr, "unknown curve id");
  ARK_REQUIRE(ctx, out_pt && aligned32(out_pt) && (n == 0 || (scalars && pts && aligned32(scalars) && aligned32(pts))), "null or misaligned array");
  return ops->msm(ctx, n, scalars, pts, out_pt);
}

// CurvePoint::msm_authenticated (curve.rs:619-642): (sum share_i * P_i, sum mac_i * P_i)
int arkmpc_pt_msm_authenticated(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac, const uint64_t* pts,
                                uint64_t* out_ps) {
  ARK_CHECK_CTX(ctx);
  const CurveOps* ops = ops_for(curve);
  ARK_REQUIRE(ctx, ops != nullptr, "unknown curve id");
  ARK_REQUIRE(ctx, out_ps && aligned32(out_ps) && (n == 0 || (s_share && s_mac && pts && aligned32(s_share) && aligned32(s_mac) && aligned32(pts))),
              "null or misaligned array");
  int rc = ops->msm(ctx, n, s_share, pts, out_ps);
  if (rc != ARKMPC_OK) return rc;
  return ops->msm(ctx, n, s_mac, pts, reinterpret_cast<char*>(out_ps) + ops->point_bytes);
}

int arkmpc_pt_normalize(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, uint64_t* out_xy) {
  ARK_PT_PROLOGUE(pts, out_xy);
  return ops->normalize(ctx, n, pts, out_xy);
}

/* inverse of arkmpc_pt_normalize */
int arkmpc_pt_from_affine(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* xy, uint64_t* out_pts) {
  ARK_PT_PROLOGUE(xy, out_pts);
  return ops->from_affine(ctx, n, xy, out_pts);
}

}  // extern "C"
