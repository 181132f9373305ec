/* arkmpc_b200 — C ABI of the Blackwell-native online-phase gate engine for ark-mpc.
 *
 * This is the drop-in boundary: the batched gate evaluation that the reference runs inside Rust
 * closures handed to `MpcFabric::new_batch_gate_op` (/root/reference/online-phase/src/fabric.rs:841-854)
 * is exposed here as plain C entry points over device memory.  The reference has no FFI on this path;
 * each function below names the reference closure / operator it replaces.  INTEGRATION.md shows the
 * Rust `extern "C"` block and the shim a maintainer would add.
 *
 * Conventions
 *  - Every function returns 0 (ARKMPC_OK) or a negative arkmpc_status; nothing throws across the ABI.
 *  - Field elements are the reference's memory image: canonical Montgomery residues (R = 2^256) in
 *    4 little-endian u64 limbs = 32 bytes (`Scalar<C>`, algebra/scalar/scalar.rs:46).
 *  - Device layout is PLANAR: a vector of n scalars is one contiguous plane of n*32 bytes; a vector
 *    of n `ScalarShare`s (algebra/scalar/share.rs:32-37) is two planes, `share` and `mac`.  The
 *    reference's AoS image {share,mac} (64 B) converts with arkmpc_share_unzip / arkmpc_share_zip.
 *    Planes must be 32-byte aligned.  Output planes may alias input planes element-for-element.
 *  - Pointers named `*_dev`/planes are DEVICE pointers; `key` arguments are HOST pointers to 4 u64
 *    (Montgomery image of the party's MAC-key share, `MpcFabric::mac_key()`); `*_host` are host buffers.
 *  - All work is enqueued on the context's stream and is asynchronous unless stated.  A context is bound to
 *    one device and is THREAD-SAFE: any number of threads may call into the same context concurrently
 *    (the reference runs gate closures on its executor thread or on arbitrary rayon workers,
 *    fabric/executor/multi_threaded/executor.rs:208-217, and clones result handles across tokio tasks,
 *    fabric/result.rs:262-266).  Each call holds the context's lock while it enqueues; everything goes to
 *    the one context stream, so work runs on the GPU in the order the calls RETURN — a gate scheduled after
 *    its inputs' calls have returned (on whatever threads) sees their results, with no event to manage.
 *    The calling thread's current CUDA device is restored before a call returns.  arkmpc_last_error is
 *    per thread.  Two parties in one process use two contexts, mirroring execute_mock_mpc
 *    (online-phase/src/lib.rs:157-201).
 *  - Values received from the PEER (d_peer / e_peer, peer MAC-check shares, E_peer points) must already be
 *    valid: canonical field elements (< p) and on-curve, prime-order-subgroup points.  The reference gets
 *    this from arkworks' validating deserialisation (scalar.rs:187-202, curve.rs:105-135); a shim passes
 *    only deserialised values, or checks a buffer with arkmpc_fr_validate / arkmpc_pt_validate first.
 *  - n == 0 is a no-op returning ARKMPC_OK (the reference returns empty vectors,
 *    authenticated_scalar.rs:854-856).  Length mismatches cannot occur: one n per call.
 */
#ifndef ARKMPC_B200_H
#define ARKMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARKMPC_ABI_VERSION 2

typedef enum arkmpc_status {
  ARKMPC_OK = 0,
  ARKMPC_ERR_INVALID = -1,     /* bad argument (null pointer, unknown field, misaligned plane, bad party id) */
  ARKMPC_ERR_CUDA = -2,        /* a CUDA runtime call failed; see arkmpc_last_error */
  ARKMPC_ERR_NO_DEVICE = -3,   /* no usable sm_100 device */
  ARKMPC_ERR_OOM = -4,
  ARKMPC_ERR_UNSUPPORTED = -5, /* e.g. NCCL entry point on a build without NCCL */
  ARKMPC_ERR_NCCL = -6
} arkmpc_status;

/* Scalar fields (C::ScalarField of the supported curves). */
typedef enum arkmpc_field { ARKMPC_BN254_FR = 0, ARKMPC_CURVE25519_FR = 1 } arkmpc_field;
/* Curve groups. */
typedef enum arkmpc_curve { ARKMPC_BN254_G1 = 0, ARKMPC_CURVE25519_EDWARDS = 1 } arkmpc_curve;

typedef struct arkmpc_ctx arkmpc_ctx;

/* ---- library / context ---- */
int arkmpc_abi_version(void);
const char* arkmpc_status_string(int status);
int arkmpc_device_count(int* count);
int arkmpc_ctx_create(int device, arkmpc_ctx** out);
int arkmpc_ctx_destroy(arkmpc_ctx* ctx);
/* Run on a caller-owned cudaStream_t.  The handle is used verbatim: NULL is the CUDA default stream
 * (what torch's default stream reports), not "no stream".  A fresh context runs on its own
 * non-blocking stream; arkmpc_ctx_reset_stream returns to it. */
int arkmpc_ctx_set_stream(arkmpc_ctx* ctx, void* cuda_stream);
int arkmpc_ctx_reset_stream(arkmpc_ctx* ctx);
/* Scheduling hint for the executor thread (the fabric knows its gate DAG).  Beaver kernels (arkmpc_fr_beaver_mask /
 * _recombine / _recombine_sum / _recombine_gather*) launched WITH the hint form a group with the un-hinted launch that
 * precedes them; the members of a group must be mutually independent — none reads or overwrites what another writes or
 * reads (e.g. the same gate of the two parties in a mock run, or batch_muls on unrelated operands).  A hinted launch starts
 * its loads and arithmetic while the earlier members of its group are still draining instead of waiting for them (its
 * programmatic-dependency wait moves to its last instruction); dependencies on everything launched before the group, and of
 * everything launched after it, are honoured as usual.  The hint is consumed by one launch.  It is meant for ONE scheduling
 * thread per context: with concurrent callers the predecessor is unknown, so do not use it there. */
int arkmpc_ctx_hint_independent(arkmpc_ctx* ctx);
void* arkmpc_ctx_get_stream(arkmpc_ctx* ctx);
int arkmpc_ctx_device(arkmpc_ctx* ctx);
int arkmpc_ctx_sync(arkmpc_ctx* ctx);
int arkmpc_ctx_sm_count(arkmpc_ctx* ctx);
/* Human-readable detail of the last failure on this context (valid until the next call). */
const char* arkmpc_last_error(arkmpc_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t arkmpc_ctx_launch_count(arkmpc_ctx* ctx);

/* ---- memory ----
 * The reference allocates a fresh Vec per gate result (fabric/result.rs:47-64); a host that mirrors it allocates and frees
 * tens of device buffers per batch.  arkmpc_free therefore does not return memory to the driver: the block goes to a
 * per-device cache (by size class) and the next arkmpc_malloc of that class gets it back without any host synchronisation.
 * Ordering: the stream of the context that receives a cached block waits for everything submitted, up to that moment, on
 * the current stream of every live context of the device (device buffers travel between the parties' contexts by
 * reference, so any of them may still be working on the block); work the caller submitted on any other stream must be
 * synchronised before the free.  arkmpc_mem_trim returns the
 * cached blocks to the driver (synchronising); ARKMPC_ALLOC_CACHE_MB caps the cache (default 4096, 0 = cudaMalloc / cudaFree). */
int arkmpc_malloc(arkmpc_ctx* ctx, size_t bytes, void** dev_ptr);
int arkmpc_free(arkmpc_ctx* ctx, void* dev_ptr);
int arkmpc_mem_trim(arkmpc_ctx* ctx);
int arkmpc_mem_cached_bytes(arkmpc_ctx* ctx, size_t* bytes); /* bytes held in free blocks on the context's device */
int arkmpc_host_alloc(arkmpc_ctx* ctx, size_t bytes, void** pinned_ptr); /* pinned host memory */
int arkmpc_host_free(arkmpc_ctx* ctx, void* pinned_ptr);
/* Asynchronous on the context stream.  A source of up to 256 KB, and any pageable source, has been read when the call
 * returns (small copies are staged through a pinned ring); a larger pinned source must stay valid until the stream is
 * synchronised. */
int arkmpc_memcpy_h2d(arkmpc_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int arkmpc_memcpy_d2h(arkmpc_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* async on stream */
int arkmpc_memcpy_d2d(arkmpc_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes);

/* ---- layout: reference AoS ScalarShare image <-> planes (share.rs:32-37) ---- */
int arkmpc_share_unzip(arkmpc_ctx* ctx, size_t n, const uint64_t* aos_dev, uint64_t* share_plane, uint64_t* mac_plane);
int arkmpc_share_zip(arkmpc_ctx* ctx, size_t n, const uint64_t* share_plane, const uint64_t* mac_plane, uint64_t* aos_dev);

/* ---- the hot path: authenticated Beaver multiplication (authenticated_scalar.rs:848-879) ---- */

/* Phase 1, replaces the two `batch_sub` gates (:863-864, closure :679-684) as far as `open_batch`
 * consumes them (:141-145 sends only the share component):
 *   d_mine[i] = x_share[i] - a_share[i],  e_mine[i] = y_share[i] - b_share[i]. */
int arkmpc_fr_beaver_mask(arkmpc_ctx* ctx, int field, size_t n,
                          const uint64_t* x_share, const uint64_t* y_share,
                          const uint64_t* a_share, const uint64_t* b_share,
                          uint64_t* d_mine, uint64_t* e_mine);

/* Phase 2, replaces the open-add gate (:161-171), `ScalarResult::batch_mul` (scalar_result.rs:257-278),
 * 2x `batch_mul_public` (:883-916), `batch_add_public` (:493-528, share.rs:74-77) and 2x `batch_add`
 * (:457-489) with one fused kernel:
 *   d = d_mine + d_peer, e = e_mine + e_peer,
 *   out = d*e (party 0 share only; mac gets key*d*e) + d*[b] + e*[a] + [c].
 * d_open / e_open are optional (NULL to skip) outputs of the opened masks. */
int arkmpc_fr_beaver_recombine(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* d_mine, const uint64_t* e_mine,
                               const uint64_t* d_peer, const uint64_t* e_peer,
                               const uint64_t* a_share, const uint64_t* a_mac,
                               const uint64_t* b_share, const uint64_t* b_mac,
                               const uint64_t* c_share, const uint64_t* c_mac,
                               uint64_t* out_share, uint64_t* out_mac,
                               uint64_t* d_open, uint64_t* e_open);

/* Phase 2 fused with the `Sum` that follows it in an inner product (`a.iter().zip(b).map(|(a, b)| a * b).sum()`,
 * integration/src/circuits.rs:22-50; Sum for AuthenticatedScalarResult, authenticated_scalar.rs:563-576): same inputs as
 * arkmpc_fr_beaver_recombine, but the n products are never written; out_share / out_mac receive the ONE ScalarShare
 * sum_i [x_i * y_i] (32 bytes each).  Bit-identical to arkmpc_fr_beaver_recombine followed by arkmpc_fr_share_sum;
 * n == 0 gives the additive identity like an empty sum(). */
int arkmpc_fr_beaver_recombine_sum(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                   const uint64_t* d_mine, const uint64_t* e_mine,
                                   const uint64_t* d_peer, const uint64_t* e_peer,
                                   const uint64_t* a_share, const uint64_t* a_mac,
                                   const uint64_t* b_share, const uint64_t* b_mac,
                                   const uint64_t* c_share, const uint64_t* c_mac,
                                   uint64_t* out_share, uint64_t* out_mac);

/* ---- public-scalar vector gates (algebra/scalar/scalar_result.rs:170-278) ---- */
int arkmpc_fr_add(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out); /* also the open-add */
int arkmpc_fr_sub(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out);
int arkmpc_fr_mul(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out);
int arkmpc_fr_neg(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out);
/* out[i] = a[i] * s (s: one host scalar) */
int arkmpc_fr_scale(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, const uint64_t* s_host, uint64_t* out);

/* ---- linear gates on share vectors (authenticated_scalar.rs:457-948, share.rs:74-131) ---- */
int arkmpc_fr_share_add(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                        const uint64_t* b_share, const uint64_t* b_mac, uint64_t* out_share, uint64_t* out_mac);
int arkmpc_fr_share_sub(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                        const uint64_t* b_share, const uint64_t* b_mac, uint64_t* out_share, uint64_t* out_mac);
int arkmpc_fr_share_neg(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                        uint64_t* out_share, uint64_t* out_mac);
/* batch_add_public / batch_sub_public (:493-528, :692-745): share += v on party 0 only; mac += key*v */
int arkmpc_fr_share_add_public(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* v,
                               uint64_t* out_share, uint64_t* out_mac);
int arkmpc_fr_share_sub_public(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* v,
                               uint64_t* out_share, uint64_t* out_mac);
/* batch_mul_public / batch_mul_constant (:883-948): share*v, mac*v */
int arkmpc_fr_share_mul_public(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                               const uint64_t* v, uint64_t* out_share, uint64_t* out_mac);

/* ---- open_authenticated_batch pieces (:278-354) ---- */
/* check[i] = key * opened[i] - mac[i]   (:299-311) */
int arkmpc_fr_mac_check(arkmpc_ctx* ctx, int field, const uint64_t* key_host, size_t n,
                        const uint64_t* opened, const uint64_t* mac, uint64_t* check);
/* *all_zero_host = 1 iff mine[i] + peer[i] == 0 for every i  (:217-219).  Synchronous. */
int arkmpc_fr_sum_is_zero(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* mine, const uint64_t* peer, int* all_zero_host);
/* *all_canonical_host = 1 iff a[i] < p for every i: the check arkworks' deserialisation applies to values that arrive from
 * the peer (scalar.rs:187-202) before they may be used as d_peer / e_peer / peer check shares.  Synchronous. */
int arkmpc_fr_validate(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, int* all_canonical_host);
/* canonical big-endian 32-byte encoding of each element for the hash commitment (scalar.rs:118-127) */
int arkmpc_fr_to_bytes_be(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint8_t* out_dev);

/* ---- Sum (share.rs:104-111): out_share/out_mac are single device scalars; scratch managed by ctx ---- */
int arkmpc_fr_share_sum(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a_share, const uint64_t* a_mac,
                        uint64_t* out_share, uint64_t* out_mac);
int arkmpc_fr_sum(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out);

/* ---- batch inversion and FFT on shares ----
 * out[i] = a[i]^-1, zeros stay zero: `Scalar::batch_inverse` (scalar.rs:93-100 -> ark_ff::batch_inversion), the public step of
 * AuthenticatedScalarResult::batch_inverse (authenticated_scalar.rs:55-82). */
int arkmpc_fr_batch_inverse(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* a, uint64_t* out);
/* ark-poly Radix2EvaluationDomain::fft / ifft over a domain of size 2^log2n (the caller zero-pads, as `D::new(x.len())` does):
 * out[j] = sum_i in[i] w^(ij), w = TWO_ADIC_ROOT_OF_UNITY^(2^(28-log2n)); inverse != 0 gives the inverse transform (scaled by n^-1).
 * Out of place (in != out); natural order in and out.  BN254 Fr only (ARKMPC_ERR_UNSUPPORTED otherwise: Curve25519 Fr has
 * two-adicity 2).  arkmpc_fr_share_fft transforms the share and the mac plane (share.rs:162-192, authenticated_scalar.rs:1011-1070). */
int arkmpc_fr_fft(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in, uint64_t* out);
int arkmpc_fr_share_fft(arkmpc_ctx* ctx, int field, int log2n, int inverse, const uint64_t* in_share, const uint64_t* in_mac,
                        uint64_t* out_share, uint64_t* out_mac);

/* ---- representation helpers ---- */
int arkmpc_fr_to_mont(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* plain, uint64_t* mont);   /* x -> x*R */
int arkmpc_fr_from_mont(arkmpc_ctx* ctx, int field, size_t n, const uint64_t* mont, uint64_t* plain); /* x*R -> x */
/* Deterministic synthetic elements, uniform in [0,p): element i = synth(seed, first_index + i)
 * (`Scalar::random` stand-in for benches, scalar.rs:76-78); identical to oracle synth generators. */
int arkmpc_fr_random(arkmpc_ctx* ctx, int field, uint64_t seed, uint64_t first_index, size_t n, uint64_t* out);

/* ---- multi-GPU: batch_open results gathered on every device (north_star's one collective) ----
 * The batch is sharded by contiguous index range, one process per GPU; gates need no inter-GPU traffic.  When the
 * opened d || e of a batch_mul must exist on every device, each rank owns two gathered planes of world*n scalars.
 * Either run arkmpc_fr_beaver_recombine with d_open/e_open pointing at this rank's row block and then an NCCL
 * all-gather (host side: torch.distributed), or use the fused entry point below, which stores the opened values into
 * every rank's planes from inside the kernel through CUDA IPC peer mappings.  Buffers exported with arkmpc_ipc_export
 * must come from arkmpc_malloc (whole allocations).  The caller synchronises ranks (stream sync + barrier) before
 * reading gathered rows written by peers. */
#define ARKMPC_IPC_HANDLE_BYTES 64
#define ARKMPC_MAX_PEERS 8
int arkmpc_ipc_export(arkmpc_ctx* ctx, void* dev_ptr, uint8_t* handle_out /* ARKMPC_IPC_HANDLE_BYTES */);
int arkmpc_ipc_import(arkmpc_ctx* ctx, const uint8_t* handle, void** peer_dev_ptr);
int arkmpc_ipc_release(arkmpc_ctx* ctx, void* peer_dev_ptr);
int arkmpc_fr_beaver_recombine_gather(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                      const uint64_t* d_mine, const uint64_t* e_mine,
                                      const uint64_t* d_peer, const uint64_t* e_peer,
                                      const uint64_t* a_share, const uint64_t* a_mac,
                                      const uint64_t* b_share, const uint64_t* b_mac,
                                      const uint64_t* c_share, const uint64_t* c_mac,
                                      uint64_t* out_share, uint64_t* out_mac,
                                      int world, int rank,
                                      uint64_t* const* gather_d /* [world] base of rank k's (world*n)-scalar plane */,
                                      uint64_t* const* gather_e);

/* The same gather with NVSwitch MULTICAST stores: the gathered planes of all ranks live in one multicast window, and each
 * opened element leaves the GPU once (multimem.st); the switch replicates it into every rank's copy (this rank's included).
 * Window life cycle (collective over the ranks of one NVSwitch domain; the host provides the barriers):
 *   1. rank 0: arkmpc_mc_open(owner_pid = 0) creates the multicast object; arkmpc_mc_export gives (pid, fd);
 *      the host hands both integers to the other ranks (any channel), which call arkmpc_mc_open(owner_pid, owner_fd) —
 *      the descriptor is duplicated with pidfd_getfd; owner_pid < 0 means owner_fd is already a descriptor of this process.
 *   2. barrier; every rank: arkmpc_mc_bind -> local_ptr (this rank's copy, ordinary device memory) and multicast_ptr
 *      (stores through it land in EVERY rank's copy at the same offset).  barrier.
 *   3. use: arkmpc_fr_beaver_recombine_gather_mc / arkmpc_mc_allgather_rows with pointers derived from multicast_ptr; read
 *      the gathered rows through local_ptr after a stream sync + barrier.
 *   4. stream sync + barrier; arkmpc_mc_close.
 * ARKMPC_ERR_UNSUPPORTED (and *supported = 0) when the device or driver has no multicast support: fall back to
 * arkmpc_fr_beaver_recombine_gather or arkmpc_allgather_open. */
typedef struct arkmpc_mc arkmpc_mc;
int arkmpc_mc_supported(arkmpc_ctx* ctx, int* supported);
int arkmpc_mc_open(arkmpc_ctx* ctx, size_t bytes, int world, int rank, int owner_pid, int owner_fd, arkmpc_mc** out);
int arkmpc_mc_export(arkmpc_mc* window, int* pid, int* fd);
int arkmpc_mc_bind(arkmpc_mc* window, void** local_ptr, void** multicast_ptr);
int arkmpc_mc_close(arkmpc_mc* window);
int arkmpc_fr_beaver_recombine_gather_mc(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                         const uint64_t* d_mine, const uint64_t* e_mine,
                                         const uint64_t* d_peer, const uint64_t* e_peer,
                                         const uint64_t* a_share, const uint64_t* a_mac,
                                         const uint64_t* b_share, const uint64_t* b_mac,
                                         const uint64_t* c_share, const uint64_t* c_mac,
                                         uint64_t* out_share, uint64_t* out_mac, int world, int rank,
                                         uint64_t* gather_d_multicast /* multicast address of the (world*n)-scalar d plane */,
                                         uint64_t* gather_e_multicast);
/* rows [rank*n, (rank+1)*n) of every rank's gathered plane <- local_rows (no arithmetic) */
int arkmpc_mc_allgather_rows(arkmpc_ctx* ctx, size_t n, int rank, const uint64_t* local_rows, uint64_t* gathered_multicast);

/* The plain collective (SURVEY §8b `arkmpc_allgather_open`): ncclAllGather of this rank's opened d and e rows into the
 * gathered planes of every rank, on the context's stream.  NCCL (libnccl.so.2) is resolved at run time; rank 0 obtains an id
 * with arkmpc_nccl_unique_id, the host distributes its ARKMPC_NCCL_ID_BYTES bytes, every rank calls arkmpc_nccl_init
 * (collective).  e_local / e_all may both be NULL to gather one plane.  ARKMPC_ERR_UNSUPPORTED if NCCL cannot be loaded,
 * ARKMPC_ERR_NCCL if a NCCL call fails. */
#define ARKMPC_NCCL_ID_BYTES 128
int arkmpc_nccl_unique_id(uint8_t* id_out /* ARKMPC_NCCL_ID_BYTES */);
int arkmpc_nccl_init(arkmpc_ctx* ctx, int world, int rank, const uint8_t* id);
int arkmpc_nccl_destroy(arkmpc_ctx* ctx);
int arkmpc_allgather_open(arkmpc_ctx* ctx, size_t n_local, const uint64_t* d_local, const uint64_t* e_local,
                          uint64_t* d_all, uint64_t* e_all);

/* ---- point gates (algebra/curve/authenticated_curve.rs, curve/share.rs, curve/curve.rs) ----
 * Points are DEVICE arrays in the reference's AoS memory image: BN254 `G1Projective` {x,y,z} = 96 B (Jacobian,
 * identity z = 0); Curve25519 `EdwardsProjective` {x,y,t,z} = 128 B (extended twisted Edwards); every coordinate a
 * canonical Montgomery residue of the base field.  A vector of n `PointShare`s (curve/share.rs:25-30) is n consecutive
 * {share, mac} point pairs (`*_ps` arguments, 2n points).  Scalars / scalar shares are planes as above.
 * Projective outputs are valid representatives, not necessarily arkworks' (curve.rs:46 compares projectively);
 * arkmpc_pt_normalize gives the canonical affine form parity is defined on.  Precondition of the fused Beaver
 * recombination: points lie in the prime-order subgroup (multiples of the generator), as every honestly shared
 * point does.  All arrays must be 32-byte aligned. */
size_t arkmpc_point_bytes(int curve); /* 96 / 128; 0 for an unknown curve */

/* CurvePointResult + CurvePointResult, the open-add (:98-108), and on 2n points AuthenticatedPointResult::batch_add /
 * batch_sub (:396-421, :520-545): out[i] = a[i] +/- b[i] */
int arkmpc_pt_add(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_pts, const uint64_t* b_pts, uint64_t* out_pts);
int arkmpc_pt_sub(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_pts, const uint64_t* b_pts, uint64_t* out_pts);
int arkmpc_pt_neg(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_pts, uint64_t* out_pts); /* batch_neg :604-621 */
/* batch_add_public / batch_sub_public (:429-465, :553-590; curve/share.rs:57-65) */
int arkmpc_pt_share_add_public(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* a_ps, const uint64_t* pub_pts, uint64_t* out_ps);
int arkmpc_pt_share_sub_public(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* a_ps, const uint64_t* pub_pts, uint64_t* out_ps);
/* CurvePointResult::batch_mul (curve.rs:459-479): out[i] = s[i] * P[i] */
int arkmpc_pt_mul(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out_pts);
/* AuthenticatedPointResult::batch_mul_public (:718-751): out[i] = (s[i]*share[i], s[i]*mac[i]) */
int arkmpc_pt_share_mul_public(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* a_ps, uint64_t* out_ps);
/* CurvePointResult::batch_mul_authenticated (curve.rs:483-517): out[i] = (s_share[i]*P[i], s_mac[i]*P[i]) */
int arkmpc_pt_mul_authenticated(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac,
                                const uint64_t* pts, uint64_t* out_ps);
/* batch_mul_generator (:754-780): out[i] = (s_share[i]*G, s_mac[i]*G); arkmpc_pt_mul_generator_public: out[i] = s[i]*G */
int arkmpc_pt_mul_generator(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac, uint64_t* out_ps);
int arkmpc_pt_mul_generator_public(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, uint64_t* out_pts);

/* AuthenticatedPointResult::batch_mul (:682-714), phase 1: d_mine = x.share - a.share (scalar plane),
 * E_mine = P.share - b.share*G (n points; what open_batch sends, :66-109). */
int arkmpc_pt_beaver_mask(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* x_share, const uint64_t* P_ps,
                          const uint64_t* a_share, const uint64_t* b_share, uint64_t* d_mine, uint64_t* E_mine_pts);
/* phase 2: d = d_mine + d_peer, E = E_mine + E_peer, out = d*E (public, add_public) + d*[bG] + [a]*E + [c]*G as n
 * PointShares.  d_open / E_open_pts optional (both or neither). */
int arkmpc_pt_beaver_recombine(arkmpc_ctx* ctx, int curve, int party_id, const uint64_t* key_host, size_t n,
                               const uint64_t* d_mine, const uint64_t* d_peer, const uint64_t* E_mine_pts, const uint64_t* E_peer_pts,
                               const uint64_t* a_share, const uint64_t* a_mac, const uint64_t* b_share, const uint64_t* b_mac,
                               const uint64_t* c_share, const uint64_t* c_mac, uint64_t* out_ps,
                               uint64_t* d_open, uint64_t* E_open_pts);

/* open_authenticated_batch on points (:193-283): check[i] = key * opened[i] - mac_i (mac_i from the PointShare vector) */
int arkmpc_pt_mac_check(arkmpc_ctx* ctx, int curve, const uint64_t* key_host, size_t n, const uint64_t* opened_pts,
                        const uint64_t* a_ps, uint64_t* check_pts);
/* *all_identity_host = 1 iff mine[i] + peer[i] is the identity for every i (:128-131).  Synchronous. */
int arkmpc_pt_sum_is_identity(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* mine_pts, const uint64_t* peer_pts, int* all_identity_host);
/* Sums (CurvePoint / PointShare `Sum`, the fold of AuthenticatedPointResult::msm :798-803): out = sum of n points (the identity for
 * n == 0); arkmpc_pt_share_sum sums the share and the mac points of n PointShares into one PointShare. */
int arkmpc_pt_sum(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, uint64_t* out_pt);
int arkmpc_pt_share_sum(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_ps, uint64_t* out_ps);
/* Public multiscalar multiplication sum_i s[i] * P[i]: `CurvePoint::msm` (curve.rs:549-560, which calls ark-ec's Pippenger), and
 * `CurvePoint::msm_authenticated` (curve.rs:619-642): out = (sum share[i]*P[i], sum mac[i]*P[i]).  Bucket method on the device
 * (csrc/curve_msm.cuh); below 2^15 points, parallel scalar multiplications and a sum.  Scratch is allocated stream-ordered. */
int arkmpc_pt_msm(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out_pt);
int arkmpc_pt_msm_authenticated(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* s_share, const uint64_t* s_mac,
                                const uint64_t* pts, uint64_t* out_ps);
/* *all_valid_host = 1 iff every point is on the curve AND in the prime-order subgroup ([r]P = identity; BN254 G1 has cofactor
 * 1, Curve25519 has cofactor 8): the check arkworks' deserialisation applies to points that arrive from the peer
 * (curve.rs:105-135).  arkmpc_pt_beaver_recombine regroups the reference's scalar multiplications as
 * ((a + d) mod r) * E, which equals a*E + d*E only for E in the prime-order subgroup: a shim validates E_peer (or takes it
 * from a validating deserialiser) before the recombination.  One scalar multiplication per point on Curve25519.  Synchronous. */
int arkmpc_pt_validate(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, int* all_valid_host);
/* PointShare vector <-> separate vectors of share points and mac points (either output of split may be NULL) */
int arkmpc_pt_share_split(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* a_ps, uint64_t* share_pts, uint64_t* mac_pts);
int arkmpc_pt_share_join(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* share_pts, const uint64_t* mac_pts, uint64_t* out_ps);
/* Canonical affine form: out_xy[i] = (x, y) Montgomery, 64 B; the BN254 identity maps to (0,0), the Edwards identity is (0,1). */
int arkmpc_pt_normalize(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* pts, uint64_t* out_xy);
/* The inverse: affine images -> points in the reference's projective memory image with Z = 1 (BN254: (0,0) -> identity).  Used to
 * ingest PointBatch payloads from the wire (ark_mpc_b200/wire.py); the caller validates peer points with arkmpc_pt_validate. */
int arkmpc_pt_from_affine(arkmpc_ctx* ctx, int curve, size_t n, const uint64_t* xy, uint64_t* out_pts);

/* ---- end-to-end over HOST buffers (reference AoS images), both protocol phases ----
 * begin: uploads x, y and the triple (a, b, c) (n ScalarShares each, AoS, host), runs the mask kernel and
 *        writes this party's d_mine || e_mine (2n scalars) to de_mine_host.  Synchronous on return.
 * finish: uploads the peer's d_peer || e_peer (2n scalars), runs the fused recombine kernel and writes n
 *        AoS ScalarShares to out_host (and, if non-NULL, the opened d || e to de_open_host).  Synchronous.
 * Transfers are chunked and overlapped with the kernels on internal streams. */
typedef struct arkmpc_batch_mul arkmpc_batch_mul;
int arkmpc_fr_batch_mul_begin_host(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                   const uint64_t* x_host, const uint64_t* y_host, const uint64_t* a_host,
                                   const uint64_t* b_host, const uint64_t* c_host, uint64_t* de_mine_host,
                                   arkmpc_batch_mul** session);
/* The same with x and y given as planes of their SHARE halves (n x 32 bytes each): the MACs of the operands are not inputs
 * of a Beaver multiplication (open_batch sends share.share() only, authenticated_scalar.rs:141-145; the product's MAC comes
 * from the triple's), so a host that keeps — or receives — the shares on their own uploads 320 instead of 384 bytes per gate. */
int arkmpc_fr_batch_mul_begin_host_shares(arkmpc_ctx* ctx, int field, int party_id, const uint64_t* key_host, size_t n,
                                          const uint64_t* x_share_host, const uint64_t* y_share_host,
                                          const uint64_t* a_host, const uint64_t* b_host, const uint64_t* c_host,
                                          uint64_t* de_mine_host, arkmpc_batch_mul** session);
int arkmpc_fr_batch_mul_finish_host(arkmpc_batch_mul* session, const uint64_t* de_peer_host, uint64_t* out_host,
                                    uint64_t* de_open_host);
int arkmpc_fr_batch_mul_abort(arkmpc_batch_mul* session);
/* Bytes one party's begin + finish move over PCIe for n gates (x.share, y.share, a, b, c, the peer's d || e up; own d || e,
 * the result shares and, with_open, the opened d || e down).  begin reads the 32-byte share halves of x and y straight from
 * the caller's buffers when they are pinned (arkmpc_host_alloc / cudaHostRegister): the MAC halves never cross the link.
 * A failed begin leaves no session (*session = NULL); finish consumes the session whether it succeeds or not. */
int arkmpc_fr_batch_mul_host_bytes(arkmpc_ctx* ctx, size_t n, int with_open, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

#ifdef __cplusplus
}
#endif
#endif /* ARKMPC_B200_H */
