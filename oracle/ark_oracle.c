/* CPU oracle (plain C) for the ark-mpc online-phase hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  The product library (ark_mpc_b200/csrc) never links,
 * loads or calls this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / `--impl reference` legs of bench.py do.
 *
 * It restates, for the CPU, the reference's arithmetic and op sequence
 * (paths relative to /root/reference/online-phase/src):
 *   - Scalar<C> = ark-ff 0.4 Fp256<MontBackend<_,4>> (algebra/scalar/scalar.rs:46,210-286;
 *     dependency pinned "0.4" in online-phase/Cargo.toml:91, source NOT vendored):
 *     canonical Montgomery residues, R = 2^256, four LE u64 limbs; mul = CIOS,
 *     add = add + conditional subtract, neg(0) = 0.
 *   - ScalarShare{share,mac} AoS 64 B and its operators (algebra/scalar/share.rs:32-131)
 *   - the UNFUSED batch_mul gate sequence (algebra/scalar/authenticated_scalar.rs:848-879):
 *     2x batch_sub (:662-688, share AND mac via self + (-rhs)), open add (:161-171),
 *     ScalarResult::batch_mul (scalar_result.rs:257-278), 2x batch_mul_public (:883-916),
 *     batch_add_public (:493-528), 2x batch_add (:457-489), every intermediate vector
 *     materialised as the reference's executor does.
 *   - MAC-check share (:299-311) and sum (share.rs:104-111).
 * It is validated against oracle/pyoracle.py (exact big-int) in tests/test_oracle.py.
 * It omits the reference executor's per-element ResultValue clone/insert bookkeeping
 * (fabric/executor/single_threaded.rs:334-373), so as a timed baseline it FLATTERS the reference.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { fe share, mac; } sshare; /* share.rs:32-37 */

typedef struct {
  fe p;          /* modulus */
  uint64_t inv;  /* -p^-1 mod 2^64 */
  fe r, r2;      /* R mod p, R^2 mod p */
  int bits;
} field_t;

static const field_t FIELDS[4] = {
    /* 0: BN254 Fr */
    {{{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
     0xc2e1f593efffffffull,
     {{0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}},
     {{0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull}},
     254},
    /* 1: Curve25519 Fr (l = 2^252 + 27742317777372353535851937790883648493) */
    {{{0x5812631a5cf5d3edull, 0x14def9dea2f79cd6ull, 0x0000000000000000ull, 0x1000000000000000ull}},
     0xd2b51da312547e1bull,
     {{0xd6ec31748d98951dull, 0xc6ef5bf4737dcf70ull, 0xfffffffffffffffeull, 0x0fffffffffffffffull}},
     {{0xa40611e3449c0f01ull, 0xd00e1ba768859347ull, 0xceec73d217f5be65ull, 0x0399411b7c309a3dull}},
     253},
    /* 2: BN254 Fq */
    {{{0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
     0x87d20782e4866389ull,
     {{0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full}},
     {{0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full}},
     254},
    /* 3: Curve25519 Fq (2^255 - 19) */
    {{{0xffffffffffffffedull, 0xffffffffffffffffull, 0xffffffffffffffffull, 0x7fffffffffffffffull}},
     0x86bca1af286bca1bull,
     {{0x26ull, 0, 0, 0}},
     {{0x5a4ull, 0, 0, 0}},
     255},
};

int orc_num_fields(void) { return 4; }
const uint64_t* orc_field_modulus(int f) { return FIELDS[f].p.l; }
const uint64_t* orc_field_r(int f) { return FIELDS[f].r.l; }
const uint64_t* orc_field_r2(int f) { return FIELDS[f].r2.l; }
uint64_t orc_field_inv(int f) { return FIELDS[f].inv; }

/* ---- field arithmetic (ark-ff MontBackend semantics) ---- */
static inline int fe_geq(const fe* a, const fe* b) {
  for (int i = 3; i >= 0; i--) {
    if (a->l[i] > b->l[i]) return 1;
    if (a->l[i] < b->l[i]) return 0;
  }
  return 1;
}
static inline uint64_t fe_sub_raw(fe* r, const fe* a, const fe* b) {
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a->l[i] - b->l[i] - (uint64_t)br;
    r->l[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  return (uint64_t)br;
}
static inline void fe_add(const field_t* F, fe* r, const fe* a, const fe* b) {
  u128 c = 0;
  fe t;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    t.l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || fe_geq(&t, &F->p)) fe_sub_raw(&t, &t, &F->p);
  *r = t;
}
static inline void fe_neg(const field_t* F, fe* r, const fe* a) {
  if ((a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0) { *r = *a; return; }
  fe_sub_raw(r, &F->p, a);
}
static inline void fe_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
  fe t;
  if (fe_sub_raw(&t, a, b)) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t.l[i] + F->p.l[i];
      t.l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  *r = t;
}
/* CIOS Montgomery multiplication, 4 x u64 (the algorithm of ark-ff's MontBackend::mul_assign) */
static inline void fe_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * F->inv;
    c = ((u128)m * F->p.l[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * F->p.l[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  fe o = {{t[0], t[1], t[2], t[3]}};
  if (t[4] || fe_geq(&o, &F->p)) fe_sub_raw(&o, &o, &F->p);
  *r = o;
}

/* ---- element-level exports (used by the tests to cross-check against pyoracle) ---- */
void orc_fe_add(int f, uint64_t* r, const uint64_t* a, const uint64_t* b) { fe_add(&FIELDS[f], (fe*)r, (const fe*)a, (const fe*)b); }
void orc_fe_sub(int f, uint64_t* r, const uint64_t* a, const uint64_t* b) { fe_sub(&FIELDS[f], (fe*)r, (const fe*)a, (const fe*)b); }
void orc_fe_neg(int f, uint64_t* r, const uint64_t* a) { fe_neg(&FIELDS[f], (fe*)r, (const fe*)a); }
void orc_fe_mul(int f, uint64_t* r, const uint64_t* a, const uint64_t* b) { fe_mul(&FIELDS[f], (fe*)r, (const fe*)a, (const fe*)b); }
/* plain integer (4 limbs, < p) -> Montgomery image and back */
void orc_to_mont(int f, size_t n, uint64_t* out, const uint64_t* in) {
  for (size_t i = 0; i < n; i++) fe_mul(&FIELDS[f], (fe*)(out + 4 * i), (const fe*)(in + 4 * i), &FIELDS[f].r2);
}
void orc_from_mont(int f, size_t n, uint64_t* out, const uint64_t* in) {
  fe one = {{1, 0, 0, 0}};
  for (size_t i = 0; i < n; i++) fe_mul(&FIELDS[f], (fe*)(out + 4 * i), (const fe*)(in + 4 * i), &one);
}

/* ---- deterministic synthetic elements (same generator as pyoracle.synth_element
 *      and the product's CUDA generator; values are Montgomery images) ---- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline void synth_one(const field_t* F, uint64_t seed, uint64_t index, fe* out) {
  const uint64_t top_mask = (1ull << (F->bits - 192)) - 1;
  for (uint64_t t = 0;; t++) {
    fe v;
    for (int j = 0; j < 4; j++) v.l[j] = splitmix64((seed ^ splitmix64(index * 4 + (uint64_t)j)) + t * 0xD1342543DE82EF95ull);
    v.l[3] &= top_mask;
    if (!fe_geq(&v, &F->p)) { *out = v; return; }
  }
}
void orc_synth(int f, uint64_t seed, uint64_t first_index, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; i++) synth_one(&FIELDS[f], seed, first_index + i, (fe*)(out + 4 * i));
}

/* ---- batch gates in the reference's unfused form (AoS) ---- */
static void batch_sub(const field_t* F, size_t n, sshare* o, const sshare* a, const sshare* b) {
  for (size_t i = 0; i < n; i++) { /* share.rs:95-101: self + (-rhs) on share and mac */
    fe t;
    fe_neg(F, &t, &b[i].share); fe_add(F, &o[i].share, &a[i].share, &t);
    fe_neg(F, &t, &b[i].mac);   fe_add(F, &o[i].mac, &a[i].mac, &t);
  }
}
static void batch_add(const field_t* F, size_t n, sshare* o, const sshare* a, const sshare* b) {
  for (size_t i = 0; i < n; i++) { fe_add(F, &o[i].share, &a[i].share, &b[i].share); fe_add(F, &o[i].mac, &a[i].mac, &b[i].mac); }
}
static void batch_mul_public(const field_t* F, size_t n, sshare* o, const sshare* a, const fe* s) {
  for (size_t i = 0; i < n; i++) { fe_mul(F, &o[i].share, &a[i].share, &s[i]); fe_mul(F, &o[i].mac, &a[i].mac, &s[i]); }
}
static void batch_add_public(const field_t* F, int party, const fe* key, size_t n, sshare* o, const sshare* a, const fe* v) {
  for (size_t i = 0; i < n; i++) { /* share.rs:74-77 */
    fe kv;
    fe_mul(F, &kv, key, &v[i]);
    if (party == 0) fe_add(F, &o[i].share, &a[i].share, &v[i]); else o[i].share = a[i].share;
    fe_add(F, &o[i].mac, &a[i].mac, &kv);
  }
}
static void scalar_batch_mul(const field_t* F, size_t n, fe* o, const fe* a, const fe* b) {
  for (size_t i = 0; i < n; i++) fe_mul(F, &o[i], &a[i], &b[i]);
}
static void scalar_batch_add(const field_t* F, size_t n, fe* o, const fe* a, const fe* b) {
  for (size_t i = 0; i < n; i++) fe_add(F, &o[i], &a[i], &b[i]);
}

void orc_batch_add(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b) { batch_add(&FIELDS[f], n, (sshare*)o, (const sshare*)a, (const sshare*)b); }
void orc_batch_sub(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b) { batch_sub(&FIELDS[f], n, (sshare*)o, (const sshare*)a, (const sshare*)b); }
void orc_batch_neg(int f, size_t n, uint64_t* o, const uint64_t* a) {
  for (size_t i = 0; i < n; i++) { fe_neg(&FIELDS[f], &((sshare*)o)[i].share, &((const sshare*)a)[i].share); fe_neg(&FIELDS[f], &((sshare*)o)[i].mac, &((const sshare*)a)[i].mac); }
}
void orc_batch_mul_public(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* s) { batch_mul_public(&FIELDS[f], n, (sshare*)o, (const sshare*)a, (const fe*)s); }
void orc_batch_add_public(int f, int party, const uint64_t* key, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* v) {
  batch_add_public(&FIELDS[f], party, (const fe*)key, n, (sshare*)o, (const sshare*)a, (const fe*)v);
}
void orc_batch_sub_public(int f, int party, const uint64_t* key, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* v) {
  const field_t* F = &FIELDS[f]; /* share.rs:80-82: add_public(-rhs) */
  for (size_t i = 0; i < n; i++) {
    fe nv; fe_neg(F, &nv, &((const fe*)v)[i]);
    batch_add_public(F, party, (const fe*)key, 1, (sshare*)o + i, (const sshare*)a + i, &nv);
  }
}
void orc_scalar_batch_mul(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b) { scalar_batch_mul(&FIELDS[f], n, (fe*)o, (const fe*)a, (const fe*)b); }
void orc_scalar_batch_add(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b) { scalar_batch_add(&FIELDS[f], n, (fe*)o, (const fe*)a, (const fe*)b); }
void orc_scalar_batch_sub(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b) {
  for (size_t i = 0; i < n; i++) fe_sub(&FIELDS[f], (fe*)o + i, (const fe*)a + i, (const fe*)b + i);
}

/* own share components of d = [x - a], e = [y - b]  (authenticated_scalar.rs:863-867, :141-145).
 * Runs the reference's full batch_sub (share and mac) and then extracts the share halves. */
void orc_beaver_mask(int f, size_t n, const uint64_t* x, const uint64_t* y, const uint64_t* a, const uint64_t* b,
                     uint64_t* d_mine, uint64_t* e_mine, uint64_t* scratch /* 2n shares */) {
  const field_t* F = &FIELDS[f];
  sshare* ml = (sshare*)scratch;
  sshare* mr = ml + n;
  batch_sub(F, n, ml, (const sshare*)x, (const sshare*)a);
  batch_sub(F, n, mr, (const sshare*)y, (const sshare*)b);
  for (size_t i = 0; i < n; i++) { ((fe*)d_mine)[i] = ml[i].share; ((fe*)e_mine)[i] = mr[i].share; }
}

/* :871-878 with every intermediate materialised; scratch holds n scalars + 5n shares */
void orc_beaver_recombine(int f, int party, const uint64_t* key, size_t n, const uint64_t* d, const uint64_t* e,
                          const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* out, uint64_t* scratch) {
  const field_t* F = &FIELDS[f];
  fe* de = (fe*)scratch;
  sshare* db = (sshare*)(de + n);
  sshare* ea = db + n;
  sshare* de_db = ea + n;
  sshare* ea_c = de_db + n;
  scalar_batch_mul(F, n, de, (const fe*)d, (const fe*)e);
  batch_mul_public(F, n, db, (const sshare*)b, (const fe*)d);
  batch_mul_public(F, n, ea, (const sshare*)a, (const fe*)e);
  batch_add_public(F, party, (const fe*)key, n, de_db, db, de);
  batch_add(F, n, ea_c, ea, (const sshare*)c);
  batch_add(F, n, (sshare*)out, de_db, ea_c);
}

void orc_mac_check(int f, const uint64_t* key, size_t n, const uint64_t* opened, const uint64_t* shares, uint64_t* out) {
  const field_t* F = &FIELDS[f]; /* :299-311: mac_key * value - share.mac */
  for (size_t i = 0; i < n; i++) {
    fe kv; fe_mul(F, &kv, (const fe*)key, (const fe*)opened + i);
    fe_sub(F, (fe*)out + i, &kv, &((const sshare*)shares)[i].mac);
  }
}
void orc_share_sum(int f, size_t n, const uint64_t* shares, uint64_t* out) {
  const field_t* F = &FIELDS[f]; /* share.rs:104-111 */
  sshare acc; memset(&acc, 0, sizeof acc);
  for (size_t i = 0; i < n; i++) { fe_add(F, &acc.share, &acc.share, &((const sshare*)shares)[i].share); fe_add(F, &acc.mac, &acc.mac, &((const sshare*)shares)[i].mac); }
  *(sshare*)out = acc;
}

/* ---- two-party batch_mul with a static index partition over `threads` host threads.
 * Thread t owns [lo,hi) of BOTH parties (the mock network moves the d/e vectors by pointer,
 * network/mock.rs:131-134, so the exchange is a read of the peer's vector). ---- */
typedef struct {
  int f; size_t lo, hi;
  const uint64_t* key[2]; const uint64_t* x[2]; const uint64_t* y[2];
  const uint64_t* a[2]; const uint64_t* b[2]; const uint64_t* c[2];
  uint64_t* out[2]; uint64_t* d_open; uint64_t* e_open;
} job_t;

static void* job_run(void* arg) {
  job_t* j = (job_t*)arg;
  const field_t* F = &FIELDS[j->f];
  size_t n = j->hi - j->lo;
  if (!n) return NULL;
  /* per party: d_mine, e_mine (2n fe) ; scratch: 2n shares mask + (n fe + 5n shares) recombine */
  fe* dm[2]; fe* em[2];
  for (int p = 0; p < 2; p++) { dm[p] = (fe*)malloc(n * sizeof(fe)); em[p] = (fe*)malloc(n * sizeof(fe)); }
  uint64_t* scratch = (uint64_t*)malloc(n * (sizeof(fe) + 5 * sizeof(sshare)));
  fe* d = (fe*)malloc(n * sizeof(fe));
  fe* e = (fe*)malloc(n * sizeof(fe));
  for (int p = 0; p < 2; p++)
    orc_beaver_mask(j->f, n, j->x[p] + 8 * j->lo, j->y[p] + 8 * j->lo, j->a[p] + 8 * j->lo, j->b[p] + 8 * j->lo,
                    (uint64_t*)dm[p], (uint64_t*)em[p], scratch);
  /* each party adds own + peer; identical values, computed once per party in the reference.
   * We compute it for both parties to keep the per-party work faithful. */
  for (int p = 0; p < 2; p++) {
    scalar_batch_add(F, n, d, dm[0], dm[1]);
    scalar_batch_add(F, n, e, em[0], em[1]);
    orc_beaver_recombine(j->f, p, j->key[p], n, (const uint64_t*)d, (const uint64_t*)e, j->a[p] + 8 * j->lo,
                         j->b[p] + 8 * j->lo, j->c[p] + 8 * j->lo, j->out[p] + 8 * j->lo, scratch);
  }
  if (j->d_open) memcpy(j->d_open + 4 * j->lo, d, n * sizeof(fe));
  if (j->e_open) memcpy(j->e_open + 4 * j->lo, e, n * sizeof(fe));
  for (int p = 0; p < 2; p++) { free(dm[p]); free(em[p]); }
  free(scratch); free(d); free(e);
  return NULL;
}

/* x0..c1: AoS share vectors of both parties; out0/out1: AoS result shares; d_open/e_open optional. */
int orc_two_party_batch_mul(int f, size_t n, int threads, const uint64_t* key0, const uint64_t* key1,
                            const uint64_t* x0, const uint64_t* y0, const uint64_t* a0, const uint64_t* b0, const uint64_t* c0,
                            const uint64_t* x1, const uint64_t* y1, const uint64_t* a1, const uint64_t* b1, const uint64_t* c1,
                            uint64_t* out0, uint64_t* out1, uint64_t* d_open, uint64_t* e_open) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  job_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < threads; t++) {
    job_t* j = &jobs[t];
    j->f = f; j->lo = n * (size_t)t / threads; j->hi = n * (size_t)(t + 1) / threads;
    j->key[0] = key0; j->key[1] = key1; j->x[0] = x0; j->x[1] = x1; j->y[0] = y0; j->y[1] = y1;
    j->a[0] = a0; j->a[1] = a1; j->b[0] = b0; j->b[1] = b1; j->c[0] = c0; j->c[1] = c1;
    j->out[0] = out0; j->out[1] = out1; j->d_open = d_open; j->e_open = e_open;
  }
  if (threads == 1) { job_run(&jobs[0]); return 0; }
  for (int t = 0; t < threads; t++) if (pthread_create(&th[t], NULL, job_run, &jobs[t])) return -1;
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* ====================================================================================================
 * Curve groups and the point Beaver multiplication (AuthenticatedPointResult::batch_mul,
 * algebra/curve/authenticated_curve.rs:682-714) in the reference's UNFUSED form: 4 fixed-base and 6
 * variable-base scalar multiplications per element and party, each a per-element MSB-first double-and-add
 * as ark-ec 0.4 `Projective * ScalarField` does (curve.rs:403-409 -> `mul_bigint`; crate not vendored).
 * Memory images: BN254 G1Projective {x,y,z} Jacobian (identity z = 0); Curve25519 EdwardsProjective
 * {x,y,t,z} extended twisted Edwards, a = -1.  Coordinates are Montgomery residues of the base field.
 * Projective representatives are not canonical; orc_pt_normalize gives the affine form results are compared on.
 * ==================================================================================================== */
typedef struct { fe c[4]; } pt_t; /* x,y,z,(unused) for curve 0 ; x,y,t,z for curve 1 */
typedef struct { int fq, fr, ncoord; } curve_t;
static const curve_t CURVES[2] = {{2, 0, 3}, {3, 1, 4}};
int orc_point_words(int curve) { return CURVES[curve].ncoord * 4; }

static const fe ED_2D = {{0x01db17fdbe8fd3f4ull, 0x21430eef5f8c52e7ull, 0xcb27240f78310d20ull, 0x590456b4e53f8a4dull}}; /* 2d * R mod p */
static const fe ED_GX = {{0xe2cabc553f9da287ull, 0x9ca598562396e489ull, 0x9879936bade4b5b7ull, 0x759e23707e6077d0ull}};
static const fe ED_GY = {{0x333333333333334aull, 0x3333333333333333ull, 0x3333333333333333ull, 0x3333333333333333ull}};

static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) { return a->l[0] == b->l[0] && a->l[1] == b->l[1] && a->l[2] == b->l[2] && a->l[3] == b->l[3]; }
static void fe_inv(const field_t* F, fe* r, const fe* a) { /* Fermat */
  fe e = F->p, acc = F->r;
  e.l[0] -= 2;
  for (int i = 255; i >= 0; i--) {
    fe_mul(F, &acc, &acc, &acc);
    if ((e.l[i >> 6] >> (i & 63)) & 1) fe_mul(F, &acc, &acc, a);
  }
  *r = acc;
}

static void pt_identity(int cv, pt_t* p) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  memset(p, 0, sizeof *p);
  if (cv == 0) { p->c[0] = F->r; p->c[1] = F->r; }           /* (1,1,0) */
  else { p->c[1] = F->r; p->c[3] = F->r; }                    /* (0,1,0,1) */
}
static void pt_generator(int cv, pt_t* p) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  memset(p, 0, sizeof *p);
  if (cv == 0) { p->c[0] = F->r; fe_add(F, &p->c[1], &F->r, &F->r); p->c[2] = F->r; }
  else { p->c[0] = ED_GX; p->c[1] = ED_GY; fe_mul(F, &p->c[2], &ED_GX, &ED_GY); p->c[3] = F->r; }
}
static int pt_is_identity(int cv, const pt_t* p) {
  if (cv == 0) return fe_is_zero(&p->c[2]);
  return fe_is_zero(&p->c[0]) && fe_eq(&p->c[1], &p->c[3]);
}
static void pt_neg(int cv, pt_t* r, const pt_t* p) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  *r = *p;
  if (cv == 0) fe_neg(F, &r->c[1], &p->c[1]);
  else { fe_neg(F, &r->c[0], &p->c[0]); fe_neg(F, &r->c[2], &p->c[2]); }
}
static void pt_dbl(int cv, pt_t* r, const pt_t* p) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  if (cv == 0) { /* dbl-2009-l */
    fe A, B, C, D, E, G, t, X3, Y3, Z3;
    fe_mul(F, &A, &p->c[0], &p->c[0]); fe_mul(F, &B, &p->c[1], &p->c[1]); fe_mul(F, &C, &B, &B);
    fe_add(F, &t, &p->c[0], &B); fe_mul(F, &t, &t, &t); fe_sub(F, &t, &t, &A); fe_sub(F, &t, &t, &C); fe_add(F, &D, &t, &t);
    fe_add(F, &E, &A, &A); fe_add(F, &E, &E, &A); fe_mul(F, &G, &E, &E);
    fe_add(F, &t, &D, &D); fe_sub(F, &X3, &G, &t);
    fe_sub(F, &t, &D, &X3); fe_mul(F, &t, &E, &t);
    fe_add(F, &C, &C, &C); fe_add(F, &C, &C, &C); fe_add(F, &C, &C, &C); fe_sub(F, &Y3, &t, &C);
    fe_mul(F, &Z3, &p->c[1], &p->c[2]); fe_add(F, &Z3, &Z3, &Z3);
    r->c[0] = X3; r->c[1] = Y3; r->c[2] = Z3;
  } else { /* dbl-2008-hwcd, a = -1 */
    fe A, B, C, E, G, Fv, H, t;
    fe_mul(F, &A, &p->c[0], &p->c[0]); fe_mul(F, &B, &p->c[1], &p->c[1]); fe_mul(F, &C, &p->c[3], &p->c[3]); fe_add(F, &C, &C, &C);
    fe_add(F, &t, &p->c[0], &p->c[1]); fe_mul(F, &E, &t, &t); fe_sub(F, &E, &E, &A); fe_sub(F, &E, &E, &B);
    fe_sub(F, &G, &B, &A); fe_sub(F, &Fv, &G, &C); fe_add(F, &H, &A, &B); fe_neg(F, &H, &H);
    fe_mul(F, &r->c[0], &E, &Fv); fe_mul(F, &r->c[1], &G, &H); fe_mul(F, &r->c[2], &E, &H); fe_mul(F, &r->c[3], &Fv, &G);
  }
}
static void pt_add(int cv, pt_t* r, const pt_t* p, const pt_t* q) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  if (cv == 0) { /* add-2007-bl with the exceptional cases */
    if (pt_is_identity(cv, q)) { *r = *p; return; }
    if (pt_is_identity(cv, p)) { *r = *q; return; }
    fe Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
    fe_mul(F, &Z1Z1, &p->c[2], &p->c[2]); fe_mul(F, &Z2Z2, &q->c[2], &q->c[2]);
    fe_mul(F, &U1, &p->c[0], &Z2Z2); fe_mul(F, &U2, &q->c[0], &Z1Z1);
    fe_mul(F, &S1, &p->c[1], &q->c[2]); fe_mul(F, &S1, &S1, &Z2Z2);
    fe_mul(F, &S2, &q->c[1], &p->c[2]); fe_mul(F, &S2, &S2, &Z1Z1);
    fe_sub(F, &H, &U2, &U1); fe_sub(F, &rr, &S2, &S1);
    if (fe_is_zero(&H)) { if (fe_is_zero(&rr)) pt_dbl(cv, r, p); else pt_identity(cv, r); return; }
    fe_add(F, &rr, &rr, &rr); fe_add(F, &I, &H, &H); fe_mul(F, &I, &I, &I); fe_mul(F, &J, &H, &I); fe_mul(F, &V, &U1, &I);
    fe_mul(F, &X3, &rr, &rr); fe_sub(F, &X3, &X3, &J); fe_add(F, &t, &V, &V); fe_sub(F, &X3, &X3, &t);
    fe_sub(F, &t, &V, &X3); fe_mul(F, &Y3, &rr, &t); fe_mul(F, &t, &S1, &J); fe_add(F, &t, &t, &t); fe_sub(F, &Y3, &Y3, &t);
    fe_add(F, &t, &p->c[2], &q->c[2]); fe_mul(F, &t, &t, &t); fe_sub(F, &t, &t, &Z1Z1); fe_sub(F, &t, &t, &Z2Z2); fe_mul(F, &Z3, &t, &H);
    r->c[0] = X3; r->c[1] = Y3; r->c[2] = Z3;
  } else { /* add-2008-hwcd-3, complete */
    fe A, B, C, D, E, Fv, G, H, t, u;
    fe_sub(F, &t, &p->c[1], &p->c[0]); fe_sub(F, &u, &q->c[1], &q->c[0]); fe_mul(F, &A, &t, &u);
    fe_add(F, &t, &p->c[1], &p->c[0]); fe_add(F, &u, &q->c[1], &q->c[0]); fe_mul(F, &B, &t, &u);
    fe_mul(F, &C, &p->c[2], &q->c[2]); fe_mul(F, &C, &C, &ED_2D);
    fe_mul(F, &D, &p->c[3], &q->c[3]); fe_add(F, &D, &D, &D);
    fe_sub(F, &E, &B, &A); fe_sub(F, &Fv, &D, &C); fe_add(F, &G, &D, &C); fe_add(F, &H, &B, &A);
    fe_mul(F, &r->c[0], &E, &Fv); fe_mul(F, &r->c[1], &G, &H); fe_mul(F, &r->c[2], &E, &H); fe_mul(F, &r->c[3], &Fv, &G);
  }
}
/* ark-ec `mul_bigint`: MSB-first double-and-add over the canonical integer of the scalar */
static void pt_mul(int cv, pt_t* r, const pt_t* p, const fe* s_mont) {
  const field_t* FR = &FIELDS[CURVES[cv].fr];
  fe one = {{1, 0, 0, 0}}, k;
  fe_mul(FR, &k, s_mont, &one);
  pt_t acc; pt_identity(cv, &acc);
  int started = 0;
  for (int i = 255; i >= 0; i--) {
    int bit = (int)((k.l[i >> 6] >> (i & 63)) & 1);
    if (!started && !bit) continue;
    started = 1;
    pt_dbl(cv, &acc, &acc);
    if (bit) pt_add(cv, &acc, &acc, p);
  }
  *r = acc;
}
static void pt_sub(int cv, pt_t* r, const pt_t* p, const pt_t* q) { pt_t n; pt_neg(cv, &n, q); pt_add(cv, r, p, &n); }

static void pt_load(int cv, pt_t* p, const uint64_t* src) { memset(p, 0, sizeof *p); memcpy(p, src, CURVES[cv].ncoord * 32); }
static void pt_store(int cv, uint64_t* dst, const pt_t* p) { memcpy(dst, p, CURVES[cv].ncoord * 32); }

/* affine (x,y), 64 B; BN254 identity -> (0,0) */
void orc_pt_normalize(int cv, size_t n, const uint64_t* pts, uint64_t* out_xy) {
  const field_t* F = &FIELDS[CURVES[cv].fq];
  const int w = CURVES[cv].ncoord * 4;
  for (size_t i = 0; i < n; i++) {
    pt_t p; pt_load(cv, &p, pts + i * w);
    fe x, y, zi, zi2;
    if (cv == 0) {
      if (pt_is_identity(cv, &p)) { memset(&x, 0, sizeof x); memset(&y, 0, sizeof y); }
      else { fe_inv(F, &zi, &p.c[2]); fe_mul(F, &zi2, &zi, &zi); fe_mul(F, &x, &p.c[0], &zi2); fe_mul(F, &zi2, &zi2, &zi); fe_mul(F, &y, &p.c[1], &zi2); }
    } else { fe_inv(F, &zi, &p.c[3]); fe_mul(F, &x, &p.c[0], &zi); fe_mul(F, &y, &p.c[1], &zi); }
    memcpy(out_xy + i * 8, &x, 32); memcpy(out_xy + i * 8 + 4, &y, 32);
  }
}
void orc_pt_generator(int cv, uint64_t* out) { pt_t g; pt_generator(cv, &g); pt_store(cv, out, &g); }
void orc_pt_mul(int cv, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out) {
  const int w = CURVES[cv].ncoord * 4;
  for (size_t i = 0; i < n; i++) { pt_t p, r; pt_load(cv, &p, pts + i * w); pt_mul(cv, &r, &p, (const fe*)scalars + i); pt_store(cv, out + i * w, &r); }
}
void orc_pt_mul_generator(int cv, size_t n, const uint64_t* scalars, uint64_t* out) {
  const int w = CURVES[cv].ncoord * 4;
  pt_t g; pt_generator(cv, &g);
  for (size_t i = 0; i < n; i++) { pt_t r; pt_mul(cv, &r, &g, (const fe*)scalars + i); pt_store(cv, out + i * w, &r); }
}
void orc_pt_add(int cv, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out, int sub) {
  const int w = CURVES[cv].ncoord * 4;
  for (size_t i = 0; i < n; i++) { pt_t p, q, r; pt_load(cv, &p, a + i * w); pt_load(cv, &q, b + i * w); if (sub) pt_sub(cv, &r, &p, &q); else pt_add(cv, &r, &p, &q); pt_store(cv, out + i * w, &r); }
}
/* PointShare::add_public (curve/share.rs:57-60) on AoS PointShares */
void orc_pt_share_add_public(int cv, int party, const uint64_t* key, size_t n, const uint64_t* a_ps, const uint64_t* pub, uint64_t* out_ps, int sub) {
  const int w = CURVES[cv].ncoord * 4;
  for (size_t i = 0; i < n; i++) {
    pt_t s, m, P, kp; pt_load(cv, &s, a_ps + i * 2 * w); pt_load(cv, &m, a_ps + i * 2 * w + w); pt_load(cv, &P, pub + i * w);
    if (sub) pt_neg(cv, &P, &P);
    if (party == 0) pt_add(cv, &s, &s, &P);
    pt_mul(cv, &kp, &P, (const fe*)key); pt_add(cv, &m, &m, &kp);
    pt_store(cv, out_ps + i * 2 * w, &s); pt_store(cv, out_ps + i * 2 * w + w, &m);
  }
}

/* One party, one element range, phase 1 (:696-700): bG = (b.share*G, b.mac*G) [kept for phase 2], d = x - a, E = P - bG.
 * The reference computes share AND mac of both masked values; only the share halves are opened. */
static void point_mask_range(int cv, size_t n, const sshare* x, const uint64_t* P_ps, const sshare* a, const sshare* b,
                             fe* d_mine, pt_t* E_mine, pt_t* bG_s, pt_t* bG_m) {
  const field_t* FR = &FIELDS[CURVES[cv].fr];
  const int w = CURVES[cv].ncoord * 4;
  pt_t g; pt_generator(cv, &g);
  for (size_t i = 0; i < n; i++) {
    pt_mul(cv, &bG_s[i], &g, &b[i].share); pt_mul(cv, &bG_m[i], &g, &b[i].mac);        /* batch_mul_generator :754-780 */
    sshare dm; batch_sub(FR, 1, &dm, &x[i], &a[i]); d_mine[i] = dm.share;                /* batch_sub :699 */
    pt_t Ps, Pm, Em; pt_load(cv, &Ps, P_ps + i * 2 * w); pt_load(cv, &Pm, P_ps + i * 2 * w + w);
    pt_sub(cv, &E_mine[i], &Ps, &bG_s[i]); pt_sub(cv, &Em, &Pm, &bG_m[i]);               /* batch_sub :700 (mac half unused by open) */
  }
}
/* phase 2 (:704-713) */
static void point_recombine_range(int cv, int party, const fe* key, size_t n, const fe* d, const pt_t* E, const sshare* a, const sshare* c,
                                  const pt_t* bG_s, const pt_t* bG_m, uint64_t* out_ps) {
  const int w = CURVES[cv].ncoord * 4;
  pt_t g; pt_generator(cv, &g);
  for (size_t i = 0; i < n; i++) {
    pt_t deG, dbs, dbm, aes, aem, cs, cm, kde, s, m;
    pt_mul(cv, &deG, &E[i], &d[i]);                                      /* CurvePointResult::batch_mul curve.rs:459-479 */
    pt_mul(cv, &dbs, &bG_s[i], &d[i]); pt_mul(cv, &dbm, &bG_m[i], &d[i]); /* batch_mul_public :718-751 */
    pt_mul(cv, &aes, &E[i], &a[i].share); pt_mul(cv, &aem, &E[i], &a[i].mac); /* batch_mul_authenticated curve.rs:483-517 */
    pt_mul(cv, &cs, &g, &c[i].share); pt_mul(cv, &cm, &g, &c[i].mac);     /* batch_mul_generator */
    s = dbs; if (party == 0) pt_add(cv, &s, &s, &deG);                    /* batch_add_public: curve/share.rs:57-60 */
    pt_mul(cv, &kde, &deG, key); pt_add(cv, &m, &dbm, &kde);
    pt_add(cv, &aes, &aes, &cs); pt_add(cv, &aem, &aem, &cm);             /* batch_add :711 */
    pt_add(cv, &s, &s, &aes); pt_add(cv, &m, &m, &aem);                   /* batch_add :713 */
    pt_store(cv, out_ps + i * 2 * w, &s); pt_store(cv, out_ps + i * 2 * w + w, &m);
  }
}

typedef struct {
  int cv; size_t lo, hi;
  const uint64_t* key[2]; const uint64_t* x[2]; const uint64_t* P[2]; const uint64_t* a[2]; const uint64_t* b[2]; const uint64_t* c[2];
  uint64_t* out[2]; uint64_t* d_open; uint64_t* E_open;
} pjob_t;

static void* pjob_run(void* arg) {
  pjob_t* j = (pjob_t*)arg;
  const int cv = j->cv;
  const field_t* FR = &FIELDS[CURVES[cv].fr];
  const int w = CURVES[cv].ncoord * 4;
  size_t n = j->hi - j->lo, lo = j->lo;
  if (!n) return NULL;
  fe* dm[2]; pt_t* Em[2]; pt_t* bs[2]; pt_t* bm[2];
  for (int p = 0; p < 2; p++) {
    dm[p] = (fe*)malloc(n * sizeof(fe)); Em[p] = (pt_t*)malloc(n * sizeof(pt_t)); bs[p] = (pt_t*)malloc(n * sizeof(pt_t)); bm[p] = (pt_t*)malloc(n * sizeof(pt_t));
    point_mask_range(cv, n, (const sshare*)j->x[p] + lo, j->P[p] + lo * 2 * w, (const sshare*)j->a[p] + lo, (const sshare*)j->b[p] + lo, dm[p], Em[p], bs[p], bm[p]);
  }
  fe* d = (fe*)malloc(n * sizeof(fe)); pt_t* E = (pt_t*)malloc(n * sizeof(pt_t));
  for (int p = 0; p < 2; p++) { /* each party opens for itself (:66-109, scalar :161-171) */
    scalar_batch_add(FR, n, d, dm[0], dm[1]);
    for (size_t i = 0; i < n; i++) pt_add(cv, &E[i], &Em[0][i], &Em[1][i]);
    point_recombine_range(cv, p, (const fe*)j->key[p], n, d, E, (const sshare*)j->a[p] + lo, (const sshare*)j->c[p] + lo, bs[p], bm[p], j->out[p] + lo * 2 * w);
  }
  if (j->d_open) memcpy(j->d_open + 4 * lo, d, n * sizeof(fe));
  if (j->E_open) for (size_t i = 0; i < n; i++) pt_store(cv, j->E_open + (lo + i) * w, &E[i]);
  for (int p = 0; p < 2; p++) { free(dm[p]); free(Em[p]); free(bs[p]); free(bm[p]); }
  free(d); free(E);
  return NULL;
}

/* x, a, b, c: AoS ScalarShares; P, out: AoS PointShares; d_open (n scalars) / E_open (n points) optional */
int orc_two_party_point_mul(int cv, size_t n, int threads, const uint64_t* key0, const uint64_t* key1,
                            const uint64_t* x0, const uint64_t* P0, const uint64_t* a0, const uint64_t* b0, const uint64_t* c0,
                            const uint64_t* x1, const uint64_t* P1, const uint64_t* a1, const uint64_t* b1, const uint64_t* c1,
                            uint64_t* out0, uint64_t* out1, uint64_t* d_open, uint64_t* E_open) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pjob_t jobs[256];
  pthread_t th[256];
  for (int t = 0; t < threads; t++) {
    pjob_t* j = &jobs[t];
    j->cv = cv; j->lo = n * (size_t)t / threads; j->hi = n * (size_t)(t + 1) / threads;
    j->key[0] = key0; j->key[1] = key1; j->x[0] = x0; j->x[1] = x1; j->P[0] = P0; j->P[1] = P1;
    j->a[0] = a0; j->a[1] = a1; j->b[0] = b0; j->b[1] = b1; j->c[0] = c0; j->c[1] = c1;
    j->out[0] = out0; j->out[1] = out1; j->d_open = d_open; j->E_open = E_open;
  }
  if (threads == 1) { pjob_run(&jobs[0]); return 0; }
  for (int t = 0; t < threads; t++) if (pthread_create(&th[t], NULL, pjob_run, &jobs[t])) return -1;
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* ====================================================================================================
 * Batch inversion and FFT on shares: Scalar::batch_inverse (scalar.rs:93-100 -> ark_ff::batch_inversion,
 * zeros stay zero) and ark-poly Radix2EvaluationDomain::fft / ifft as used by ScalarShare::fft_helper (share.rs:162-192):
 * X_j = sum_i x_i w^(ij), w = TWO_ADIC_ROOT_OF_UNITY^(2^(28 - log2n)), natural order; ifft = inverse, scaled by n^-1.
 * BN254 Fr only: ark-bn254 fixes GENERATOR = 5, TWO_ADICITY = 28 (crate not vendored; the root below is the widely published
 * 19103219067921713944291392827692070036145651957329286315305642004821462161904, checked in tests/test_oracle_ntt.py).
 * ==================================================================================================== */
static const fe BN254_FR_ROOT28 = {{0x636e735580d13d9cull, 0xa22bf3742445ffd6ull, 0x56452ac01eb203d8ull, 0x1860ef942963f9e7ull}};

void orc_batch_inverse(int f, size_t n, uint64_t* out, const uint64_t* in) {
  const field_t* F = &FIELDS[f];
  for (size_t i = 0; i < n; i++) {
    const fe* x = (const fe*)in + i;
    if (fe_is_zero(x)) ((fe*)out)[i] = *x; else fe_inv(F, (fe*)out + i, x);
  }
}

int orc_fft(int f, int log2n, int inverse, const uint64_t* in, uint64_t* out) {
  if (f != 0 || log2n < 0 || log2n > 28) return -1;
  const field_t* F = &FIELDS[f];
  const size_t n = (size_t)1 << log2n;
  fe w = BN254_FR_ROOT28;
  for (int i = log2n; i < 28; i++) fe_mul(F, &w, &w, &w);
  if (inverse) fe_inv(F, &w, &w);
  fe* x = (fe*)out;
  for (size_t p = 0; p < n; p++) { /* bit-reversed copy */
    size_t r = 0;
    for (int b = 0; b < log2n; b++) r |= ((p >> b) & 1) << (log2n - 1 - b);
    x[p] = ((const fe*)in)[r];
  }
  for (int s = 1; s <= log2n; s++) {
    const size_t m = (size_t)1 << s, half = m >> 1;
    fe wm = w; /* w^(n/m) */
    for (int i = s; i < log2n; i++) fe_mul(F, &wm, &wm, &wm);
    for (size_t k = 0; k < n; k += m) {
      fe wj = F->r;
      for (size_t j = 0; j < half; j++) {
        fe t, u = x[k + j];
        fe_mul(F, &t, &wj, &x[k + j + half]);
        fe_add(F, &x[k + j], &u, &t);
        fe_sub(F, &x[k + j + half], &u, &t);
        fe_mul(F, &wj, &wj, &wm);
      }
    }
  }
  if (inverse) {
    fe two, nn = F->r, ninv;
    fe_add(F, &two, &F->r, &F->r);
    for (int i = 0; i < log2n; i++) fe_mul(F, &nn, &nn, &two);
    fe_inv(F, &ninv, &nn);
    for (size_t i = 0; i < n; i++) fe_mul(F, &x[i], &x[i], &ninv);
  }
  return 0;
}
