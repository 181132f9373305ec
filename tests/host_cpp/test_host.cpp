// Two-party tests of the C++ host mirror (host/arkmpc_host.hpp) over the C ABI, written the way the reference's own tests
// are (`execute_mock_mpc`, /root/reference/online-phase/src/lib.rs:116-201; cases from algebra/scalar/authenticated_scalar.rs:1131-1715,
// algebra/curve/authenticated_curve.rs:1217-1295, integration/src/authenticated_scalar.rs:49-75): share inputs, run the gates on
// the device, open with the MAC check, compare with the CPU oracle (oracle/ark_oracle.c, test infrastructure) on plaintext values.
// Build (tests/test_host_cpp.py): g++ -std=c++17 -Iinclude -Ihost tests/host_cpp/test_host.cpp -Lark_mpc_b200/lib -larkmpc_b200 -Loracle -lark_oracle -lpthread
#include <cstdio>
#include <cstdlib>

#include "arkmpc_host.hpp"

extern "C" {
void orc_synth(int f, uint64_t seed, uint64_t first_index, size_t n, uint64_t* out);
void orc_scalar_batch_mul(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b);
void orc_scalar_batch_add(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b);
void orc_pt_mul_generator(int cv, size_t n, const uint64_t* scalars, uint64_t* out);
void orc_pt_mul(int cv, size_t n, const uint64_t* scalars, const uint64_t* pts, uint64_t* out);
void orc_pt_normalize(int cv, size_t n, const uint64_t* pts, uint64_t* out_xy);
void orc_batch_inverse(int f, size_t n, uint64_t* out, const uint64_t* in);
int orc_fft(int f, int log2n, int inverse, const uint64_t* in, uint64_t* out);
void orc_pt_add(int cv, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out, int sub);
}

using namespace arkmpc;
static int failures = 0;
#define EXPECT(cond, msg)                                   \
  do {                                                      \
    if (!(cond)) { printf("FAIL: %s (%s:%d)\n", msg, __FILE__, __LINE__); failures++; } \
  } while (0)

static std::function<std::unique_ptr<PreprocessingPhase>(int)> party_id_source(const CurveInfo& cv) {
  return [cv](int p) { return std::unique_ptr<PreprocessingPhase>(new PartyIDBeaverSource(p, cv)); };
}

static void test_scalars(const CurveInfo& cv, const char* name) {
  const size_t n = 100;
  HostScalars a, b;
  a.limbs.resize(n * 4); b.limbs.resize(n * 4);
  orc_synth(cv.field, 1, 0, n, a.limbs.data());
  orc_synth(cv.field, 2, 0, n, b.limbs.data());
  std::vector<uint64_t> want_mul(n * 4), want_add(n * 4);
  orc_scalar_batch_mul(cv.field, n, want_mul.data(), a.limbs.data(), b.limbs.data());
  orc_scalar_batch_add(cv.field, n, want_add.data(), a.limbs.data(), b.limbs.data());
  using Out = std::vector<std::vector<uint64_t>>;
  auto res = execute_mock_mpc<Out>(cv, party_id_source(cv), [&](MpcFabric& f) {
    ScalarResult va = f.allocate_scalars(a), vb = f.allocate_scalars(b);
    auto A = f.batch_share_scalar(f.party_id() == 0 ? &va : nullptr, n, 0);
    auto B = f.batch_share_scalar(f.party_id() == 1 ? &vb : nullptr, n, 1);
    using S = AuthenticatedScalarResult;
    Out o;
    o.push_back(S::open_authenticated_batch(S::batch_mul(A, B)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_add(A, B)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_add_public(A, vb)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_mul_public(A, vb)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_sub(S::batch_add(A, B), B)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_mul(A, B).sum()).result().to_host());
    auto e = S::batch_mul(f.allocate_scalar_shares(HostShares{}), f.allocate_scalar_shares(HostShares{}));  // empty batch (:854-856)
    o.push_back(std::vector<uint64_t>{(uint64_t)e.len()});
    return o;
  });
  for (const Out* o : {&res.first, &res.second}) {
    EXPECT((*o)[0] == want_mul, "batch_mul opens to a*b");
    EXPECT((*o)[1] == want_add, "batch_add opens to a+b");
    EXPECT((*o)[2] == want_add, "batch_add_public opens to a+b");
    EXPECT((*o)[3] == want_mul, "batch_mul_public opens to a*b");
    EXPECT((*o)[4] == a.limbs, "a+b-b opens to a");
    std::vector<uint64_t> acc(4, 0), t(4);
    for (size_t i = 0; i < n; i++) { orc_scalar_batch_add(cv.field, 1, t.data(), acc.data(), want_mul.data() + 4 * i); acc = t; }
    EXPECT((*o)[5] == acc, "inner product (batch_mul + sum) opens to sum a_i*b_i");
    EXPECT((*o)[6][0] == 0, "empty batch_mul returns an empty batch");
  }
  printf("%s: scalar gates done\n", name);
}

static void test_party_id_kat(const CurveInfo& cv) {  // offline_prep.rs:137-158: the mock triple opens to 2 * 3 = 6 under key 1
  using Out = std::vector<std::vector<uint64_t>>;
  auto res = execute_mock_mpc<Out>(cv, party_id_source(cv), [&](MpcFabric& f) {
    auto [a, b, c] = f.next_triple_batch(3);
    using S = AuthenticatedScalarResult;
    return Out{S::open_authenticated_batch(a).result().to_host(), S::open_authenticated_batch(b).result().to_host(),
               S::open_authenticated_batch(c).result().to_host()};
  });
  const uint64_t want[3] = {2, 3, 6};
  for (int k = 0; k < 3; k++) {
    Limbs m = mont_small(cv, want[k]);
    for (int i = 0; i < 3; i++) EXPECT(std::equal(m.begin(), m.end(), res.first[k].begin() + 4 * i) && res.first[k] == res.second[k], "PartyIDBeaverSource KAT");
  }
}

static void test_corruption(const CurveInfo& cv) {  // integration/src/authenticated_scalar.rs:49-75
  const size_t n = 16;
  HostScalars a;
  a.limbs.resize(n * 4);
  orc_synth(cv.field, 5, 0, n, a.limbs.data());
  auto res = execute_mock_mpc<int>(cv, party_id_source(cv), [&](MpcFabric& f) {
    ScalarResult va = f.allocate_scalars(a);
    auto A = f.batch_share_scalar(f.party_id() == 0 ? &va : nullptr, n, 0);
    if (f.party_id() == 0) {  // corrupt one MAC share: mac[3] += mac[3] + 1 style tweak through the public-add gate on a 1-element view
      Limbs one = mont_small(cv, 1);
      f.ctx()->check(arkmpc_memcpy_h2d(f.raw(), A.mac->at(3 * 32), one.data(), 32), "corrupt");
      f.ctx()->sync();
    }
    try {
      AuthenticatedScalarResult::open_authenticated_batch(A).result();
      return 0;
    } catch (const AuthenticationError&) {
      return 1;
    }
  });
  EXPECT(res.first == 1 && res.second == 1, "a corrupted MAC makes open_authenticated fail on both parties");
}

// the "next" rows: public batch inversion, FFT on shares (BN254 only), msm_authenticated with public points
static void test_next_rows(const CurveInfo& cv, const char* name) {
  const size_t n = 64, w = cv.point_words;
  HostScalars a, t;
  a.limbs.resize(n * 4); t.limbs.resize(n * 4);
  orc_synth(cv.field, 31, 0, n, a.limbs.data());
  orc_synth(cv.field, 32, 0, n, t.limbs.data());
  std::vector<uint64_t> want_inv(n * 4), want_fft(n * 4), P(n * w), prod(n * w), acc(w), tmp(w), want_msm(8);
  orc_batch_inverse(cv.field, n, want_inv.data(), a.limbs.data());
  const bool has_fft = orc_fft(cv.field, 6, 0, a.limbs.data(), want_fft.data()) == 0;
  orc_pt_mul_generator(cv.curve, n, t.limbs.data(), P.data());
  orc_pt_mul(cv.curve, n, a.limbs.data(), P.data(), prod.data());
  memcpy(acc.data(), prod.data(), w * 8);
  for (size_t i = 1; i < n; i++) { orc_pt_add(cv.curve, 1, acc.data(), prod.data() + i * w, tmp.data(), 0); acc = tmp; }
  orc_pt_normalize(cv.curve, 1, acc.data(), want_msm.data());
  using Out = std::vector<std::vector<uint64_t>>;
  auto res = execute_mock_mpc<Out>(cv, party_id_source(cv), [&](MpcFabric& f) {
    ScalarResult va = f.allocate_scalars(a);
    CurvePointResult pts{&f, f.upload(P.data(), n * w * 8), n};
    auto A = f.batch_share_scalar(f.party_id() == 0 ? &va : nullptr, n, 0);
    using S = AuthenticatedScalarResult;
    Out o;
    o.push_back(ScalarResult::batch_inverse(va).to_host());
    if (has_fft) o.push_back(S::open_authenticated_batch(S::fft(A)).result().to_host());
    else o.push_back({});
    auto m = AuthenticatedPointResult::open_authenticated_batch(CurvePointResult::msm_authenticated(A, pts));
    o.push_back(m.result().to_affine_host());
    o.push_back(CurvePointResult::msm(va, pts).to_affine_host());
    o.push_back(S::open_authenticated_batch(S::batch_inverse(A)).result().to_host());
    o.push_back(S::open_authenticated_batch(S::batch_mul(S::batch_div(A, A), A)).result().to_host());  // (a / a) * a = a
    return o;
  });
  for (const Out* o : {&res.first, &res.second}) {
    EXPECT((*o)[0] == want_inv, "public batch_inverse");
    if (has_fft) EXPECT((*o)[1] == want_fft, "fft on shares opens to the transform of the plaintext");
    EXPECT((*o)[2] == want_msm, "msm_authenticated opens to sum a_i * P_i");
    EXPECT((*o)[3] == want_msm, "public msm");
    EXPECT((*o)[4] == want_inv, "authenticated batch_inverse opens to the inverses");
    EXPECT((*o)[5] == a.limbs, "batch_div: (a / a) * a opens to a");
  }
  printf("%s: inverse / fft / msm done\n", name);
}

static void test_points(const CurveInfo& cv, const char* name) {
  const size_t n = 24, w = cv.point_words;
  HostScalars x, s;
  x.limbs.resize(n * 4); s.limbs.resize(n * 4);
  orc_synth(cv.field, 11, 0, n, x.limbs.data());
  orc_synth(cv.field, 12, 0, n, s.limbs.data());
  std::vector<uint64_t> P(n * w), xP(n * w), want(n * 8), wantP(n * 8);
  orc_pt_mul_generator(cv.curve, n, s.limbs.data(), P.data());
  orc_pt_mul(cv.curve, n, x.limbs.data(), P.data(), xP.data());
  orc_pt_normalize(cv.curve, n, xP.data(), want.data());
  orc_pt_normalize(cv.curve, n, P.data(), wantP.data());
  using Out = std::vector<std::vector<uint64_t>>;
  auto res = execute_mock_mpc<Out>(cv, party_id_source(cv), [&](MpcFabric& f) {
    ScalarResult vx = f.allocate_scalars(x);
    CurvePointResult pts{&f, f.upload(P.data(), n * w * 8), n};
    auto X = f.batch_share_scalar(f.party_id() == 0 ? &vx : nullptr, n, 0);
    auto Pt = f.batch_share_point(f.party_id() == 1 ? &pts : nullptr, n, 1);
    using A = AuthenticatedPointResult;
    Out o;
    o.push_back(A::open_authenticated_batch(A::batch_mul(X, Pt)).result().to_affine_host());
    o.push_back(A::open_authenticated_batch(A::batch_sub(A::batch_add(Pt, Pt), Pt)).result().to_affine_host());
    o.push_back(A::open_authenticated_batch(A::batch_mul_public(vx, Pt)).result().to_affine_host());
    return o;
  });
  for (const Out* o : {&res.first, &res.second}) {
    EXPECT((*o)[0] == want, "[x]*[P] opens to x*P");
    EXPECT((*o)[1] == wantP, "P+P-P opens to P");
    EXPECT((*o)[2] == want, "x*[P] (public scalar) opens to x*P");
  }
  printf("%s: point gates done\n", name);
}

// `test_host --sha3 <len> <seed>`: digest of a deterministic message (byte i = (seed + 131 * i) >> 3), fed in three uneven pieces,
// so that the commitment hash can be compared with hashlib on a box without a GPU
static int sha3_mode(size_t len, unsigned seed) {
  std::vector<uint8_t> msg(len);
  for (size_t i = 0; i < len; i++) msg[i] = (uint8_t)((seed + 131u * i) >> 3);
  Sha3_256 h;
  const size_t c1 = len / 3, c2 = len / 3 + (len > 7 ? 7 : 0);
  h.update(msg.data(), c1);
  h.update(msg.data() + c1, c2 - c1 > len - c1 ? len - c1 : c2 - c1);
  const size_t done = c1 + (c2 - c1 > len - c1 ? len - c1 : c2 - c1);
  h.update(msg.data() + done, len - done);
  for (uint8_t b : h.finalize()) printf("%02x", b);
  printf("\n");
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "--sha3") return sha3_mode((size_t)atol(argv[2]), (unsigned)atol(argv[3]));
  int count = 0;
  if (arkmpc_device_count(&count) != ARKMPC_OK || count == 0) { printf("no CUDA device: the host mirror has no CPU fallback\n"); return 2; }
  for (auto& kv : {std::make_pair(bn254(), "bn254"), std::make_pair(curve25519(), "curve25519")}) {
    test_scalars(kv.first, kv.second);
    test_party_id_kat(kv.first);
    test_corruption(kv.first);
    test_points(kv.first, kv.second);
    test_next_rows(kv.first, kv.second);
  }
  try {  // length mismatch is a programming error, as the reference's assert (authenticated_scalar.rs:852)
    auto cv = bn254();
    execute_mock_mpc<int>(cv, party_id_source(cv), [&](MpcFabric& f) {
      auto t3 = f.next_triple_batch(3);
      auto t4 = f.next_triple_batch(4);
      AuthenticatedScalarResult::batch_add(std::get<0>(t3), std::get<0>(t4));
      return 0;
    });
    EXPECT(false, "length mismatch must throw");
  } catch (const std::invalid_argument&) {
  }
  printf(failures ? "HOST MIRROR TESTS FAILED: %d\n" : "host mirror tests OK (%d failures)\n", failures);
  return failures ? 1 : 0;
}
