// arkmpc_host.hpp — C++ host side above the C ABI (include/arkmpc_b200.h), mirroring the reference's operator surface
// for the hot path.  Header-only, C++17, depends on nothing but the C ABI: no CUDA headers, no torch.
//
// The reference's host is Rust (no Rust toolchain in this image), so this is the compiled-language mirror of what
// INTEGRATION.md binds: same names, argument meaning and error behaviour as (paths under
// /root/reference/online-phase/src):
//   MpcFabric                   fabric.rs:402-978     party_id / mac_key / next_triple_batch / batch_share_scalar /
//                                                      batch_share_point; party 0 sends first (:751-765)
//   PreprocessingPhase          offline_prep.rs:12-82 ; PartyIDBeaverSource :88-170
//   MockNetwork                 network/mock.rs:63-143 (payloads move by reference)
//   ScalarResult                algebra/scalar/scalar_result.rs:170-278
//   AuthenticatedScalarResult   algebra/scalar/authenticated_scalar.rs:129-948
//   AuthenticatedPointResult    algebra/curve/authenticated_curve.rs:66-806
//   MpcError::AuthenticationError  error.rs:9-18 (thrown by open results, authenticated_scalar.rs:368-385)
//   HashCommitment              commitment.rs:63-89 (SHA3-256 over BE bytes || blinder, reduced mod p) — host side
// One handle denotes a WHOLE BATCH in device memory (SURVEY §8b "result carrier"); the reference panics on length
// mismatch (authenticated_scalar.rs:852) -> std::invalid_argument here; empty batches return empty results (:854-856).
// `ark_mpc_b200/fabric.py` is the same mirror for the Python test-suite; tests/host_cpp/test_host.cpp exercises this one.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "arkmpc_b200.h"

namespace arkmpc {

constexpr int PARTY0 = 0, PARTY1 = 1;
using Limbs = std::array<uint64_t, 4>;  // one field element, Montgomery image (Scalar<C>, scalar.rs:46)

struct MpcError : std::runtime_error { using std::runtime_error::runtime_error; };
struct AuthenticationError : MpcError { AuthenticationError() : MpcError("MAC check failed") {} };

// ---------------------------------------------------------------------------------------------------------------
// SHA3-256 (FIPS 202) for the hash commitment; the reference uses the `sha3` crate (commitment.rs:36-41).
// ---------------------------------------------------------------------------------------------------------------
namespace detail {
inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
// Keccak-f[1600], the 25 lanes in locals and each round written out (theta, rho + pi, chi, iota): ~2.5 ns per absorbed byte; the
// commitment hashes n x 32 bytes per open_authenticated_batch and was 80 % of a 1024-gate iteration with a table-driven round.
inline void keccak_f(uint64_t st[25]) {
  static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                                  0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                                  0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                                  0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                                  0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  uint64_t a00 = st[0], a01 = st[1], a02 = st[2], a03 = st[3], a04 = st[4], a05 = st[5], a06 = st[6], a07 = st[7], a08 = st[8], a09 = st[9],
           a10 = st[10], a11 = st[11], a12 = st[12], a13 = st[13], a14 = st[14], a15 = st[15], a16 = st[16], a17 = st[17], a18 = st[18],
           a19 = st[19], a20 = st[20], a21 = st[21], a22 = st[22], a23 = st[23], a24 = st[24];
  for (int r = 0; r < 24; r++) {
    // theta
    const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                   c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
    const uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
    // rho + pi: b[y][2x+3y] = rot(a[x][y])
    const uint64_t b00 = a00 ^ d0, b10 = rotl64(a01 ^ d1, 1), b20 = rotl64(a02 ^ d2, 62), b05 = rotl64(a03 ^ d3, 28), b15 = rotl64(a04 ^ d4, 27),
                   b16 = rotl64(a05 ^ d0, 36), b01 = rotl64(a06 ^ d1, 44), b11 = rotl64(a07 ^ d2, 6), b21 = rotl64(a08 ^ d3, 55), b06 = rotl64(a09 ^ d4, 20),
                   b07 = rotl64(a10 ^ d0, 3), b17 = rotl64(a11 ^ d1, 10), b02 = rotl64(a12 ^ d2, 43), b12 = rotl64(a13 ^ d3, 25), b22 = rotl64(a14 ^ d4, 39),
                   b23 = rotl64(a15 ^ d0, 41), b08 = rotl64(a16 ^ d1, 45), b18 = rotl64(a17 ^ d2, 15), b03 = rotl64(a18 ^ d3, 21), b13 = rotl64(a19 ^ d4, 8),
                   b14 = rotl64(a20 ^ d0, 18), b24 = rotl64(a21 ^ d1, 2), b09 = rotl64(a22 ^ d2, 61), b19 = rotl64(a23 ^ d3, 56), b04 = rotl64(a24 ^ d4, 14);
    // chi (+ iota on lane 0)
    a00 = b00 ^ (~b01 & b02) ^ RC[r]; a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
    a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
    a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
    a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
    a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
  }
  st[0] = a00; st[1] = a01; st[2] = a02; st[3] = a03; st[4] = a04; st[5] = a05; st[6] = a06; st[7] = a07; st[8] = a08; st[9] = a09;
  st[10] = a10; st[11] = a11; st[12] = a12; st[13] = a13; st[14] = a14; st[15] = a15; st[16] = a16; st[17] = a17; st[18] = a18; st[19] = a19;
  st[20] = a20; st[21] = a21; st[22] = a22; st[23] = a23; st[24] = a24;
}
}  // namespace detail

class Sha3_256 {
 public:
  void update(const uint8_t* data, size_t len) {
    size_t i = 0;
    for (; i < len && pos_ != 0; i++) absorb_byte(data[i]);  // finish a partial block
    for (; i + kRate <= len; i += kRate) {                   // whole blocks, a lane at a time (little-endian host)
      for (size_t j = 0; j < kRate / 8; j++) {
        uint64_t w;
        memcpy(&w, data + i + 8 * j, 8);
        st_[j] ^= w;
      }
      detail::keccak_f(st_);
    }
    for (; i < len; i++) absorb_byte(data[i]);
  }
  std::array<uint8_t, 32> finalize() {
    reinterpret_cast<uint8_t*>(st_)[pos_] ^= 0x06;
    reinterpret_cast<uint8_t*>(st_)[kRate - 1] ^= 0x80;
    detail::keccak_f(st_);
    std::array<uint8_t, 32> out;
    memcpy(out.data(), st_, 32);
    return out;
  }

 private:
  static constexpr size_t kRate = 136;
  void absorb_byte(uint8_t v) {
    reinterpret_cast<uint8_t*>(st_)[pos_++] ^= v;
    if (pos_ == kRate) { detail::keccak_f(st_); pos_ = 0; }
  }
  uint64_t st_[25] = {0};
  size_t pos_ = 0;
};

// ---------------------------------------------------------------------------------------------------------------
// Host-side profile (ARKMPC_HOST_PROFILE=1): wall time spent per category, summed over both parties' threads.  At the reference
// bench's n = 1024 the arithmetic is microseconds; this says where the milliseconds go.
// ---------------------------------------------------------------------------------------------------------------
namespace prof {
enum Cat { kMalloc, kFree, kSync, kH2D, kD2H, kWait, kHash, kSource, kCats };
inline const char* name(int c) { static const char* n[] = {"malloc", "free", "sync", "h2d", "d2h", "peer_wait", "sha3", "preprocessing"}; return n[c]; }
inline bool enabled() { static const bool on = [] { const char* e = std::getenv("ARKMPC_HOST_PROFILE"); return e && *e && *e != '0'; }(); return on; }
inline std::atomic<uint64_t>& ns(int c) { static std::atomic<uint64_t> v[kCats]; return v[c]; }
inline std::atomic<uint64_t>& calls(int c) { static std::atomic<uint64_t> v[kCats]; return v[c]; }
struct Scope {
  int c;
  std::chrono::steady_clock::time_point t0;
  explicit Scope(int cat) : c(cat) { if (enabled()) t0 = std::chrono::steady_clock::now(); }
  ~Scope() {
    if (!enabled()) return;
    ns(c) += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    calls(c)++;
  }
};
inline void reset() { for (int c = 0; c < kCats; c++) { ns(c) = 0; calls(c) = 0; } }
inline void report(FILE* f, double per) {  // `per`: divide by (e.g. the number of iterations)
  for (int c = 0; c < kCats; c++) fprintf(f, "  %-14s %9.1f us  %6.1f calls\n", name(c), ns(c) / 1e3 / per, calls(c) / per);
}
}  // namespace prof

// ---------------------------------------------------------------------------------------------------------------
// Context and device memory
// ---------------------------------------------------------------------------------------------------------------
class Context {
 public:
  explicit Context(int device = 0) {
    int rc = arkmpc_ctx_create(device, &raw_);
    if (rc != ARKMPC_OK) throw MpcError(std::string("arkmpc_ctx_create: ") + arkmpc_status_string(rc) + " (the gate engine has no CPU fallback)");
  }
  ~Context() { if (raw_) arkmpc_ctx_destroy(raw_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  arkmpc_ctx* raw() const { return raw_; }
  // the reference panics on misuse (authenticated_scalar.rs:852, result.rs:127-133); a failed call is an exception here
  void check(int rc, const char* what) const {
    if (rc != ARKMPC_OK) throw MpcError(std::string(what) + ": " + arkmpc_status_string(rc) + ": " + arkmpc_last_error(raw_));
  }
  void sync() const { prof::Scope ps(prof::kSync); check(arkmpc_ctx_sync(raw_), "arkmpc_ctx_sync"); }

 private:
  arkmpc_ctx* raw_ = nullptr;
};

class DevBuf {
 public:
  DevBuf(std::shared_ptr<Context> ctx, size_t bytes) : ctx_(std::move(ctx)), bytes_(bytes) {
    prof::Scope ps(prof::kMalloc);
    ctx_->check(arkmpc_malloc(ctx_->raw(), bytes, &p_), "arkmpc_malloc");
  }
  ~DevBuf() { prof::Scope ps(prof::kFree); if (p_) arkmpc_free(ctx_->raw(), p_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  uint64_t* u64() const { return static_cast<uint64_t*>(p_); }
  uint64_t* at(size_t byte_off) const { return reinterpret_cast<uint64_t*>(static_cast<char*>(p_) + byte_off); }
  size_t bytes() const { return bytes_; }

 private:
  std::shared_ptr<Context> ctx_;
  void* p_ = nullptr;
  size_t bytes_;
};
using Buf = std::shared_ptr<DevBuf>;

struct CurveInfo { int field; int curve; size_t point_words; Limbs modulus; Limbs r_mod_p; };  // scalar-field modulus and R = 2^256 mod p
inline CurveInfo bn254() {
  return {ARKMPC_BN254_FR, ARKMPC_BN254_G1, 12, {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull},
          {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full}};
}
inline CurveInfo curve25519() {
  return {ARKMPC_CURVE25519_FR, ARKMPC_CURVE25519_EDWARDS, 16, {0x5812631a5cf5d3edull, 0x14def9dea2f79cd6ull, 0ull, 0x1000000000000000ull},
          {0xd6ec31748d98951dull, 0xc6ef5bf4737dcf70ull, 0xfffffffffffffffeull, 0x0fffffffffffffffull}};
}

// Montgomery image of a small integer: v * R mod p by double-and-add on R mod p (p < 2^255, so sums fit 256 bits)
inline Limbs mont_small(const CurveInfo& cv, uint64_t v) {
  auto add_mod = [&](const Limbs& a, const Limbs& b) {
    Limbs s;
    unsigned __int128 c = 0;
    for (int j = 0; j < 4; j++) { c += (unsigned __int128)a[j] + b[j]; s[j] = (uint64_t)c; c >>= 64; }
    bool ge = true;
    for (int j = 3; j >= 0; j--) if (s[j] != cv.modulus[j]) { ge = s[j] > cv.modulus[j]; break; }
    if (ge) {
      unsigned __int128 br = 0;
      for (int j = 0; j < 4; j++) { unsigned __int128 t = (unsigned __int128)s[j] - cv.modulus[j] - (uint64_t)br; s[j] = (uint64_t)t; br = (t >> 64) & 1; }
    }
    return s;
  };
  Limbs acc{0, 0, 0, 0};
  for (int i = 63; i >= 0; i--) {
    acc = add_mod(acc, acc);
    if ((v >> i) & 1) acc = add_mod(acc, cv.r_mod_p);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Network (network/mock.rs): an unbounded in-memory duplex; device buffers travel by reference
// ---------------------------------------------------------------------------------------------------------------
struct Message { Buf buf; std::vector<uint64_t> host; bool poison = false; };

class Channel {
 public:
  void put(Message m) { { std::lock_guard<std::mutex> l(mu_); q_.push(std::move(m)); } cv_.notify_one(); }
  Message get() {
    prof::Scope ps(prof::kWait);
    std::unique_lock<std::mutex> l(mu_);
    cv_.wait(l, [&] { return !q_.empty(); });
    Message m = std::move(q_.front());
    q_.pop();
    return m;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::queue<Message> q_;
};

class MockNetwork {
 public:
  MockNetwork(int party_id, std::shared_ptr<Channel> send, std::shared_ptr<Channel> recv) : party_(party_id), send_(std::move(send)), recv_(std::move(recv)) {}
  static std::pair<MockNetwork, MockNetwork> new_duplex_pair() {
    auto a = std::make_shared<Channel>(), b = std::make_shared<Channel>();
    return {MockNetwork(PARTY0, a, b), MockNetwork(PARTY1, b, a)};
  }
  int party_id() const { return party_; }
  void send_message(Message m) { send_->put(std::move(m)); }
  Message receive_message() {
    Message m = recv_->get();
    if (m.poison) throw MpcError("the counterparty failed");
    return m;
  }
  void poison() { Message m; m.poison = true; send_->put(std::move(m)); }

 private:
  int party_;
  std::shared_ptr<Channel> send_, recv_;
};

// ---------------------------------------------------------------------------------------------------------------
// Preprocessing (offline_prep.rs).  Host values are what the Rust trait returns in memory: AoS ScalarShare images.
// ---------------------------------------------------------------------------------------------------------------
struct HostShares { std::vector<uint64_t> aos; size_t n() const { return aos.size() / 8; } };  // n x {share[4], mac[4]}
struct HostScalars { std::vector<uint64_t> limbs; size_t n() const { return limbs.size() / 4; } };

class PreprocessingPhase {
 public:
  virtual ~PreprocessingPhase() = default;
  virtual Limbs get_mac_key_share() = 0;
  virtual std::tuple<HostShares, HostShares, HostShares> next_triplet_batch(size_t n) = 0;
  virtual std::pair<HostScalars, HostShares> next_local_input_mask_batch(size_t n) = 0;
  virtual HostShares next_counterparty_input_mask_batch(size_t n) = 0;
  virtual HostShares next_shared_value_batch(size_t n) = 0;  // offline_prep.rs: next_shared_value_batch
};

// ---------------------------------------------------------------------------------------------------------------
// Fabric and result handles
// ---------------------------------------------------------------------------------------------------------------
class MpcFabric;

struct ScalarResult {  // a batch of public scalars: one device plane
  MpcFabric* fabric = nullptr;
  Buf values;
  size_t n = 0;
  size_t len() const { return n; }
  static ScalarResult batch_add(const ScalarResult& a, const ScalarResult& b);
  static ScalarResult batch_sub(const ScalarResult& a, const ScalarResult& b);
  static ScalarResult batch_mul(const ScalarResult& a, const ScalarResult& b);  // scalar_result.rs:257-278
  static ScalarResult batch_neg(const ScalarResult& a);
  static ScalarResult batch_inverse(const ScalarResult& a);                     // scalar.rs:93-100 (zeros stay zero)
  std::vector<uint64_t> to_host() const;  // n x 4 Montgomery limbs
};

struct AuthenticatedScalarOpenResult {  // authenticated_scalar.rs:358-385
  ScalarResult value;
  bool mac_check = false;
  const ScalarResult& result() const {
    if (!mac_check) throw AuthenticationError();
    return value;
  }
};

struct AuthenticatedScalarResult {  // a batch of ScalarShares: two device planes
  MpcFabric* fabric = nullptr;
  Buf share, mac;
  size_t n = 0;
  size_t len() const { return n; }
  static AuthenticatedScalarResult batch_add(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b);
  static AuthenticatedScalarResult batch_sub(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b);
  static AuthenticatedScalarResult batch_neg(const AuthenticatedScalarResult& a);
  static AuthenticatedScalarResult batch_add_public(const AuthenticatedScalarResult& a, const ScalarResult& b);
  static AuthenticatedScalarResult batch_sub_public(const AuthenticatedScalarResult& a, const ScalarResult& b);
  static AuthenticatedScalarResult batch_mul_public(const AuthenticatedScalarResult& a, const ScalarResult& b);
  static AuthenticatedScalarResult batch_mul(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b);  // :848-879
  static ScalarResult open_batch(const AuthenticatedScalarResult& v);                                                  // :129-172
  static AuthenticatedScalarOpenResult open_authenticated_batch(const AuthenticatedScalarResult& v);                   // :278-354
  AuthenticatedScalarResult sum() const;                                                                               // :563-576
  static AuthenticatedScalarResult batch_inverse(const AuthenticatedScalarResult& v);                                  // :55-82
  static AuthenticatedScalarResult batch_div(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b);  // :974-977
  // ark-poly fft / ifft on the share and mac planes (:1011-1070); len() must be a power of two (the caller pads, as D::new does)
  static AuthenticatedScalarResult fft(const AuthenticatedScalarResult& x, bool inverse = false);
};

struct AuthenticatedPointResult;
struct CurvePointResult {  // a batch of public points, AoS projective image
  MpcFabric* fabric = nullptr;
  Buf points;
  size_t n = 0;
  size_t len() const { return n; }
  static CurvePointResult msm(const ScalarResult& scalars, const CurvePointResult& points);                                // curve.rs:549-560
  static AuthenticatedPointResult msm_authenticated(const AuthenticatedScalarResult& scalars, const CurvePointResult& points);  // :619-642
  std::vector<uint64_t> to_affine_host() const;  // n x 8: canonical affine (x, y)
};

struct AuthenticatedPointOpenResult {
  CurvePointResult value;
  bool mac_check = false;
  const CurvePointResult& result() const {
    if (!mac_check) throw AuthenticationError();
    return value;
  }
};

struct AuthenticatedPointResult {  // a batch of PointShares {share, mac}
  MpcFabric* fabric = nullptr;
  Buf shares;
  size_t n = 0;
  size_t len() const { return n; }
  static AuthenticatedPointResult batch_add(const AuthenticatedPointResult& a, const AuthenticatedPointResult& b);       // :396-421
  static AuthenticatedPointResult batch_sub(const AuthenticatedPointResult& a, const AuthenticatedPointResult& b);       // :520-545
  static AuthenticatedPointResult batch_neg(const AuthenticatedPointResult& a);                                           // :604-621
  static AuthenticatedPointResult batch_add_public(const AuthenticatedPointResult& a, const CurvePointResult& b);        // :429-465
  static AuthenticatedPointResult batch_mul_public(const ScalarResult& a, const AuthenticatedPointResult& b);            // :718-751
  static AuthenticatedPointResult batch_mul_generator(const AuthenticatedScalarResult& a);                               // :754-780
  static AuthenticatedPointResult batch_mul(const AuthenticatedScalarResult& a, const AuthenticatedPointResult& b);      // :682-714
  static CurvePointResult open_batch(const AuthenticatedPointResult& v);                                                  // :66-109
  static AuthenticatedPointOpenResult open_authenticated_batch(const AuthenticatedPointResult& v);                       // :193-283
};

class MpcFabric {
 public:
  MpcFabric(MockNetwork net, std::unique_ptr<PreprocessingPhase> src, CurveInfo cv, int device = 0)
      : net_(std::move(net)), src_(std::move(src)), cv_(cv), ctx_(std::make_shared<Context>(device)) {
    party_ = net_.party_id();
    key_ = src_->get_mac_key_share();
  }
  int party_id() const { return party_; }
  const Limbs& mac_key() const { return key_; }
  const CurveInfo& curve() const { return cv_; }
  const std::shared_ptr<Context>& ctx() const { return ctx_; }
  arkmpc_ctx* raw() const { return ctx_->raw(); }
  size_t num_gates() const { return gates_; }  // fabric.rs:479-481
  void count_gate(size_t k = 1) { gates_ += k; }

  Buf alloc(size_t bytes) const { return std::make_shared<DevBuf>(ctx_, bytes ? bytes : 32); }
  Buf upload(const uint64_t* host, size_t bytes) const {
    Buf b = alloc(bytes);
    { prof::Scope ps(prof::kH2D); if (bytes) ctx_->check(arkmpc_memcpy_h2d(raw(), b->u64(), host, bytes), "arkmpc_memcpy_h2d"); }
    // no synchronisation: arkmpc_memcpy_h2d has read a pageable source (every host vector here) when it returns
    return b;
  }
  std::vector<uint64_t> download(const Buf& b, size_t bytes) const {
    std::vector<uint64_t> out(bytes / 8);
    { prof::Scope ps(prof::kD2H); if (bytes) ctx_->check(arkmpc_memcpy_d2h(raw(), out.data(), b->u64(), bytes), "arkmpc_memcpy_d2h"); }
    ctx_->sync();
    return out;
  }

  // plain integers (4 LE limbs each, < p) -> public scalars in Montgomery form
  ScalarResult allocate_scalars_plain(const std::vector<uint64_t>& plain_limbs) {
    const size_t n = plain_limbs.size() / 4;
    Buf plain = upload(plain_limbs.data(), n * 32), mont = alloc(n * 32);
    ctx_->check(arkmpc_fr_to_mont(raw(), cv_.field, n, plain->u64(), mont->u64()), "arkmpc_fr_to_mont");
    return ScalarResult{this, mont, n};
  }
  ScalarResult allocate_scalars(const HostScalars& mont) { return ScalarResult{this, upload(mont.limbs.data(), mont.n() * 32), mont.n()}; }
  AuthenticatedScalarResult allocate_scalar_shares(const HostShares& s) {  // fabric.rs:676-686
    const size_t n = s.n();
    Buf aos = upload(s.aos.data(), n * 64), sh = alloc(n * 32), mc = alloc(n * 32);
    ctx_->check(arkmpc_share_unzip(raw(), n, aos->u64(), sh->u64(), mc->u64()), "arkmpc_share_unzip");
    // `aos` is released on return: arkmpc_free is ordered after the work submitted so far (include/arkmpc_b200.h, memory)
    return AuthenticatedScalarResult{this, sh, mc, n};
  }
  std::tuple<AuthenticatedScalarResult, AuthenticatedScalarResult, AuthenticatedScalarResult> next_triple_batch(size_t n) {  // fabric.rs:894-915
    std::tuple<HostShares, HostShares, HostShares> t;
    { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); t = src_->next_triplet_batch(n); }
    return {allocate_scalar_shares(std::get<0>(t)), allocate_scalar_shares(std::get<1>(t)), allocate_scalar_shares(std::get<2>(t))};
  }

  AuthenticatedScalarResult random_shared_scalars(size_t n) {  // fabric.rs:950-965
    HostShares v;
    { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); v = src_->next_shared_value_batch(n); }
    return allocate_scalar_shares(v);
  }

  // -- network: party 0 sends then receives, party 1 receives then sends (fabric.rs:751-765) --
  void send(const Buf& b) { ctx_->sync(); Message m; m.buf = b; net_.send_message(std::move(m)); }
  Buf receive() { return net_.receive_message().buf; }
  Buf exchange(const Buf& mine) {
    if (party_ == PARTY0) { send(mine); return receive(); }
    Buf peer = receive();
    send(mine);
    return peer;
  }
  std::vector<uint64_t> exchange_host(const std::vector<uint64_t>& mine) {
    Message m; m.host = mine;
    if (party_ == PARTY0) { net_.send_message(std::move(m)); return net_.receive_message().host; }
    auto peer = net_.receive_message().host;
    net_.send_message(std::move(m));
    return peer;
  }
  Buf share_plaintext(const Buf& mine, int sender) {  // fabric.rs:786-814
    if (party_ == sender) { send(mine); return mine; }
    return receive();
  }
  void poison_peer() { net_.poison(); }

  // -- input sharing (fabric.rs:578-649).  The receiver passes only the length. --
  AuthenticatedScalarResult batch_share_scalar(const ScalarResult* vals, size_t n, int sender) {
    Buf masked;
    HostShares mask_shares;
    if (party_ == sender) {
      std::pair<HostScalars, HostShares> m;
      { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); m = src_->next_local_input_mask_batch(n); }
      ScalarResult masks = allocate_scalars(m.first);
      masked = share_plaintext(ScalarResult::batch_sub(*vals, masks).values, sender);
      mask_shares = std::move(m.second);
    } else {
      { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); mask_shares = src_->next_counterparty_input_mask_batch(n); }
      masked = share_plaintext(nullptr, sender);
    }
    return AuthenticatedScalarResult::batch_add_public(allocate_scalar_shares(mask_shares), ScalarResult{this, masked, n});
  }
  AuthenticatedPointResult batch_share_point(const CurvePointResult* pts, size_t n, int sender) {
    const size_t pb = cv_.point_words * 8;
    Buf masked;
    HostShares mask_shares;
    if (party_ == sender) {
      std::pair<HostScalars, HostShares> m;
      { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); m = src_->next_local_input_mask_batch(n); }
      ScalarResult masks = allocate_scalars(m.first);
      Buf mg = alloc(n * pb), diff = alloc(n * pb);
      ctx_->check(arkmpc_pt_mul_generator_public(raw(), cv_.curve, n, masks.values->u64(), mg->u64()), "arkmpc_pt_mul_generator_public");
      ctx_->check(arkmpc_pt_sub(raw(), cv_.curve, n, pts->points->u64(), mg->u64(), diff->u64()), "arkmpc_pt_sub");
      masked = share_plaintext(diff, sender);
      mask_shares = std::move(m.second);
    } else {
      { prof::Scope ps(prof::kSource); std::lock_guard<std::mutex> l(src_mu_); mask_shares = src_->next_counterparty_input_mask_batch(n); }
      masked = share_plaintext(nullptr, sender);
    }
    AuthenticatedPointResult masks_g = AuthenticatedPointResult::batch_mul_generator(allocate_scalar_shares(mask_shares));
    return AuthenticatedPointResult::batch_add_public(masks_g, CurvePointResult{this, masked, n});
  }

  // commitment.rs:63-89
  Limbs commit(const std::vector<uint8_t>& value_bytes, const Limbs& blinder_plain) const {
    prof::Scope ps(prof::kHash);
    Sha3_256 h;
    h.update(value_bytes.data(), value_bytes.size());
    uint8_t be[32];
    for (int i = 0; i < 32; i++) be[i] = (uint8_t)(blinder_plain[3 - i / 8] >> (8 * (7 - i % 8)));
    h.update(be, 32);
    auto d = h.finalize();
    return reduce_be(d.data());
  }
  // 32 big-endian bytes mod p (`from_be_bytes_mod_order`): 2^256 < 16 p for both fields, a few conditional subtractions
  Limbs reduce_be(const uint8_t* be) const {
    Limbs v;
    for (int j = 0; j < 4; j++) {
      uint64_t w = 0;
      for (int k = 0; k < 8; k++) w = (w << 8) | be[(3 - j) * 8 + k];
      v[j] = w;
    }
    auto geq = [&](const Limbs& a) { for (int j = 3; j >= 0; j--) { if (a[j] != cv_.modulus[j]) return a[j] > cv_.modulus[j]; } return true; };
    while (geq(v)) {
      unsigned __int128 br = 0;
      for (int j = 0; j < 4; j++) {
        unsigned __int128 t = (unsigned __int128)v[j] - cv_.modulus[j] - (uint64_t)br;
        v[j] = (uint64_t)t;
        br = (t >> 64) & 1;
      }
    }
    return v;
  }
  Limbs random_blinder() {
    std::random_device rd;
    Limbs b;
    for (auto& w : b) w = ((uint64_t)rd() << 32) | rd();
    b[3] &= 0x0fffffffffffffffull;  // < 2^252 <= p for both fields
    return b;
  }

 private:
  MockNetwork net_;
  std::unique_ptr<PreprocessingPhase> src_;
  std::mutex src_mu_;  // fabric.rs:208 Arc<Mutex<Box<dyn PreprocessingPhase>>>
  CurveInfo cv_;
  std::shared_ptr<Context> ctx_;
  int party_ = 0;
  Limbs key_{};
  size_t gates_ = 0;
};

// PartyIDBeaverSource (offline_prep.rs:88-170): a = 2, b = 3, c = 6, [a] = (1,1), [b] = (3,0), [c] = (2,4); key share = party id;
// input masks are 3.
class PartyIDBeaverSource : public PreprocessingPhase {
 public:
  PartyIDBeaverSource(int party_id, const CurveInfo& cv) : party_((uint64_t)party_id), mont_([cv](uint64_t v) { return mont_small(cv, v); }) {}
  Limbs get_mac_key_share() override { return mont_(party_); }
  std::tuple<HostShares, HostShares, HostShares> next_triplet_batch(size_t n) override {
    const uint64_t a = 1, b = party_ == 0 ? 3 : 0, c = party_ == 0 ? 2 : 4;
    return {fill(a, party_ * 2, n), fill(b, party_ * 3, n), fill(c, party_ * 6, n)};
  }
  std::pair<HostScalars, HostShares> next_local_input_mask_batch(size_t n) override {
    HostScalars m;
    Limbs three = mont_(3);
    m.limbs.resize(n * 4);
    for (size_t i = 0; i < n; i++) memcpy(m.limbs.data() + 4 * i, three.data(), 32);
    return {m, fill(party_ * 3, party_ * 3, n)};
  }
  HostShares next_counterparty_input_mask_batch(size_t n) override { return fill(3 * party_, party_ * 3 * party_, n); }
  HostShares next_shared_value_batch(size_t n) override { return fill(party_, party_, n); }  // :166-168: a sharing of 1 under key 1

 private:
  HostShares fill(uint64_t share, uint64_t mac, size_t n) {
    Limbs s = mont_(share), m = mont_(mac);
    HostShares out;
    out.aos.resize(n * 8);
    if (n == 0) return out;
    uint64_t* w = out.aos.data();
    memcpy(w, s.data(), 32);
    memcpy(w + 4, m.data(), 32);
    for (size_t have = 1; have < n; have *= 2) memcpy(w + 8 * have, w, 64 * std::min(have, n - have));  // doubling fill
    return out;
  }
  uint64_t party_;
  std::function<Limbs(uint64_t)> mont_;
};

// ---------------------------------------------------------------------------------------------------------------
// Gate implementations
// ---------------------------------------------------------------------------------------------------------------
namespace detail {
inline void same_len(size_t a, size_t b, const char* what) {
  if (a != b) throw std::invalid_argument(std::string(what) + " requires equal length inputs");
}
}  // namespace detail

#define ARKMPC_F (a.fabric)
inline ScalarResult ScalarResult::batch_add(const ScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_add");
  Buf o = ARKMPC_F->alloc(a.n * 32);
  ARKMPC_F->ctx()->check(arkmpc_fr_add(ARKMPC_F->raw(), ARKMPC_F->curve().field, a.n, a.values->u64(), b.values->u64(), o->u64()), "arkmpc_fr_add");
  ARKMPC_F->count_gate();
  return {a.fabric, o, a.n};
}
inline ScalarResult ScalarResult::batch_sub(const ScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_sub");
  Buf o = ARKMPC_F->alloc(a.n * 32);
  ARKMPC_F->ctx()->check(arkmpc_fr_sub(ARKMPC_F->raw(), ARKMPC_F->curve().field, a.n, a.values->u64(), b.values->u64(), o->u64()), "arkmpc_fr_sub");
  ARKMPC_F->count_gate();
  return {a.fabric, o, a.n};
}
inline ScalarResult ScalarResult::batch_mul(const ScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_mul");
  Buf o = ARKMPC_F->alloc(a.n * 32);
  ARKMPC_F->ctx()->check(arkmpc_fr_mul(ARKMPC_F->raw(), ARKMPC_F->curve().field, a.n, a.values->u64(), b.values->u64(), o->u64()), "arkmpc_fr_mul");
  ARKMPC_F->count_gate();
  return {a.fabric, o, a.n};
}
inline ScalarResult ScalarResult::batch_neg(const ScalarResult& a) {
  Buf o = ARKMPC_F->alloc(a.n * 32);
  ARKMPC_F->ctx()->check(arkmpc_fr_neg(ARKMPC_F->raw(), ARKMPC_F->curve().field, a.n, a.values->u64(), o->u64()), "arkmpc_fr_neg");
  ARKMPC_F->count_gate();
  return {a.fabric, o, a.n};
}
inline ScalarResult ScalarResult::batch_inverse(const ScalarResult& a) {
  Buf o = ARKMPC_F->alloc(a.n * 32);
  ARKMPC_F->ctx()->check(arkmpc_fr_batch_inverse(ARKMPC_F->raw(), ARKMPC_F->curve().field, a.n, a.values->u64(), o->u64()), "arkmpc_fr_batch_inverse");
  ARKMPC_F->count_gate();
  return {a.fabric, o, a.n};
}
inline std::vector<uint64_t> ScalarResult::to_host() const { return fabric->download(values, n * 32); }

namespace detail {
using ShareBin = int (*)(arkmpc_ctx*, int, size_t, const uint64_t*, const uint64_t*, const uint64_t*, const uint64_t*, uint64_t*, uint64_t*);
inline AuthenticatedScalarResult share_binary(ShareBin fn, const char* what, const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b) {
  same_len(a.n, b.n, what);
  MpcFabric* f = a.fabric;
  Buf s = f->alloc(a.n * 32), m = f->alloc(a.n * 32);
  f->ctx()->check(fn(f->raw(), f->curve().field, a.n, a.share->u64(), a.mac->u64(), b.share->u64(), b.mac->u64(), s->u64(), m->u64()), what);
  f->count_gate();
  return {f, s, m, a.n};
}
}  // namespace detail

inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_add(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b) {
  return detail::share_binary(arkmpc_fr_share_add, "batch_add", a, b);
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_sub(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b) {
  return detail::share_binary(arkmpc_fr_share_sub, "batch_sub", a, b);
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_neg(const AuthenticatedScalarResult& a) {
  MpcFabric* f = a.fabric;
  Buf s = f->alloc(a.n * 32), m = f->alloc(a.n * 32);
  f->ctx()->check(arkmpc_fr_share_neg(f->raw(), f->curve().field, a.n, a.share->u64(), a.mac->u64(), s->u64(), m->u64()), "batch_neg");
  f->count_gate();
  return {f, s, m, a.n};
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_add_public(const AuthenticatedScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_add_public");
  MpcFabric* f = a.fabric;
  Buf s = f->alloc(a.n * 32), m = f->alloc(a.n * 32);
  f->ctx()->check(arkmpc_fr_share_add_public(f->raw(), f->curve().field, f->party_id(), f->mac_key().data(), a.n, a.share->u64(), a.mac->u64(),
                                             b.values->u64(), s->u64(), m->u64()), "batch_add_public");
  f->count_gate();
  return {f, s, m, a.n};
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_sub_public(const AuthenticatedScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_sub_public");
  MpcFabric* f = a.fabric;
  Buf s = f->alloc(a.n * 32), m = f->alloc(a.n * 32);
  f->ctx()->check(arkmpc_fr_share_sub_public(f->raw(), f->curve().field, f->party_id(), f->mac_key().data(), a.n, a.share->u64(), a.mac->u64(),
                                             b.values->u64(), s->u64(), m->u64()), "batch_sub_public");
  f->count_gate();
  return {f, s, m, a.n};
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_mul_public(const AuthenticatedScalarResult& a, const ScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_mul_public");
  MpcFabric* f = a.fabric;
  Buf s = f->alloc(a.n * 32), m = f->alloc(a.n * 32);
  f->ctx()->check(arkmpc_fr_share_mul_public(f->raw(), f->curve().field, a.n, a.share->u64(), a.mac->u64(), b.values->u64(), s->u64(), m->u64()),
                  "batch_mul_public");
  f->count_gate();
  return {f, s, m, a.n};
}

inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_mul(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b) {
  detail::same_len(a.n, b.n, "batch_mul");
  MpcFabric* f = a.fabric;
  const size_t n = a.n;
  if (n == 0) return {f, f->alloc(0), f->alloc(0), 0};  // :854-856
  auto [ba, bb, bc] = f->next_triple_batch(n);
  const int fid = f->curve().field;
  Buf de_mine = f->alloc(2 * n * 32);  // d || e, like `all_masks` (:866)
  f->ctx()->check(arkmpc_fr_beaver_mask(f->raw(), fid, n, a.share->u64(), b.share->u64(), ba.share->u64(), bb.share->u64(), de_mine->u64(),
                                        de_mine->at(n * 32)), "arkmpc_fr_beaver_mask");
  Buf de_peer = f->exchange(de_mine);  // the network half of open_batch (:129-160)
  Buf s = f->alloc(n * 32), m = f->alloc(n * 32);
  f->ctx()->check(arkmpc_fr_beaver_recombine(f->raw(), fid, f->party_id(), f->mac_key().data(), n, de_mine->u64(), de_mine->at(n * 32), de_peer->u64(),
                                             de_peer->at(n * 32), ba.share->u64(), ba.mac->u64(), bb.share->u64(), bb.mac->u64(), bc.share->u64(),
                                             bc.mac->u64(), s->u64(), m->u64(), nullptr, nullptr), "arkmpc_fr_beaver_recombine");
  // the triple planes and de buffers are released on return; arkmpc_free orders their reuse after this kernel
  f->count_gate(2);
  return {f, s, m, n};
}

inline ScalarResult AuthenticatedScalarResult::open_batch(const AuthenticatedScalarResult& v) {
  MpcFabric* f = v.fabric;
  if (v.n == 0) return {f, f->alloc(0), 0};
  Buf peer = f->exchange(v.share);
  Buf o = f->alloc(v.n * 32);
  f->ctx()->check(arkmpc_fr_add(f->raw(), f->curve().field, v.n, v.share->u64(), peer->u64(), o->u64()), "open_batch");
  f->count_gate();
  return {f, o, v.n};
}

// Two rounds (Bar-Ilan & Beaver): mask with shared randomness, open with the MAC check, invert in public, unmask.
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_inverse(const AuthenticatedScalarResult& v) {
  if (v.n == 0) throw std::invalid_argument("cannot invert empty batch of scalars");
  AuthenticatedScalarResult r = v.fabric->random_shared_scalars(v.n);
  ScalarResult opened = open_authenticated_batch(batch_mul(v, r)).result();
  return batch_mul_public(r, ScalarResult::batch_inverse(opened));
}
inline AuthenticatedScalarResult AuthenticatedScalarResult::batch_div(const AuthenticatedScalarResult& a, const AuthenticatedScalarResult& b) {
  return batch_mul(a, batch_inverse(b));
}

inline AuthenticatedScalarOpenResult AuthenticatedScalarResult::open_authenticated_batch(const AuthenticatedScalarResult& v) {
  MpcFabric* f = v.fabric;
  const size_t n = v.n;
  if (n == 0) return {ScalarResult{f, f->alloc(0), 0}, true};
  const int fid = f->curve().field;
  ScalarResult opened = open_batch(v);
  Buf checks = f->alloc(n * 32), bytes = f->alloc(n * 32);
  f->ctx()->check(arkmpc_fr_mac_check(f->raw(), fid, f->mac_key().data(), n, opened.values->u64(), v.mac->u64(), checks->u64()), "arkmpc_fr_mac_check");  // :299-311
  f->ctx()->check(arkmpc_fr_to_bytes_be(f->raw(), fid, n, checks->u64(), reinterpret_cast<uint8_t*>(bytes->u64())), "arkmpc_fr_to_bytes_be");
  auto to_bytes = [&](const Buf& b) { auto w = f->download(b, n * 32); std::vector<uint8_t> o(n * 32); memcpy(o.data(), w.data(), n * 32); return o; };
  Limbs blinder = f->random_blinder();
  Limbs my_comm = f->commit(to_bytes(bytes), blinder);
  auto peer_comm = f->exchange_host(std::vector<uint64_t>(my_comm.begin(), my_comm.end()));
  Buf peer_checks = f->exchange(checks);                                                              // :323
  auto peer_blinder = f->exchange_host(std::vector<uint64_t>(blinder.begin(), blinder.end()));
  // batch_verify_mac_check (:201-220)
  Buf peer_bytes = f->alloc(n * 32);
  f->ctx()->check(arkmpc_fr_to_bytes_be(f->raw(), fid, n, peer_checks->u64(), reinterpret_cast<uint8_t*>(peer_bytes->u64())), "arkmpc_fr_to_bytes_be");
  Limbs pb{peer_blinder[0], peer_blinder[1], peer_blinder[2], peer_blinder[3]};
  Limbs expect = f->commit(to_bytes(peer_bytes), pb);
  bool ok = std::equal(expect.begin(), expect.end(), peer_comm.begin());
  int zero = 0;
  f->ctx()->check(arkmpc_fr_sum_is_zero(f->raw(), fid, n, checks->u64(), peer_checks->u64(), &zero), "arkmpc_fr_sum_is_zero");
  f->count_gate(3);
  return {opened, ok && zero == 1};
}

inline AuthenticatedScalarResult AuthenticatedScalarResult::sum() const {
  Buf s = fabric->alloc(32), m = fabric->alloc(32);
  fabric->ctx()->check(arkmpc_fr_share_sum(fabric->raw(), fabric->curve().field, n, share->u64(), mac->u64(), s->u64(), m->u64()), "arkmpc_fr_share_sum");
  fabric->count_gate();
  return {fabric, s, m, 1};
}

inline AuthenticatedScalarResult AuthenticatedScalarResult::fft(const AuthenticatedScalarResult& x, bool inverse) {
  MpcFabric* f = x.fabric;
  if (x.n == 0 || (x.n & (x.n - 1))) throw std::invalid_argument("fft: the length must be a non-zero power of two");
  int log2n = 0;
  while (((size_t)1 << log2n) < x.n) log2n++;
  Buf s = f->alloc(x.n * 32), m = f->alloc(x.n * 32);
  f->ctx()->check(arkmpc_fr_share_fft(f->raw(), f->curve().field, log2n, inverse ? 1 : 0, x.share->u64(), x.mac->u64(), s->u64(), m->u64()), "arkmpc_fr_share_fft");
  f->count_gate();
  return {f, s, m, x.n};
}

// ---- points ----
inline CurvePointResult CurvePointResult::msm(const ScalarResult& scalars, const CurvePointResult& points) {
  detail::same_len(scalars.n, points.n, "msm");
  MpcFabric* f = points.fabric;
  Buf o = f->alloc(f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_msm(f->raw(), f->curve().curve, points.n, scalars.values->u64(), points.points->u64(), o->u64()), "arkmpc_pt_msm");
  f->count_gate();
  return {f, o, 1};
}
inline std::vector<uint64_t> CurvePointResult::to_affine_host() const {
  Buf xy = fabric->alloc(n * 64);
  fabric->ctx()->check(arkmpc_pt_normalize(fabric->raw(), fabric->curve().curve, n, points->u64(), xy->u64()), "arkmpc_pt_normalize");
  return fabric->download(xy, n * 64);
}

namespace detail {
using PtBin = int (*)(arkmpc_ctx*, int, size_t, const uint64_t*, const uint64_t*, uint64_t*);
inline AuthenticatedPointResult pshare_binary(PtBin fn, const char* what, const AuthenticatedPointResult& a, const AuthenticatedPointResult& b) {
  same_len(a.n, b.n, what);
  MpcFabric* f = a.fabric;
  Buf o = f->alloc(a.n * 2 * f->curve().point_words * 8);
  f->ctx()->check(fn(f->raw(), f->curve().curve, 2 * a.n, a.shares->u64(), b.shares->u64(), o->u64()), what);  // the 2n points of n PointShares
  f->count_gate();
  return {f, o, a.n};
}
}  // namespace detail

inline AuthenticatedPointResult AuthenticatedPointResult::batch_add(const AuthenticatedPointResult& a, const AuthenticatedPointResult& b) {
  return detail::pshare_binary(arkmpc_pt_add, "batch_add", a, b);
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_sub(const AuthenticatedPointResult& a, const AuthenticatedPointResult& b) {
  return detail::pshare_binary(arkmpc_pt_sub, "batch_sub", a, b);
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_neg(const AuthenticatedPointResult& a) {
  MpcFabric* f = a.fabric;
  Buf o = f->alloc(a.n * 2 * f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_neg(f->raw(), f->curve().curve, 2 * a.n, a.shares->u64(), o->u64()), "batch_neg");
  f->count_gate();
  return {f, o, a.n};
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_add_public(const AuthenticatedPointResult& a, const CurvePointResult& b) {
  detail::same_len(a.n, b.n, "batch_add_public");
  MpcFabric* f = a.fabric;
  Buf o = f->alloc(a.n * 2 * f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_share_add_public(f->raw(), f->curve().curve, f->party_id(), f->mac_key().data(), a.n, a.shares->u64(), b.points->u64(), o->u64()),
                  "batch_add_public");
  f->count_gate();
  return {f, o, a.n};
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_mul_public(const ScalarResult& a, const AuthenticatedPointResult& b) {
  detail::same_len(a.n, b.n, "batch_mul_public");
  MpcFabric* f = b.fabric;
  Buf o = f->alloc(b.n * 2 * f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_share_mul_public(f->raw(), f->curve().curve, b.n, a.values->u64(), b.shares->u64(), o->u64()), "batch_mul_public");
  f->count_gate();
  return {f, o, b.n};
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_mul_generator(const AuthenticatedScalarResult& a) {
  MpcFabric* f = a.fabric;
  Buf o = f->alloc(a.n * 2 * f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_mul_generator(f->raw(), f->curve().curve, a.n, a.share->u64(), a.mac->u64(), o->u64()), "batch_mul_generator");
  f->count_gate();
  return {f, o, a.n};
}
inline AuthenticatedPointResult AuthenticatedPointResult::batch_mul(const AuthenticatedScalarResult& a, const AuthenticatedPointResult& b) {
  detail::same_len(a.n, b.n, "Batch add");  // the reference's message (:689)
  MpcFabric* f = a.fabric;
  const size_t n = a.n, pb = f->curve().point_words * 8;
  if (n == 0) return {f, f->alloc(0), 0};
  auto [ba, bb, bc] = f->next_triple_batch(n);
  const int cid = f->curve().curve;
  Buf d_mine = f->alloc(n * 32), E_mine = f->alloc(n * pb);
  f->ctx()->check(arkmpc_pt_beaver_mask(f->raw(), cid, n, a.share->u64(), b.shares->u64(), ba.share->u64(), bb.share->u64(), d_mine->u64(), E_mine->u64()),
                  "arkmpc_pt_beaver_mask");
  Buf E_peer = f->exchange(E_mine);  // open_batch of the masked points (:66-109)
  Buf d_peer = f->exchange(d_mine);  // open_batch of the masked scalars
  Buf out = f->alloc(n * 2 * pb);
  f->ctx()->check(arkmpc_pt_beaver_recombine(f->raw(), cid, f->party_id(), f->mac_key().data(), n, d_mine->u64(), d_peer->u64(), E_mine->u64(), E_peer->u64(),
                                             ba.share->u64(), ba.mac->u64(), bb.share->u64(), bb.mac->u64(), bc.share->u64(), bc.mac->u64(), out->u64(),
                                             nullptr, nullptr), "arkmpc_pt_beaver_recombine");
  f->count_gate(2);
  return {f, out, n};
}
inline CurvePointResult AuthenticatedPointResult::open_batch(const AuthenticatedPointResult& v) {
  MpcFabric* f = v.fabric;
  const size_t n = v.n, pb = f->curve().point_words * 8;
  if (n == 0) return {f, f->alloc(0), 0};
  Buf mine = f->alloc(n * pb);  // this party's share points (:75-77 sends `share.share()` only)
  f->ctx()->check(arkmpc_pt_share_split(f->raw(), f->curve().curve, n, v.shares->u64(), mine->u64(), nullptr), "arkmpc_pt_share_split");
  Buf peer = f->exchange(mine);
  Buf o = f->alloc(n * pb);
  f->ctx()->check(arkmpc_pt_add(f->raw(), f->curve().curve, n, mine->u64(), peer->u64(), o->u64()), "open_batch");
  f->count_gate();
  return {f, o, n};
}
inline AuthenticatedPointOpenResult AuthenticatedPointResult::open_authenticated_batch(const AuthenticatedPointResult& v) {
  MpcFabric* f = v.fabric;
  const size_t n = v.n, pb = f->curve().point_words * 8;
  if (n == 0) return {CurvePointResult{f, f->alloc(0), 0}, true};
  const int cid = f->curve().curve;
  CurvePointResult opened = open_batch(v);
  Buf checks = f->alloc(n * pb);
  f->ctx()->check(arkmpc_pt_mac_check(f->raw(), cid, f->mac_key().data(), n, opened.points->u64(), v.shares->u64(), checks->u64()), "arkmpc_pt_mac_check");  // :217-232
  // The reference hashes arkworks' compressed encoding of each check point (commitment.rs, ToBytes); that hashing stays on the
  // host in the Rust integration.  This mirror commits to the canonical affine limbs of the whole vector.
  auto affine_bytes = [&](const Buf& pts) {
    CurvePointResult r{f, pts, n};
    auto w = r.to_affine_host();
    std::vector<uint8_t> o(w.size() * 8);
    memcpy(o.data(), w.data(), o.size());
    return o;
  };
  Limbs blinder = f->random_blinder();
  Limbs my_comm = f->commit(affine_bytes(checks), blinder);
  auto peer_comm = f->exchange_host(std::vector<uint64_t>(my_comm.begin(), my_comm.end()));
  Buf peer_checks = f->exchange(checks);
  auto peer_blinder = f->exchange_host(std::vector<uint64_t>(blinder.begin(), blinder.end()));
  Limbs pbl{peer_blinder[0], peer_blinder[1], peer_blinder[2], peer_blinder[3]};
  Limbs expect = f->commit(affine_bytes(peer_checks), pbl);
  bool ok = std::equal(expect.begin(), expect.end(), peer_comm.begin());
  int ident = 0;
  f->ctx()->check(arkmpc_pt_sum_is_identity(f->raw(), cid, n, checks->u64(), peer_checks->u64(), &ident), "arkmpc_pt_sum_is_identity");  // :128-131
  f->count_gate(3);
  return {opened, ok && ident == 1};
}
inline AuthenticatedPointResult CurvePointResult::msm_authenticated(const AuthenticatedScalarResult& scalars, const CurvePointResult& points) {
  detail::same_len(scalars.n, points.n, "msm");
  MpcFabric* f = points.fabric;
  Buf o = f->alloc(2 * f->curve().point_words * 8);
  f->ctx()->check(arkmpc_pt_msm_authenticated(f->raw(), f->curve().curve, points.n, scalars.share->u64(), scalars.mac->u64(), points.points->u64(), o->u64()),
                  "arkmpc_pt_msm_authenticated");
  f->count_gate();
  return {f, o, 1};
}
#undef ARKMPC_F

// ---------------------------------------------------------------------------------------------------------------
// Two-party in-process harness (lib.rs:116-201): runs `f(fabric)` for both parties on two threads and returns both results.
// `make_source(party_id)` builds the preprocessing source.
// ---------------------------------------------------------------------------------------------------------------
template <class T>
std::pair<T, T> execute_mock_mpc(const CurveInfo& cv, std::function<std::unique_ptr<PreprocessingPhase>(int)> make_source,
                                 std::function<T(MpcFabric&)> f, int device = 0) {
  auto nets = MockNetwork::new_duplex_pair();
  MockNetwork net[2] = {nets.first, nets.second};
  T results[2];
  std::exception_ptr errors[2];
  auto run = [&](int p) {
    std::unique_ptr<MpcFabric> fabric;
    try {
      fabric = std::make_unique<MpcFabric>(net[p], make_source(p), cv, device);
      results[p] = f(*fabric);
      fabric->ctx()->sync();
    } catch (...) {
      errors[p] = std::current_exception();
      if (fabric) fabric->poison_peer(); else net[p].poison();
    }
  };
  std::thread t0(run, 0), t1(run, 1);
  t0.join();
  t1.join();
  for (auto& e : errors)
    if (e) std::rethrow_exception(e);
  return {results[0], results[1]};
}

}  // namespace arkmpc
