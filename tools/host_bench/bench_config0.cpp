// BASELINE.json configs[0] through the C++ host mirror (host/arkmpc_host.hpp) — the reference bench's shape
// (/root/reference/online-phase/benches/batch_ops.rs:20-40): n random scalars over Curve25519 Fr (or BN254 Fr), both vectors shared by
// party 0, AuthenticatedScalarResult::batch_mul, open_authenticated_batch, results on the host; PartyIDBeaverSource; in-memory mock
// network; the clock starts inside each party's closure and the slower party counts.  A compiled host (what a Rust integration is)
// instead of the Python mirror: at n = 1024 the path is bound by host-side latency, not arithmetic.
// Prints one JSON object.  Build: g++ -O2 -std=c++17 -Iinclude -Ihost tools/host_bench/bench_config0.cpp -Lark_mpc_b200/lib -larkmpc_b200 -lpthread
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "arkmpc_host.hpp"

using namespace arkmpc;

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 1024;
  const int iters = argc > 2 ? atoi(argv[2]) : 30;
  const bool c25519 = argc <= 3 || atoi(argv[3]) == 1;
  int count = 0;
  if (arkmpc_device_count(&count) != ARKMPC_OK || count == 0) { printf("{\"error\": \"no CUDA device: no CPU fallback\"}\n"); return 2; }
  const CurveInfo cv = c25519 ? curve25519() : bn254();
  // random canonical scalars (Montgomery images of small random integers are as good as any for timing; the top limb stays below p's)
  std::mt19937_64 rng(1024);
  HostScalars a, b;
  a.limbs.resize(n * 4);
  b.limbs.resize(n * 4);
  for (size_t i = 0; i < n * 4; i++) { a.limbs[i] = rng(); b.limbs[i] = rng(); }
  for (size_t i = 0; i < n; i++) { a.limbs[4 * i + 3] &= 0x0fffffffffffffffull; b.limbs[4 * i + 3] &= 0x0fffffffffffffffull; }
  auto source = [cv](int p) { return std::unique_ptr<PreprocessingPhase>(new PartyIDBeaverSource(p, cv)); };
  std::vector<double> times;
  std::vector<uint64_t> first;
  for (int it = 0; it < iters + 3; it++) {
    if (it == 3) prof::reset();
    using Out = std::pair<double, std::vector<uint64_t>>;
    auto res = execute_mock_mpc<Out>(cv, source, [&](MpcFabric& f) {
      const auto t0 = std::chrono::steady_clock::now();
      ScalarResult va = f.allocate_scalars(a), vb = f.allocate_scalars(b);
      auto A = f.batch_share_scalar(f.party_id() == 0 ? &va : nullptr, n, 0);
      auto B = f.batch_share_scalar(f.party_id() == 0 ? &vb : nullptr, n, 0);
      using S = AuthenticatedScalarResult;
      std::vector<uint64_t> opened = S::open_authenticated_batch(S::batch_mul(A, B)).result().to_host();  // host copy = "await all"
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      return Out{dt, opened};
    });
    if (res.first.second != res.second.second) { printf("{\"error\": \"the parties opened different values\"}\n"); return 1; }
    if (it == 0) first = res.first.second;
    if (it >= 3) times.push_back(std::max(res.first.first, res.second.first));
  }
  std::sort(times.begin(), times.end());
  const double med = times[times.size() / 2];
  // a checksum of the opened products so that the caller can compare them with the oracle
  uint64_t chk = 0;
  for (uint64_t w : first) chk = chk * 1000003u + w;
  printf("{\"n\": %zu, \"iters\": %d, \"field\": \"%s\", \"median_s\": %.9f, \"min_s\": %.9f, \"mults_per_s\": %.1f, \"opened_checksum\": \"%016llx\"}\n", n,
         iters, c25519 ? "curve25519_fr" : "bn254_fr", med, times.front(), n / med, (unsigned long long)chk);
  if (prof::enabled()) {
    fprintf(stderr, "host profile, per iteration, both parties' threads summed (wall %.1f us per iteration):\n", med * 1e6);
    prof::report(stderr, iters);
  }
  return 0;
}
