// Thread-safety of one context (gate closures run on arbitrary executor workers,
// /root/reference/online-phase/src/fabric/executor/multi_threaded/executor.rs:208-217): T host threads drive ONE
// arkmpc_ctx concurrently, each with its own batches of authenticated Beaver multiplications (upload, mask, fused
// recombine of one party against host-computed peer values, download), interleaved with MAC-check / sum / sum_is_zero
// calls that use the context's shared scratch.  Every thread's results must equal the CPU oracle's
// (oracle/ark_oracle.c, test infrastructure) limb for limb.  Exit code 0 = OK, 2 = no GPU.
// Build (tests/test_host_cpp.py): g++ -std=c++17 -Iinclude tests/host_cpp/test_threads.cpp -Lark_mpc_b200/lib -larkmpc_b200 -Loracle -lark_oracle -lpthread
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "arkmpc_b200.h"

extern "C" {
void orc_synth(int f, uint64_t seed, uint64_t first_index, size_t n, uint64_t* out);
void orc_beaver_mask(int f, size_t n, const uint64_t* x, const uint64_t* y, const uint64_t* a, const uint64_t* b, uint64_t* d_mine,
                     uint64_t* e_mine, uint64_t* scratch);
void orc_beaver_recombine(int f, int party, const uint64_t* key, size_t n, const uint64_t* d, const uint64_t* e, const uint64_t* a,
                          const uint64_t* b, const uint64_t* c, uint64_t* out, uint64_t* scratch);
void orc_scalar_batch_add(int f, size_t n, uint64_t* o, const uint64_t* a, const uint64_t* b);
void orc_mac_check(int f, const uint64_t* key, size_t n, const uint64_t* opened, const uint64_t* shares, uint64_t* out);
void orc_share_sum(int f, size_t n, const uint64_t* shares, uint64_t* out);
}

static std::atomic<int> failures{0};
#define CHECK(cond, msg)                                                                  \
  do {                                                                                    \
    if (!(cond)) { printf("FAIL[t%d it%d]: %s (line %d)\n", tid, it, msg, __LINE__); failures++; return; } \
  } while (0)
#define OKCALL(expr) CHECK((expr) == ARKMPC_OK, #expr)

struct Dev {
  arkmpc_ctx* ctx;
  std::vector<void*> ptrs;
  uint64_t* alloc(size_t bytes) {
    void* p = nullptr;
    if (arkmpc_malloc(ctx, bytes, &p) != ARKMPC_OK) return nullptr;
    ptrs.push_back(p);
    return static_cast<uint64_t*>(p);
  }
  ~Dev() { for (void* p : ptrs) arkmpc_free(ctx, p); }
};

static void worker(arkmpc_ctx* ctx, int tid, int field, int iters) {
  for (int it = 0; it < iters; it++) {
    const size_t n = 1000 + 37 * (size_t)tid + 211 * (size_t)it;  // ragged sizes, different on every thread
    const int party = (tid + it) & 1;
    const uint64_t seed = 1000u * (uint64_t)tid + 10u * (uint64_t)it;
    // AoS host images of this party's shares (share, mac interleaved), the reference's memory layout
    std::vector<uint64_t> x(n * 8), y(n * 8), a(n * 8), b(n * 8), c(n * 8), dp(n * 4), ep(n * 4), key(4);
    orc_synth(field, seed + 1, 0, 2 * n, x.data());
    orc_synth(field, seed + 2, 0, 2 * n, y.data());
    orc_synth(field, seed + 3, 0, 2 * n, a.data());
    orc_synth(field, seed + 4, 0, 2 * n, b.data());
    orc_synth(field, seed + 5, 0, 2 * n, c.data());
    orc_synth(field, seed + 6, 0, n, dp.data());
    orc_synth(field, seed + 7, 0, n, ep.data());
    orc_synth(field, seed + 8, 0, 1, key.data());
    // oracle: unfused reference sequence
    std::vector<uint64_t> dm(n * 4), em(n * 4), d(n * 4), e(n * 4), want(n * 8), scratch(n * 4 + 5 * n * 8), mscratch(2 * n * 8);
    orc_beaver_mask(field, n, x.data(), y.data(), a.data(), b.data(), dm.data(), em.data(), mscratch.data());
    orc_scalar_batch_add(field, n, d.data(), dm.data(), dp.data());
    orc_scalar_batch_add(field, n, e.data(), em.data(), ep.data());
    orc_beaver_recombine(field, party, key.data(), n, d.data(), e.data(), a.data(), b.data(), c.data(), want.data(), scratch.data());
    std::vector<uint64_t> want_chk(n * 4), want_sum(8);
    orc_mac_check(field, key.data(), n, d.data(), want.data(), want_chk.data());
    orc_share_sum(field, n, want.data(), want_sum.data());

    Dev D{ctx, {}};
    uint64_t *xa = D.alloc(n * 64), *ya = D.alloc(n * 64), *aa = D.alloc(n * 64), *ba = D.alloc(n * 64), *ca = D.alloc(n * 64);
    uint64_t* pl[10];
    for (auto& p : pl) p = D.alloc(n * 32);  // x_s x_m y_s y_m a_s a_m b_s b_m c_s c_m
    uint64_t *d_m = D.alloc(n * 32), *e_m = D.alloc(n * 32), *d_p = D.alloc(n * 32), *e_p = D.alloc(n * 32);
    uint64_t *o_s = D.alloc(n * 32), *o_m = D.alloc(n * 32), *d_o = D.alloc(n * 32), *e_o = D.alloc(n * 32), *o_aos = D.alloc(n * 64);
    uint64_t *chk = D.alloc(n * 32), *neg = D.alloc(n * 32), *sum_s = D.alloc(32), *sum_m = D.alloc(32);
    CHECK(sum_m != nullptr, "device allocation");
    OKCALL(arkmpc_memcpy_h2d(ctx, xa, x.data(), n * 64));
    OKCALL(arkmpc_memcpy_h2d(ctx, ya, y.data(), n * 64));
    OKCALL(arkmpc_memcpy_h2d(ctx, aa, a.data(), n * 64));
    OKCALL(arkmpc_memcpy_h2d(ctx, ba, b.data(), n * 64));
    OKCALL(arkmpc_memcpy_h2d(ctx, ca, c.data(), n * 64));
    OKCALL(arkmpc_memcpy_h2d(ctx, d_p, dp.data(), n * 32));
    OKCALL(arkmpc_memcpy_h2d(ctx, e_p, ep.data(), n * 32));
    OKCALL(arkmpc_share_unzip(ctx, n, xa, pl[0], pl[1]));
    OKCALL(arkmpc_share_unzip(ctx, n, ya, pl[2], pl[3]));
    OKCALL(arkmpc_share_unzip(ctx, n, aa, pl[4], pl[5]));
    OKCALL(arkmpc_share_unzip(ctx, n, ba, pl[6], pl[7]));
    OKCALL(arkmpc_share_unzip(ctx, n, ca, pl[8], pl[9]));
    OKCALL(arkmpc_fr_beaver_mask(ctx, field, n, pl[0], pl[2], pl[4], pl[6], d_m, e_m));
    std::this_thread::yield();  // let another thread's calls land between the two phases
    OKCALL(arkmpc_fr_beaver_recombine(ctx, field, party, key.data(), n, d_m, e_m, d_p, e_p, pl[4], pl[5], pl[6], pl[7], pl[8], pl[9], o_s, o_m,
                                      d_o, e_o));
    OKCALL(arkmpc_share_zip(ctx, n, o_s, o_m, o_aos));
    OKCALL(arkmpc_fr_mac_check(ctx, field, key.data(), n, d_o, o_m, chk));
    OKCALL(arkmpc_fr_share_sum(ctx, field, n, o_s, o_m, sum_s, sum_m));  // uses the context's shared partial-sum scratch
    OKCALL(arkmpc_fr_neg(ctx, field, n, chk, neg));
    int zero = 0;
    OKCALL(arkmpc_fr_sum_is_zero(ctx, field, n, chk, neg, &zero));       // shared flag word, synchronises
    CHECK(zero == 1, "chk + (-chk) must be all zero");
    std::vector<uint64_t> got(n * 8), got_d(n * 4), got_chk(n * 4), got_sum(8);
    OKCALL(arkmpc_memcpy_d2h(ctx, got.data(), o_aos, n * 64));
    OKCALL(arkmpc_memcpy_d2h(ctx, got_d.data(), d_o, n * 32));
    OKCALL(arkmpc_memcpy_d2h(ctx, got_chk.data(), chk, n * 32));
    OKCALL(arkmpc_memcpy_d2h(ctx, got_sum.data(), sum_s, 32));
    OKCALL(arkmpc_memcpy_d2h(ctx, got_sum.data() + 4, sum_m, 32));
    OKCALL(arkmpc_ctx_sync(ctx));
    CHECK(got == want, "recombine output differs from the oracle");
    CHECK(got_d == d, "opened d differs from the oracle");
    CHECK(got_chk == want_chk, "MAC-check vector differs from the oracle");
    CHECK(got_sum == want_sum, "share sum differs from the oracle");
    // errors are per thread: a bad call here must not disturb the others, and its message must be ours
    CHECK(arkmpc_fr_beaver_mask(ctx, 99, n, pl[0], pl[2], pl[4], pl[6], d_m, e_m) == ARKMPC_ERR_INVALID, "unknown field must be rejected");
    CHECK(strstr(arkmpc_last_error(ctx), "field") != nullptr, "per-thread error text");
  }
}

int main(int argc, char** argv) {
  int count = 0;
  if (arkmpc_device_count(&count) != ARKMPC_OK || count == 0) {
    printf("no CUDA device: the gate engine has no CPU fallback\n");
    return 2;
  }
  const int threads = argc > 1 ? atoi(argv[1]) : 4, iters = argc > 2 ? atoi(argv[2]) : 6;
  arkmpc_ctx* ctx = nullptr;
  if (arkmpc_ctx_create(0, &ctx) != ARKMPC_OK) { printf("ctx_create failed\n"); return 1; }
  for (int field = 0; field < 2; field++) {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) th.emplace_back(worker, ctx, t, field, iters);
    for (auto& t : th) t.join();
  }
  const unsigned long long launches = arkmpc_ctx_launch_count(ctx);
  arkmpc_ctx_destroy(ctx);
  if (failures) { printf("%d failure(s)\n", failures.load()); return 1; }
  printf("thread-safety test OK: %d threads x %d iterations x 2 fields on one context, %llu kernels launched, all results equal the oracle\n",
         threads, iters, launches);
  return 0;
}
