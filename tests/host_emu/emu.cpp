// Host-emulation harness: compiles the SAME per-element device functions (fp256.cuh, beaver.cuh)
// with the carry flag emulated in software, and exports them with a C ABI so tests/test_host_emu.py
// can check the exact instruction sequences against the oracle on a CPU-only box.
// Test infrastructure; not part of the product library.
#include <cstring>
#include "../../ark_mpc_b200/csrc/beaver.cuh"
#include "../../ark_mpc_b200/csrc/ctab.hpp"
using namespace ark;

// constant-multiplier table from an 8 x u32 value, through the product's own host builder (ctab.hpp)
template <class F> static CTab tab_of(const fe8& s) {
  uint64_t h[4];
  for (int j = 0; j < 4; j++) h[j] = (uint64_t)s.v[2 * j] | ((uint64_t)s.v[2 * j + 1] << 32);
  CTab t;
  ctab_build<F>(t, h);
  return t;
}

template <class F> static void run(int op, const uint32_t* in, uint32_t* out, int party) {
  const fe8* v = reinterpret_cast<const fe8*>(in);
  fe8* o = reinterpret_cast<fe8*>(out);
  switch (op) {
    case 0: Fp<F>::add(o[0], v[0], v[1]); break;
    case 1: Fp<F>::sub(o[0], v[0], v[1]); break;
    case 2: Fp<F>::neg(o[0], v[0]); break;
    case 3: Fp<F>::mul(o[0], v[0], v[1]); break;
    case 4: Fp<F>::mul_lazy(o[0], v[0], v[1]); break;
    case 5: if constexpr (F::kLazy2) { Fp<F>::mul2_lazy(o[0], v[0], v[1], v[2], v[3]); } break;
    case 6: beaver_mask_elem<F>(o[0], o[1], v[0], v[1], v[2], v[3]); break;
    case 7: if constexpr (F::kLazy2) {
      // in: key, d_mine, e_mine, d_peer, e_peer, a_s, a_m, b_s, b_m, c_s, c_m ; out: out_s, out_m, d, e
      beaver_recombine_elem<F>(o[0], o[1], o[2], o[3], party, tab_of<F>(v[0]), v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10]);
    } break;
    case 8: if constexpr (F::kBits <= 254) { share_add_public_elem<F>(o[0], o[1], party, tab_of<F>(v[0]), v[1], v[2], v[3]); } break;
    case 9: if constexpr (F::kBits <= 254) { share_sub_public_elem<F>(o[0], o[1], party, tab_of<F>(v[0]), v[1], v[2], v[3]); } break;
    case 10: if constexpr (F::kBits <= 254) { mac_check_elem<F>(o[0], tab_of<F>(v[0]), v[1], v[2]); } break;
    case 11: o[0] = v[0]; Fp<F>::csub_p(o[0]); break;
    case 12: Fp<F>::set_one(o[0]); Fp<F>::set_r2(o[1]); break;
    case 13: Fp<F>::sqr(o[0], v[0]); break;
    case 17: Fp<F>::inv_plain(o[0], v[0]); break;
    case 18: Fp<F>::inv_mont(o[0], v[0]); break;
    case 19: Fp<F>::inv_safegcd(o[0], v[0]); break;
    case 20: Fp<F>::mul_kara(o[0], v[0], v[1]); break;
    case 21: { uint32_t T[16]; kara512(T, v[0].v, v[1].v); memcpy(o, T, sizeof T); } break;  // the 512-bit product (two elements)
    case 14: if constexpr (F::kBits <= 254) { Fp<F>::mul_ctab_lazy(o[0], tab_of<F>(v[0]), v[1]); } break;  // lazy: s * a / R, any 256-bit a
    case 15: if constexpr (F::kBits <= 254) { Fp<F>::mul_ctab(o[0], tab_of<F>(v[0]), v[1]); } break;
    case 16: if constexpr (F::kBits <= 254) { CTab t = tab_of<F>(v[0]); memcpy(o, &t, sizeof t); } break;       // the table itself (8 elements)
  }
}

extern "C" uint64_t emu_violations() { return ark::emu::violations; }

extern "C" int emu_run(int field, int op, int party, const uint32_t* in, uint32_t* out) {
  switch (field) {
    case 0: run<Bn254Fr>(op, in, out, party); return 0;
    case 1: run<Curve25519Fr>(op, in, out, party); return 0;
    case 2: run<Bn254Fq>(op, in, out, party); return 0;
    case 3: run<Curve25519Fq>(op, in, out, party); return 0;
  }
  return -1;
}

// ---- curve gates (curve.cuh / curve_gates.cuh) ----
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ark_mpc_b200/csrc/curve_gates.cuh"

template <class C> static const typename C::Aff* host_gtab() {
  static std::vector<typename C::Aff> tab;
  if (tab.empty()) {
    // Built incrementally (entry w = entry w - 1 + base_j: one addition and one normalisation each; build_gtab_entry's 12 j
    // doublings per entry would take minutes under emulation) and checked against build_gtab_entry on a sample of entries.
    tab.resize(kFixWindows * kFixEntries);
    typename C::Pt base;
    C::set_generator(base);
    for (int j = 0; j < kFixWindows; j++) {
      typename C::Pt acc;
      C::set_identity(acc);
      for (uint32_t w = 1; w < (uint32_t)kFixEntries; w++) {
        C::add(acc, base);
        C::to_aff(tab[j * kFixEntries + w], acc);
      }
      for (uint32_t w : {1u, 2u, 255u, 256u, 2049u, (uint32_t)kFixEntries - 1u}) {
        typename C::Aff chk;
        build_gtab_entry<C>(chk, j, w);
        if (memcmp(&chk, &tab[j * kFixEntries + w], sizeof chk) != 0) { fprintf(stderr, "fixed-base table mismatch at (%d, %u)\n", j, w); abort(); }
      }
      for (int i = 0; i < kFixBits; i++) C::dbl(base);
    }
  }
  return tab.data();
}

template <class C> static void run_curve(int op, const uint32_t* in, uint32_t* out, int party) {
  using Pt = typename C::Pt;
  const fe8* v = reinterpret_cast<const fe8*>(in);
  fe8* o = reinterpret_cast<fe8*>(out);
  constexpr int K = C::kCoords;
  // points cross the harness in the reference's memory image, exactly as they cross the kernels' loads and stores
  auto pt = [&](int at) { Pt p; memcpy(&p, v + at, sizeof(Pt)); C::from_image(p); return p; };
  auto put = [&](int at, const Pt& p_in) { Pt p = p_in; C::to_image(p); memcpy(o + at, &p, sizeof(Pt)); };
  switch (op) {
    case 0: { Pt p = pt(0); C::add(p, pt(K)); put(0, p); } break;
    case 1: { Pt p = pt(0); C::dbl(p); put(0, p); } break;
    case 2: { C::normalize(o[0], o[1], pt(0)); } break;
    case 3: { Pt r; pt_mul_elem<C>(r, v[0], pt(1)); put(0, r); } break;
    case 4: { Pt r; pt_mul_gen_elem<C>(r, v[0], host_gtab<C>()); put(0, r); } break;
    case 5: { Pt E; pt_beaver_mask_elem<C>(o[0], E, v[0], v[1], v[2], pt(3), host_gtab<C>()); put(1, E); } break;
    case 6: {
      // in: key d_mine d_peer a_s a_m b_s b_m c_s c_m E_mine E_peer ; out: d, E, out_s, out_m
      Pt E;
      pt_beaver_recombine_elem<C, false>(o[0], E, party, v[0], v[1], v[2], pt(9), pt(9 + K), v[3], v[4], v[5], v[6], v[7], v[8], host_gtab<C>(),
                                  [&](int which, const Pt& r) { put(1 + K + which * K, r); });
      put(1, E);
    } break;
    case 7: case 8: { Pt s, m; pt_share_add_public_elem<C>(s, m, party, op == 8, v[0], pt(1), pt(1 + K), pt(1 + 2 * K)); put(0, s); put(K, m); } break;
    case 9: { Pt r; pt_mac_check_elem<C>(r, v[0], pt(1), pt(1 + K)); put(0, r); } break;
    case 10: { Pt r0, r1; pt_mul2_elem<C>(r0, r1, v[0], v[1], pt(2)); put(0, r0); put(K, r1); } break;
    case 13: {  // as case 6 with the dual-chain variable-base passes
      Pt E;
      pt_beaver_recombine_elem<C, true>(o[0], E, party, v[0], v[1], v[2], pt(9), pt(9 + K), v[3], v[4], v[5], v[6], v[7], v[8], host_gtab<C>(),
                                        [&](int which, const Pt& r) { put(1 + K + which * K, r); });
      put(1, E);
    } break;
    case 11: { Pt p; C::set_generator(p); put(0, p); C::set_identity(p); put(K, p); } break;
    case 12: { Pt p = pt(0); C::neg(p); put(0, p); } break;
    case 14: { o[0].v[0] = pt_valid_elem<C>(pt(0)) ? 1u : 0u; } break;
  }
}

// F25519 (special-form arithmetic modulo 2p): in = plain 256-bit values, out = raw results (and canonical forms)
extern "C" int emu_f25519(int op, const uint32_t* in, uint32_t* out) {
  const fe8* v = reinterpret_cast<const fe8*>(in);
  fe8* o = reinterpret_cast<fe8*>(out);
  switch (op) {
    case 0: F25519::add(o[0], v[0], v[1]); break;
    case 1: F25519::sub(o[0], v[0], v[1]); break;
    case 2: F25519::mul(o[0], v[0], v[1]); break;
    case 3: F25519::neg(o[0], v[0]); break;
    case 4: F25519::canon(o[0], v[0]); break;
    case 5: F25519::inv(o[0], v[0]); break;
    case 6: F25519::from_image(o[0], v[0]); break;
    case 7: F25519::to_image(o[0], v[0]); break;
    case 8: o[0].v[0] = F25519::is_zero(v[0]); o[0].v[1] = F25519::eq(v[0], v[1]); break;
    case 9: F25519::sqr(o[0], v[0]); break;
    case 10: sqr512(reinterpret_cast<uint32_t*>(o), v[0].v); break;
    default: return -1;
  }
  return 0;
}

extern "C" int emu_curve(int curve, int op, int party, const uint32_t* in, uint32_t* out) {
  switch (curve) {
    case 0: run_curve<Bn254G1>(op, in, out, party); return 0;
    case 1: run_curve<Ed25519>(op, in, out, party); return 0;
  }
  return -1;
}
