// Per-element point gates (one gate per thread), shared by the kernels in curve_kernels.cuh and by the
// host-emulation harness.  Reference behaviour restated (paths under /root/reference/online-phase/src/algebra):
//   CurvePoint * Scalar                      curve/curve.rs:403-409
//   ScalarShare * CurvePoint                 scalar/share.rs:135-141   -> (share*P, mac*P)
//   PointShare * Scalar                      curve/share.rs:107-113    -> (s*share, s*mac)
//   PointShare::add_public                   curve/share.rs:57-60      -> share + P on party 0 only; mac + mac_key*P
//   AuthenticatedPointResult::batch_mul      curve/authenticated_curve.rs:682-714
//       mask       d_mine = x.share - a.share ;  E_mine = P.share - b.share*G            (:696-700; open sends shares only, :66-109)
//       recombine  [x*P] = d*E + d[bG] + [a]E + [c]G  with add_public for the public d*E   (:704-713)
// The recombination regroups the reference's 4 fixed-base + 6 variable-base scalar multiplications per element into
// two double-scalar passes that share one table of E:
//       share = (a.share [+ d on party 0]) * E + (d*b.share + c.share) * G
//       mac   = (key*d + a.mac)            * E + (d*b.mac   + c.mac)   * G
// The group elements are identical provided E lies in the prime-order subgroup (scalars are combined mod r before
// multiplying); every honestly generated share is a multiple of the generator, as in the reference's tests
// (lib.rs:48-54 random_point = generator * random scalar).
#pragma once
#include "curve.cuh"

namespace ark {

// Every gate that multiplies a variable point takes the window-table store as a workspace (`Tab`, curve.cuh): the kernels
// pass their scratch record, the overloads without it use a per-thread array (host emulation, tests).

// out = s * P
template <class C, class Tab>
ARK_D void pt_mul_elem(Tab& tab, typename C::Pt& out, const fe8& s, const typename C::Pt& P) {
  uint32_t k[8];
  build_table<C>(tab, P);
  finish_tables<C>(tab, 1);
  scalar_to_plain<typename C::R>(k, s);
  C::set_identity(out);
  var_mul<C>(out, tab, k);
}
template <class C>
ARK_D void pt_mul_elem(typename C::Pt& out, const fe8& s, const typename C::Pt& P) {
  LocalTabStore<C> store;
  LocalTab<C> tab = store.view();
  pt_mul_elem<C>(tab, out, s, P);
}

// tables of P_p = 2^(4 W p) P, p < C::kSplitParts, for the split two-pass multiplications (curve.cuh)
template <class C, class Tab>
ARK_D void build_split_tables(const Tab& tab, const typename C::Pt& P) {
  typename C::Pt Q = P;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int h = 0; h < C::kSplitParts; h++) {  // one inlined copy of the table construction
    if (h) shift_windows<C>(Q, split_shift_windows<C>());
    Tab part = tab.sub(h);
    build_table<C>(part, Q);
  }
  finish_tables<C>(tab, C::kSplitParts);
}

// (out0, out1) = (s0 * P, s1 * P): the two passes share the tables of P and 2^s P
template <class C, class Tab>
ARK_D void pt_mul2_elem(Tab& tab, typename C::Pt& out0, typename C::Pt& out1, const fe8& s0, const fe8& s1, const typename C::Pt& P) {
  uint32_t k[8];
  build_split_tables<C>(tab, P);
  scalar_to_plain<typename C::R>(k, s0);
  C::set_identity(out0);
  var_mul_split<C>(out0, tab, k);
  scalar_to_plain<typename C::R>(k, s1);
  C::set_identity(out1);
  var_mul_split<C>(out1, tab, k);
}
template <class C>
ARK_D void pt_mul2_elem(typename C::Pt& out0, typename C::Pt& out1, const fe8& s0, const fe8& s1, const typename C::Pt& P) {
  LocalTabStore<C, kMaxSplitParts> store;
  LocalTab<C> tab = store.view();
  pt_mul2_elem<C>(tab, out0, out1, s0, s1, P);
}

// out = s * G
template <class C>
ARK_D void pt_mul_gen_elem(typename C::Pt& out, const fe8& s, const typename C::Aff* gtab) {
  uint32_t k[8];
  scalar_to_plain<typename C::R>(k, s);
  C::set_identity(out);
  fix_mul_acc<C>(out, gtab, k);
}

// PointShare::add_public / sub_public
template <class C, class Tab>
ARK_D void pt_share_add_public_elem(Tab& tab, typename C::Pt& out_s, typename C::Pt& out_m, int party, bool sub, const fe8& key,
                                    const typename C::Pt& a_s, const typename C::Pt& a_m, const typename C::Pt& pub) {
  typename C::Pt P = pub;
  if (sub) C::neg(P);                                   // curve/share.rs:63-65: add_public(-rhs)
  out_s = a_s;
  if (party == 0) C::add(out_s, P);
  typename C::Pt kp;
  pt_mul_elem<C>(tab, kp, key, P);
  out_m = a_m;
  C::add(out_m, kp);
}
template <class C>
ARK_D void pt_share_add_public_elem(typename C::Pt& out_s, typename C::Pt& out_m, int party, bool sub, const fe8& key,
                                    const typename C::Pt& a_s, const typename C::Pt& a_m, const typename C::Pt& pub) {
  LocalTabStore<C> store;
  LocalTab<C> tab = store.view();
  pt_share_add_public_elem<C>(tab, out_s, out_m, party, sub, key, a_s, a_m, pub);
}

// mac_key * opened - mac  (authenticated_curve.rs:160-175 `mac_check_value`)
template <class C, class Tab>
ARK_D void pt_mac_check_elem(Tab& tab, typename C::Pt& out, const fe8& key, const typename C::Pt& opened, const typename C::Pt& mac) {
  pt_mul_elem<C>(tab, out, key, opened);
  typename C::Pt m = mac;
  C::neg(m);
  C::add(out, m);
}
template <class C>
ARK_D void pt_mac_check_elem(typename C::Pt& out, const fe8& key, const typename C::Pt& opened, const typename C::Pt& mac) {
  LocalTabStore<C> store;
  LocalTab<C> tab = store.view();
  pt_mac_check_elem<C>(tab, out, key, opened, mac);
}

// What arkworks' point deserialisation enforces on values from the peer (curve.rs:105-135: on the curve AND in the prime-order
// subgroup) and what the regrouped recombination relies on for E_peer (scalars are combined mod r before multiplying).
template <class C, class Tab>
ARK_D bool pt_valid_elem(Tab& tab, const typename C::Pt& P) {
  if (!C::on_curve(P)) return false;
  if constexpr (C::kNeedsSubgroupCheck) {
    // [r]P with the UNREDUCED group order as the scalar (var_mul only needs k < 2^254)
    using R = typename C::R;
    const uint32_t k[8] = {R::P0, R::P1, R::P2, R::P3, R::P4, R::P5, R::P6, R::P7};
    typename C::Pt acc;
    build_table<C>(tab, P);
    finish_tables<C>(tab, 1);
    C::set_identity(acc);
    var_mul<C>(acc, tab, k);
    return C::is_identity(acc);
  }
  return true;
}
template <class C>
ARK_D bool pt_valid_elem(const typename C::Pt& P) {
  LocalTabStore<C> store;
  LocalTab<C> tab = store.view();
  return pt_valid_elem<C>(tab, P);
}

template <class C>
ARK_D void pt_beaver_mask_elem(fe8& d_mine, typename C::Pt& E_mine, const fe8& x_s, const fe8& a_s, const fe8& b_s,
                               const typename C::Pt& P_s, const typename C::Aff* gtab) {
  Fp<typename C::R>::sub(d_mine, x_s, a_s);
  typename C::Pt bG;
  pt_mul_gen_elem<C>(bG, b_s, gtab);
  C::neg(bG);
  E_mine = P_s;
  C::add(E_mine, bG);
}

// Writes the opened d and E, and the two result points through `emit(which, point)` (which = 0 share, 1 mac)
// so that a kernel can store each as soon as its pass finishes.
template <class C, bool DUAL, class Tab, class Emit>
ARK_D void pt_beaver_recombine_elem(Tab& tab, fe8& d, typename C::Pt& E, int party, const fe8& key, const fe8& d_mine, const fe8& d_peer,
                                    const typename C::Pt& E_mine, const typename C::Pt& E_peer, const fe8& a_s, const fe8& a_m,
                                    const fe8& b_s, const fe8& b_m, const fe8& c_s, const fe8& c_m,
                                    const typename C::Aff* gtab, Emit emit) {
  using FR = Fp<typename C::R>;
  FR::add(d, d_mine, d_peer);
  E = E_mine;
  C::add(E, E_peer);
  build_split_tables<C>(tab, E);
  fe8 sv[2], tv[2], t;
  if (party == 0) FR::add(sv[0], a_s, d); else sv[0] = a_s;
  FR::mul(t, d, b_s);
  FR::add(tv[0], t, c_s);
  FR::mul(t, key, d);
  FR::add(sv[1], t, a_m);
  FR::mul(t, d, b_m);
  FR::add(tv[1], t, c_m);
  if constexpr (DUAL) {  // both variable-base passes in lock-step (two independent chains per thread), then the fixed-base parts
    uint32_t k0[8], k1[8];
    typename C::Pt acc0, acc1;
    C::set_identity(acc0);
    C::set_identity(acc1);
    scalar_to_plain<typename C::R>(k0, sv[0]);
    scalar_to_plain<typename C::R>(k1, sv[1]);
    var_mul2_split<C>(acc0, acc1, tab, k0, k1);
    scalar_to_plain<typename C::R>(k0, tv[0]);
    fix_mul_acc<C>(acc0, gtab, k0);
    emit(0, acc0);
    scalar_to_plain<typename C::R>(k1, tv[1]);
    fix_mul_acc<C>(acc1, gtab, k1);
    emit(1, acc1);
  } else {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int which = 0; which < 2; which++) {
      uint32_t k[8];
      typename C::Pt acc;
      C::set_identity(acc);
      scalar_to_plain<typename C::R>(k, sv[which]);
      var_mul_split<C>(acc, tab, k);
      scalar_to_plain<typename C::R>(k, tv[which]);
      fix_mul_acc<C>(acc, gtab, k);
      emit(which, acc);
    }
  }
}

template <class C, bool DUAL, class Emit>
ARK_D void pt_beaver_recombine_elem(fe8& d, typename C::Pt& E, int party, const fe8& key, const fe8& d_mine, const fe8& d_peer,
                                    const typename C::Pt& E_mine, const typename C::Pt& E_peer, const fe8& a_s, const fe8& a_m,
                                    const fe8& b_s, const fe8& b_m, const fe8& c_s, const fe8& c_m,
                                    const typename C::Aff* gtab, Emit emit) {
  LocalTabStore<C, kMaxSplitParts> store;
  LocalTab<C> tab = store.view();
  pt_beaver_recombine_elem<C, DUAL>(tab, d, E, party, key, d_mine, d_peer, E_mine, E_peer, a_s, a_m, b_s, b_m, c_s, c_m, gtab, emit);
}

}  // namespace ark
