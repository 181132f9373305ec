// Per-element gate arithmetic of the authenticated Beaver multiplication and the linear SPDZ gates.
// One thread evaluates one gate; these functions are shared by every kernel skeleton in kernels.cu
// and by the host-emulation harness (tests/host_emu/emu.cpp).
//
// Reference behaviour restated (paths under /root/reference/online-phase/src/algebra/scalar):
//   mask       d_mine = x.share - a.share, e_mine = y.share - b.share   authenticated_scalar.rs:863-867 (+ :141-145: only the
//              share component is ever sent; the MAC halves the reference's batch_sub also computes are dead values)
//   open       d = d_mine + d_peer                                        authenticated_scalar.rs:161-171
//   recombine  [xy] = de + d[b] + e[a] + [c]                              authenticated_scalar.rs:871-878, single-gate form :826-840
//              share: only party 0 adds the public de; mac += mac_key*de   share.rs:74-77
// All results are canonical Montgomery residues, hence bit-identical to the reference's.
#pragma once
#include "fp256.cuh"

namespace ark {

template <class F>
ARK_D void beaver_mask_elem(fe8& d_mine, fe8& e_mine, const fe8& x_s, const fe8& y_s, const fe8& a_s, const fe8& b_s) {
  Fp<F>::sub(d_mine, x_s, a_s);
  Fp<F>::sub(e_mine, y_s, b_s);
}

// Fused recombination.  6 modular multiplications in the reference (d*e, d*b.share, d*b.mac, e*a.share,
// e*a.mac, key*de) become 5 products and 2 + 3/8 Montgomery reductions:
//   share = REDC(d*(b.share [+ e on party 0]) + e*a.share) + c.share
//   mac   = REDC(d*(b.mac + key*e)            + e*a.mac)   + c.mac
// key*e uses the MAC key's constant-multiplier table (fp256.cuh CTab: three reduction steps instead of eight).
// Operand sums stay unreduced (< 3p + 5p/2^32 < 2^256); F::kLazy2 makes one conditional subtraction enough.
template <class F>
ARK_D void beaver_recombine_elem(fe8& out_s, fe8& out_m, fe8& d, fe8& e, int party, const CTab& key,
                                 const fe8& d_mine, const fe8& e_mine, const fe8& d_peer, const fe8& e_peer,
                                 const fe8& a_s, const fe8& a_m, const fe8& b_s, const fe8& b_m,
                                 const fe8& c_s, const fe8& c_m) {
  static_assert(F::kLazy2, "fused recombination needs 4p < 2^256");
  Fp<F>::add(d, d_mine, d_peer);
  Fp<F>::add(e, e_mine, e_peer);

  fe8 x;
  if (party == 0) Fp<F>::add_raw(x, b_s, e); else x = b_s;  // < 2p
  fe8 s;
  Fp<F>::mul2_lazy(s, d, x, e, a_s);                          // < (2p^2 + p^2)/R + p < 2p
  Fp<F>::csub_p(s);
  Fp<F>::add(out_s, s, c_s);

  fe8 ke;
  Fp<F>::mul_ctab_lazy(ke, key, e);                           // < p + 5p/2^32
  fe8 y;
  Fp<F>::add_raw(y, b_m, ke);                                 // < 2p + 5p/2^32
  fe8 m;
  Fp<F>::mul2_lazy(m, d, y, e, a_m);                          // < (2.01p^2 + p^2)/R + p < 2p
  Fp<F>::csub_p(m);
  Fp<F>::add(out_m, m, c_m);
}

// ---- linear gates on ScalarShare (share.rs:74-131) ----
template <class F>
ARK_D void share_add_public_elem(fe8& out_s, fe8& out_m, int party, const CTab& key, const fe8& s, const fe8& m, const fe8& v) {
  if (party == 0) Fp<F>::add(out_s, s, v); else out_s = s;   // share.rs:75
  fe8 kv;
  Fp<F>::mul_ctab(kv, key, v);
  Fp<F>::add(out_m, m, kv);                                   // share.rs:76
}
template <class F>
ARK_D void share_sub_public_elem(fe8& out_s, fe8& out_m, int party, const CTab& key, const fe8& s, const fe8& m, const fe8& v) {
  fe8 nv;
  Fp<F>::neg(nv, v);                                          // share.rs:80-82: add_public(-rhs)
  share_add_public_elem<F>(out_s, out_m, party, key, s, m, nv);
}
// mac_key * value - share.mac   (authenticated_scalar.rs:299-311)
template <class F>
ARK_D void mac_check_elem(fe8& out, const CTab& key, const fe8& opened, const fe8& mac) {
  fe8 kv;
  Fp<F>::mul_ctab(kv, key, opened);
  Fp<F>::sub(out, kv, mac);
}

}  // namespace ark
