// Host side of the constant-multiplier tables (CTab, fp256.cuh): k[i] = s * 2^(32 i + 96 - 256) mod p for i < 4 and
// s * 2^(32 i + 64 - 256) mod p for i >= 4, canonical, from the 4 x u64 Montgomery image of s that the reference keeps
// in memory (e.g. the MAC key share, fabric.rs `mac_key`; share.rs:76 multiplies every public value by it).
// Plain C++ (no device code): also compiled by the host-emulation harness.
#pragma once
#include <stdint.h>

#include "fp256.cuh"

namespace ark {

namespace ctab_detail {
struct U256 { uint64_t w[4]; };

template <class F>
inline U256 modulus() {
  const uint32_t p[8] = {F::P0, F::P1, F::P2, F::P3, F::P4, F::P5, F::P6, F::P7};
  U256 r;
  for (int j = 0; j < 4; j++) r.w[j] = (uint64_t)p[2 * j] | ((uint64_t)p[2 * j + 1] << 32);
  return r;
}
inline bool geq(const U256& a, const U256& b) {
  for (int j = 3; j >= 0; j--)
    if (a.w[j] != b.w[j]) return a.w[j] > b.w[j];
  return true;
}
inline void add_in(U256& a, const U256& b) {
  unsigned __int128 c = 0;
  for (int j = 0; j < 4; j++) { c += (unsigned __int128)a.w[j] + b.w[j]; a.w[j] = (uint64_t)c; c >>= 64; }
}
inline void sub_in(U256& a, const U256& b) {
  unsigned __int128 br = 0;
  for (int j = 0; j < 4; j++) {
    unsigned __int128 t = (unsigned __int128)a.w[j] - b.w[j] - br;
    a.w[j] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
}
// x/2 mod p (p odd, x < p < 2^255)
inline void halve(U256& x, const U256& p) {
  if (x.w[0] & 1) add_in(x, p);
  for (int j = 0; j < 3; j++) x.w[j] = (x.w[j] >> 1) | (x.w[j + 1] << 63);
  x.w[3] >>= 1;
}
// 2x mod p (x < p < 2^255)
inline void dbl(U256& x, const U256& p) {
  for (int j = 3; j > 0; j--) x.w[j] = (x.w[j] << 1) | (x.w[j - 1] >> 63);
  x.w[0] <<= 1;
  if (geq(x, p)) sub_in(x, p);
}
}  // namespace ctab_detail

// s_mont: 4 x u64 LE canonical (< p).  Values >= p are reduced first (peer-supplied keys never reach this path).
template <class F>
inline void ctab_build(CTab& T, const uint64_t* s_mont) {
  using namespace ctab_detail;
  const U256 p = modulus<F>();
  U256 x;
  for (int j = 0; j < 4; j++) x.w[j] = s_mont[j];
  while (geq(x, p)) sub_in(x, p);
  // exponent of two applied to s for row i: 32 i - 160 (i < 4), 32 i - 192 (i >= 4):
  //   i: 0 -> -160, 1 -> -128, 2 -> -96, 3 -> -64, 4 -> -64, 5 -> -32, 6 -> 0, 7 -> +32
  U256 v[8];
  v[6] = x;
  v[7] = x;
  for (int k = 0; k < 32; k++) dbl(v[7], p);
  v[5] = v[6];
  for (int k = 0; k < 32; k++) halve(v[5], p);
  v[4] = v[5];
  for (int k = 0; k < 32; k++) halve(v[4], p);
  v[3] = v[4];
  for (int i = 2; i >= 0; i--) {
    v[i] = v[i + 1];
    for (int k = 0; k < 32; k++) halve(v[i], p);
  }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 4; j++) {
      T.k[i][2 * j] = (uint32_t)v[i].w[j];
      T.k[i][2 * j + 1] = (uint32_t)(v[i].w[j] >> 32);
    }
}

}  // namespace ark
