// Base field of Curve25519, p = 2^255 - 19, with special-form reduction (no Montgomery form inside the kernels).
//
// The reference keeps coordinates as Montgomery residues x*2^256 mod p (ark-ff Fp256<MontBackend>, like every arkworks
// field); the library's boundary keeps that image.  Inside a point kernel a coordinate is converted once on load
// (x = image * 38^-1, because 2^256 = 38 mod p) and once on store (image = x * 38), and in between all arithmetic works
// on plain residues modulo 2p = 2^256 - 38, kept only "loosely" reduced: any 256-bit value is a valid representative.
//   * a product is a 512-bit schoolbook product (64 IMAD.WIDE in the same even/odd carry chains as the Montgomery code)
//     folded with 2^256 = 38: 8 more multiply-adds, against 64 for a Montgomery reduction;
//   * add / sub fold the carry / borrow out of bit 256 back in as +-38.
// Canonical (< p) values are produced only where they are observable: stores, comparisons.
// Dual-target like fp256.cuh (host emulation for the bit-exactness tests).
#pragma once
#include "fp256.cuh"

// multiplications stay out of line in the host-emulation build (keeps it small and fast to compile)
#ifndef ARK_FQ_MUL
#if defined(__CUDACC__)
#define ARK_FQ_MUL ARK_DM
#else
#define ARK_FQ_MUL __attribute__((noinline))
#endif
#endif

namespace ark {

struct F25519 {
  ARK_DM static void zero(fe8& r) { ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = 0; }
  ARK_DM static void one(fe8& r) { zero(r); r.v[0] = 1; }

  // r = a + b (mod 2p)
  ARK_DM static void add(fe8& r, const fe8& a, const fe8& b) {
    uint32_t t[8];
    t[0] = add_cc(a.v[0], b.v[0]);
    ARK_UNROLL for (int j = 1; j < 8; j++) t[j] = addc_cc(a.v[j], b.v[j]);
    const uint32_t c = addc(0u, 0u);                 // carry out of bit 256: worth 38
    t[0] = add_cc(t[0], 38u * c);
    ARK_UNROLL for (int j = 1; j < 8; j++) t[j] = addc_cc(t[j], 0u);
    const uint32_t c2 = addc(0u, 0u);                // only if the sum wrapped again: then t < 38 and the add below cannot carry
    t[0] += 38u * c2;
    ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = t[j];
  }
  // r = a - b (mod 2p)
  ARK_DM static void sub(fe8& r, const fe8& a, const fe8& b) {
    uint32_t t[8];
    t[0] = sub_cc(a.v[0], b.v[0]);
    ARK_UNROLL for (int j = 1; j < 8; j++) t[j] = subc_cc(a.v[j], b.v[j]);
    const uint32_t bw = subc(0u, 0u) & 1u;           // borrow out of bit 256: worth -38
    t[0] = sub_cc(t[0], 38u * bw);
    ARK_UNROLL for (int j = 1; j < 8; j++) t[j] = subc_cc(t[j], 0u);
    const uint32_t bw2 = subc(0u, 0u) & 1u;          // only if it wrapped again: then t >= 2^256 - 38 and the sub below cannot borrow
    t[0] -= 38u * bw2;
    ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = t[j];
  }
  ARK_DM static void dbl(fe8& r, const fe8& a) { add(r, a, a); }
  ARK_DM static void neg(fe8& r, const fe8& a) {
    fe8 z;
    zero(z);
    sub(r, z, a);
  }

  // fold a 9-word value lo + top*2^256 (top < 2^32 / 38) into 8 words
  ARK_DM static void fold_top(fe8& r, uint32_t* lo, uint32_t top) {
    lo[0] = add_cc(lo[0], 38u * top);
    ARK_UNROLL for (int j = 1; j < 8; j++) lo[j] = addc_cc(lo[j], 0u);
    const uint32_t c2 = addc(0u, 0u);
    lo[0] += 38u * c2;
    ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = lo[j];
  }

  // r = a * b (mod 2p): 512-bit product, then 2^256 = 38
  ARK_FQ_MUL static void mul(fe8& r, const fe8& a, const fe8& b) {
    MontAcc t;
    acc_zero(t);
    uint32_t lo[8];
    ARK_UNROLL for (int i = 0; i < 8; i++) {
      if (i == 0) acc_row(t, a.v, b.v[0]); else acc_row_first(t, a.v, b.v[i]);
      // the word of weight 2^0 is final: emit it and divide the accumulator by 2^32 (relabelling, as acc_reduce_shift does)
      lo[i] = t.E[0];
      const uint32_t fold = t.E[1];
      uint32_t nO[8];
      ARK_UNROLL for (int j = 0; j < 7; j++) nO[j] = t.E[j + 2];
      nO[7] = 0;
      ARK_UNROLL for (int j = 0; j < 8; j++) t.E[j] = t.O[j];
      t.E[8] = 0;
      ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = nO[j];
      t.fold = fold;
    }
    fe8 hi;
    acc_collapse(hi, t);                              // high half of the product, < 2^256
    fold512(r, lo, hi.v);
  }

  // r = lo + 38 * hi (mod 2p), lo and hi 8 words each: one more multiply-accumulate row on an accumulator preloaded with lo
  ARK_DM static void fold512(fe8& r, const uint32_t* lo, const uint32_t* hi) {
    MontAcc u;
    ARK_UNROLL for (int j = 0; j < 8; j++) { u.E[j] = lo[j]; u.O[j] = 0; }
    u.E[8] = 0;
    u.fold = 0;
    acc_row(u, hi, 38u);
    uint32_t s[8];
    s[0] = u.E[0];
    s[1] = add_cc(u.E[1], u.O[0]);
    ARK_UNROLL for (int j = 2; j < 8; j++) s[j] = addc_cc(u.E[j], u.O[j - 1]);
    const uint32_t top = addc(u.E[8], u.O[7]);        // < 39
    fold_top(r, s, top);
  }

  // r = a^2 (mod 2p): dedicated 512-bit square (36 wide multiply-adds instead of 64), then the same fold
  ARK_FQ_MUL static void sqr(fe8& r, const fe8& a) {
    uint32_t t[16];
    sqr512(t, a.v);
    fold512(r, t, t + 8);
  }

  // canonical representative (< p)
  ARK_DM static void canon(fe8& r, const fe8& a) {
    uint32_t t[8];
    const uint32_t top = a.v[7] >> 31;                // bit 255: worth 19
    t[0] = add_cc(a.v[0], 19u * top);
    ARK_UNROLL for (int j = 1; j < 7; j++) t[j] = addc_cc(a.v[j], 0u);
    t[7] = addc(a.v[7] & 0x7fffffffu, 0u);            // < 2^255 + 19
    // subtract p = 2^255 - 19 if t >= p  <=>  t + 19 >= 2^255
    uint32_t u[8];
    u[0] = add_cc(t[0], 19u);
    ARK_UNROLL for (int j = 1; j < 7; j++) u[j] = addc_cc(t[j], 0u);
    u[7] = addc(t[7], 0u);
    const bool ge = (u[7] >> 31) != 0;                // t >= p; then t - p = u - 2^255
    u[7] &= 0x7fffffffu;
    ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = ge ? u[j] : t[j];
  }
  ARK_DM static bool is_zero(const fe8& a) {
    fe8 c;
    canon(c, a);
    uint32_t nz = 0;
    ARK_UNROLL for (int j = 0; j < 8; j++) nz |= c.v[j];
    return nz == 0;
  }
  ARK_DM static bool eq(const fe8& a, const fe8& b) {
    fe8 d;
    sub(d, a, b);
    return is_zero(d);
  }

  // reference memory image (canonical Montgomery residue x*2^256 mod p) <-> internal plain residue
  ARK_DM static void from_image(fe8& r, const fe8& img) {
    fe8 k;  // 38^-1 mod p
    k.v[0] = 0x9435e50au; k.v[1] = 0x435e50d7u; k.v[2] = 0x35e50d79u; k.v[3] = 0x5e50d794u;
    k.v[4] = 0xe50d7943u; k.v[5] = 0x50d79435u; k.v[6] = 0x0d79435eu; k.v[7] = 0x179435e5u;
    mul(r, img, k);
  }
  ARK_DM static void to_image(fe8& r, const fe8& x) {
    fe8 k, t;
    zero(k);
    k.v[0] = 38u;
    mul(t, x, k);
    canon(r, t);
  }

  // a^-1 by division steps on the canonical residue (Fp::inv_safegcd; a Fermat chain for 2^255 - 21 is 255 squarings and 250
  // multiplications); inv(0) = 0
  ARK_DM static void inv(fe8& r, const fe8& a) {
    fe8 c;
    canon(c, a);
    if (is_zero(c)) { zero(r); return; }
    Fp<Curve25519Fq>::inv_safegcd(r, c);
  }
};

}  // namespace ark
