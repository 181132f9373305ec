// 256-bit prime-field arithmetic for sm_100a, one field element per thread.
//
// Representation contract (what the reference keeps in memory): canonical Montgomery residues
// a*2^256 mod p in 4 little-endian u64 limbs == 8 little-endian u32 limbs
// (ark-ff 0.4 Fp256<MontBackend<_,4>> behind Scalar<C>, /root/reference/online-phase/src/algebra/scalar/scalar.rs:46).
//
// The multiplier is a carry-chain CIOS Montgomery product on 32-bit limbs. Every
// (mad.lo.cc, madc.hi.cc) pair below is fused by ptxas into ONE IMAD.WIDE.U32[.X]
// (verified with cuobjdump), so an 8x8-limb row costs 8 wide multiply-adds. Products whose
// low word lands on an even limb position accumulate in E, odd positions in O
// (value = E + O*2^32); a 32-bit right shift then just swaps the roles of the two arrays,
// which keeps every 64-bit accumulator pair register-aligned with no moves.
//
// `mont_mul2` accumulates TWO products a*x + b*y and reduces ONCE. The Beaver recombination
// needs 6 field multiplications in the reference's unfused form; with lazy accumulation it is
// 5 products and 3 reductions (see beaver.cuh).
//
// This header is dual-target: under nvcc the carry primitives are inline PTX; under a plain
// host compiler they are emulated with an explicit carry flag so that the exact instruction
// sequence can be verified bit-for-bit on a CPU-only box (tests/test_host_emu.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ARK_D __device__ __forceinline__
#define ARK_DM __device__ __forceinline__
#define ARK_HDM __host__ __device__ __forceinline__
#define ARK_UNROLL _Pragma("unroll")
#else
#define ARK_D static inline __attribute__((always_inline))
#define ARK_DM inline __attribute__((always_inline))
#define ARK_HDM inline __attribute__((always_inline))
#define ARK_UNROLL
#endif

namespace ark {

// ----------------------------------------------------------------------------------------------
// Carry-flag primitives
// ----------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
ARK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
ARK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ARK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ARK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
ARK_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
#define ARK_EMU_EXPECT_NO_CARRY() do { } while (0)
#else
namespace emu { static thread_local uint32_t CF = 0; static thread_local uint64_t violations = 0; }
#define ARK_EMU_EXPECT_NO_CARRY() do { if (::ark::emu::CF) ::ark::emu::violations++; } while (0)
ARK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; emu::CF = (uint32_t)(t >> 32); return (uint32_t)t; }
ARK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + emu::CF; emu::CF = (uint32_t)(t >> 32); return (uint32_t)t; }
ARK_D uint32_t addc(uint32_t a, uint32_t b) { return a + b + emu::CF; }
ARK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; emu::CF = (uint32_t)(t >> 63); return (uint32_t)t; }
ARK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - emu::CF; emu::CF = (uint32_t)(t >> 63); return (uint32_t)t; }
ARK_D uint32_t subc(uint32_t a, uint32_t b) { return a - b - emu::CF; }
ARK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; emu::CF = (uint32_t)(t >> 32); return (uint32_t)t; }
ARK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + emu::CF; emu::CF = (uint32_t)(t >> 32); return (uint32_t)t; }
ARK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + emu::CF; emu::CF = (uint32_t)(t >> 32); return (uint32_t)t; }
ARK_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
#endif

// ----------------------------------------------------------------------------------------------
// Field parameter packs (32-bit LE limbs).  kLazy2: 4p < 2^256 (p < 2^254), which the two-product
// accumulation and the lazy (unreduced) operand sums in beaver.cuh rely on: a sum of two products of a
// canonical value with a value < 3p reduces to < (4p/R + 1) p < 2p, so ONE conditional subtraction
// yields the canonical result.
// ----------------------------------------------------------------------------------------------
struct Bn254Fr {
  static constexpr int kId = 0;
  static constexpr bool kLazy2 = true;
  static constexpr uint32_t P0 = 0xf0000001u, P1 = 0x43e1f593u, P2 = 0x79b97091u, P3 = 0x2833e848u,
                            P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xefffffffu;  // -p^-1 mod 2^32
  // R mod p and R^2 mod p (Montgomery one / conversion constant)
  static constexpr uint32_t R_0 = 0x4ffffffbu, R_1 = 0xac96341cu, R_2 = 0x9f60cd29u, R_3 = 0x36fc7695u, R_4 = 0x7879462eu, R_5 = 0x666ea36fu, R_6 = 0x9a07df2fu, R_7 = 0x0e0a77c1u;
  static constexpr uint32_t R2_0 = 0xae216da7u, R2_1 = 0x1bb8e645u, R2_2 = 0xe35c59e3u, R2_3 = 0x53fe3ab1u, R2_4 = 0x53bb8085u, R2_5 = 0x8c49833du, R2_6 = 0x7f4e44a5u, R2_7 = 0x0216d0b1u;
  static constexpr int kBits = 254;
};
struct Curve25519Fr {
  static constexpr int kId = 1;
  static constexpr bool kLazy2 = true;
  static constexpr uint32_t P0 = 0x5cf5d3edu, P1 = 0x5812631au, P2 = 0xa2f79cd6u, P3 = 0x14def9deu,
                            P4 = 0u, P5 = 0u, P6 = 0u, P7 = 0x10000000u;
  static constexpr uint32_t INV = 0x12547e1bu;
  static constexpr uint32_t R_0 = 0x8d98951du, R_1 = 0xd6ec3174u, R_2 = 0x737dcf70u, R_3 = 0xc6ef5bf4u, R_4 = 0xfffffffeu, R_5 = 0xffffffffu, R_6 = 0xffffffffu, R_7 = 0x0fffffffu;
  static constexpr uint32_t R2_0 = 0x449c0f01u, R2_1 = 0xa40611e3u, R2_2 = 0x68859347u, R2_3 = 0xd00e1ba7u, R2_4 = 0x17f5be65u, R2_5 = 0xceec73d2u, R2_6 = 0x7c309a3du, R2_7 = 0x0399411bu;
  static constexpr int kBits = 253;
};
struct Bn254Fq {
  static constexpr int kId = 2;
  static constexpr bool kLazy2 = true;
  static constexpr uint32_t P0 = 0xd87cfd47u, P1 = 0x3c208c16u, P2 = 0x6871ca8du, P3 = 0x97816a91u,
                            P4 = 0x8181585du, P5 = 0xb85045b6u, P6 = 0xe131a029u, P7 = 0x30644e72u;
  static constexpr uint32_t INV = 0xe4866389u;
  static constexpr uint32_t R_0 = 0xc58f0d9du, R_1 = 0xd35d438du, R_2 = 0xf5c70b3du, R_3 = 0x0a78eb28u, R_4 = 0x7879462cu, R_5 = 0x666ea36fu, R_6 = 0x9a07df2fu, R_7 = 0x0e0a77c1u;
  static constexpr uint32_t R2_0 = 0x538afa89u, R2_1 = 0xf32cfc5bu, R2_2 = 0xd44501fbu, R2_3 = 0xb5e71911u, R2_4 = 0x0a417ff6u, R2_5 = 0x47ab1effu, R2_6 = 0xcab8351fu, R2_7 = 0x06d89f71u;
  static constexpr int kBits = 254;
};
struct Curve25519Fq {
  static constexpr int kId = 3;
  static constexpr bool kLazy2 = false;  // 2^255 - 19: 4p does not fit 256 bits
  static constexpr uint32_t P0 = 0xffffffedu, P1 = 0xffffffffu, P2 = 0xffffffffu, P3 = 0xffffffffu,
                            P4 = 0xffffffffu, P5 = 0xffffffffu, P6 = 0xffffffffu, P7 = 0x7fffffffu;
  static constexpr uint32_t INV = 0x286bca1bu;
  static constexpr uint32_t R_0 = 0x26u, R_1 = 0, R_2 = 0, R_3 = 0, R_4 = 0, R_5 = 0, R_6 = 0, R_7 = 0;
  static constexpr uint32_t R2_0 = 0x5a4u, R2_1 = 0, R2_2 = 0, R2_3 = 0, R2_4 = 0, R2_5 = 0, R2_6 = 0, R2_7 = 0;
  static constexpr int kBits = 255;
};


// A field element in registers: 8 x u32, little endian.
struct fe8 { uint32_t v[8]; };

// ----------------------------------------------------------------------------------------------
// Multiply-accumulate rows.  acc[0..7] += {a[0],a[2],a[4],a[6]} * w  (lo parts on even slots,
// hi parts on odd slots => 4 fused IMAD.WIDE.U32.X), returns with the chain's carry-out pending
// in the flag; the caller decides where it goes.
// ----------------------------------------------------------------------------------------------
template <bool CARRY_IN>
ARK_D void mad_row4(uint32_t* acc, uint32_t a0, uint32_t a2, uint32_t a4, uint32_t a6, uint32_t w) {
  acc[0] = CARRY_IN ? madc_lo_cc(a0, w, acc[0]) : mad_lo_cc(a0, w, acc[0]);
  acc[1] = madc_hi_cc(a0, w, acc[1]);
  acc[2] = madc_lo_cc(a2, w, acc[2]);
  acc[3] = madc_hi_cc(a2, w, acc[3]);
  acc[4] = madc_lo_cc(a4, w, acc[4]);
  acc[5] = madc_hi_cc(a4, w, acc[5]);
  acc[6] = madc_lo_cc(a6, w, acc[6]);
  acc[7] = madc_hi_cc(a6, w, acc[7]);
}

// Same, for a compile-time modulus limb quadruple: zero limbs cost an add-with-carry, not a multiply
// (Curve25519 Fr has three zero limbs).
template <uint32_t PJ>
ARK_D void mad_plimb(uint32_t& lo, uint32_t& hi, uint32_t m, bool first) {
  if (PJ == 0u) {
    lo = first ? add_cc(lo, 0u) : addc_cc(lo, 0u);
    hi = addc_cc(hi, 0u);
  } else {
    lo = first ? mad_lo_cc(PJ, m, lo) : madc_lo_cc(PJ, m, lo);
    hi = madc_hi_cc(PJ, m, hi);
  }
}

// Accumulator: value = sum E[j] 2^(32j) + 2^32 * sum O[j] 2^(32j).  Invariant:
// the value stays < 2^288 inside an iteration and < 2^256 after each shift, hence E has 9 words,
// O has 8 and O never carries out.
struct MontAcc {
  uint32_t E[9];
  uint32_t O[8];
  uint32_t fold;  // word of weight 2^0 left over by the previous shift, not yet added into E[0]
};

ARK_D void acc_zero(MontAcc& t) {
  ARK_UNROLL for (int j = 0; j < 9; j++) t.E[j] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = 0;
  t.fold = 0;
}

// t += a * w, first row after a shift (adds the folded word; its carry has weight 2^32 == O[0])
ARK_D void acc_row_first(MontAcc& t, const uint32_t* a, uint32_t w) {
  t.E[0] = add_cc(t.E[0], t.fold);
  mad_row4<true>(t.O, a[1], a[3], a[5], a[7], w);
  ARK_EMU_EXPECT_NO_CARRY();
  mad_row4<false>(t.E, a[0], a[2], a[4], a[6], w);
  t.E[8] = addc(t.E[8], 0u);
}
ARK_D void acc_row(MontAcc& t, const uint32_t* a, uint32_t w) {
  mad_row4<false>(t.O, a[1], a[3], a[5], a[7], w);
  ARK_EMU_EXPECT_NO_CARRY();
  mad_row4<false>(t.E, a[0], a[2], a[4], a[6], w);
  t.E[8] = addc(t.E[8], 0u);
}

// One Montgomery step: t = (t + m*p) / 2^32 with m chosen so the low word cancels.
template <class F>
ARK_D void acc_reduce_shift(MontAcc& t) {
  const uint32_t m = mul_lo(t.E[0], F::INV);
  mad_plimb<F::P1>(t.O[0], t.O[1], m, true);
  mad_plimb<F::P3>(t.O[2], t.O[3], m, false);
  mad_plimb<F::P5>(t.O[4], t.O[5], m, false);
  mad_plimb<F::P7>(t.O[6], t.O[7], m, false);
  ARK_EMU_EXPECT_NO_CARRY();
  mad_plimb<F::P0>(t.E[0], t.E[1], m, true);
  mad_plimb<F::P2>(t.E[2], t.E[3], m, false);
  mad_plimb<F::P4>(t.E[4], t.E[5], m, false);
  mad_plimb<F::P6>(t.E[6], t.E[7], m, false);
  t.E[8] = addc(t.E[8], 0u);
  // E[0] is now 0.  Divide by 2^32: O becomes the even array, E[2..8] the odd one, E[1] is folded later.
  const uint32_t fold = t.E[1];
  uint32_t nO[8];
  ARK_UNROLL for (int j = 0; j < 7; j++) nO[j] = t.E[j + 2];
  nO[7] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.E[j] = t.O[j];
  t.E[8] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = nO[j];
  t.fold = fold;
}

// r = E + fold + (O << 32); the caller guarantees the value is < 2^256.
ARK_D void acc_collapse(fe8& r, const MontAcc& t) {
  r.v[0] = add_cc(t.E[0], t.fold);
  ARK_UNROLL for (int j = 1; j < 8; j++) r.v[j] = addc_cc(t.E[j], t.O[j - 1]);
  ARK_EMU_EXPECT_NO_CARRY();
}

// ----------------------------------------------------------------------------------------------
// Multiplication by a batch-constant s (MAC key share, R^2, 1, ...): instead of a*s followed by a full 8-step
// Montgomery reduction (64 + 64 wide multiply-adds), the host precomputes the eight residues
//   k[i] = s * 2^(32 i + 96 - 256) mod p (i < 4),   k[i] = s * 2^(32 i + 64 - 256) mod p (i >= 4)
// (ctab.hpp) and the kernel evaluates sum_i a_i * k[i] with only THREE reduction steps: rows 0..3, shift, rows 4..7,
// shift, shift (64 + 24 wide multiply-adds).  Bounds (p < 2^254): after rows 0..3 the value is < 4 * 2^32 p < 2^288;
// each shift maps V to (V + m p)/2^32, so the result is < p + 5p/2^32: one conditional subtraction canonicalises it,
// and as a lazy operand it counts as "< 2p".  The table sits in the kernel's constant bank (a __grid_constant__
// parameter), so the k[i][j] operands cost no registers.
// ----------------------------------------------------------------------------------------------
struct CTab { uint32_t k[8][8]; };

// One Montgomery step on an accumulator whose folded word is still pending (no multiplication row in between): the
// carry of E[0] + fold enters the O chain, exactly as in acc_row_first.
template <class F>
ARK_D void acc_reduce_shift_pending(MontAcc& t) {
  t.E[0] = add_cc(t.E[0], t.fold);
  const uint32_t m = mul_lo(t.E[0], F::INV);
  mad_plimb<F::P1>(t.O[0], t.O[1], m, false);
  mad_plimb<F::P3>(t.O[2], t.O[3], m, false);
  mad_plimb<F::P5>(t.O[4], t.O[5], m, false);
  mad_plimb<F::P7>(t.O[6], t.O[7], m, false);
  ARK_EMU_EXPECT_NO_CARRY();
  mad_plimb<F::P0>(t.E[0], t.E[1], m, true);
  mad_plimb<F::P2>(t.E[2], t.E[3], m, false);
  mad_plimb<F::P4>(t.E[4], t.E[5], m, false);
  mad_plimb<F::P6>(t.E[6], t.E[7], m, false);
  t.E[8] = addc(t.E[8], 0u);
  const uint32_t fold = t.E[1];
  uint32_t nO[8];
  ARK_UNROLL for (int j = 0; j < 7; j++) nO[j] = t.E[j + 2];
  nO[7] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.E[j] = t.O[j];
  t.E[8] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = nO[j];
  t.fold = fold;
}

// ----------------------------------------------------------------------------------------------
// 512-bit square of an 8-limb value: r[0..15] = a^2.
// a^2 = D + 2S with D = sum a_i^2 2^(64 i) and S = sum_{i<j} a_i a_j 2^(32(i+j)).  S is a product with the multiplicand limbs
// j <= i of row i skipped (28 wide multiply-adds instead of 64, same even/odd carry chains and word-per-row emission as
// F25519::mul); skipped odd pairs still ripple the carry of the folded word.  Then one funnel-shift pass doubles S and one
// 16-word carry chain adds the 8 diagonal squares (8 more wide multiply-adds).
// ----------------------------------------------------------------------------------------------
// row I of S: t += a_I * (a_j for j > I), t already shifted I limbs (relative position of a_I * a_j is j)
template <int I>
ARK_D void sqr_row(MontAcc& t, const uint32_t* a) {
  const uint32_t w = a[I];
  // odd limbs j = 1,3,5,7 -> O pairs (0,1),(2,3),(4,5),(6,7); the carry of E[0] + fold enters at O[0]
  if (I > 0) t.E[0] = add_cc(t.E[0], t.fold);
  bool chain_open = I > 0;  // a carry may be pending from the fold
  ARK_UNROLL for (int k = 0; k < 4; k++) {
    const int j = 2 * k + 1;
    if (j > I) {
      if (chain_open) { t.O[2 * k] = madc_lo_cc(a[j], w, t.O[2 * k]); } else { t.O[2 * k] = mad_lo_cc(a[j], w, t.O[2 * k]); }
      t.O[2 * k + 1] = madc_hi_cc(a[j], w, t.O[2 * k + 1]);
      chain_open = true;
    } else if (chain_open) {
      t.O[2 * k] = addc_cc(t.O[2 * k], 0u);
      t.O[2 * k + 1] = addc_cc(t.O[2 * k + 1], 0u);
    }
  }
  ARK_EMU_EXPECT_NO_CARRY();
  // even limbs j = 2,4,6 (j = 0 is never > I) -> E pairs (2,3),(4,5),(6,7); carry out into E[8]
  bool e_open = false;
  ARK_UNROLL for (int k = 1; k < 4; k++) {
    const int j = 2 * k;
    if (j > I) {
      if (e_open) { t.E[2 * k] = madc_lo_cc(a[j], w, t.E[2 * k]); } else { t.E[2 * k] = mad_lo_cc(a[j], w, t.E[2 * k]); }
      t.E[2 * k + 1] = madc_hi_cc(a[j], w, t.E[2 * k + 1]);
      e_open = true;
    }
  }
  if (e_open) t.E[8] = addc(t.E[8], 0u);
}

ARK_D void acc_shift_emit(MontAcc& t, uint32_t& out) {
  out = t.E[0];
  const uint32_t fold = t.E[1];
  uint32_t nO[8];
  ARK_UNROLL for (int j = 0; j < 7; j++) nO[j] = t.E[j + 2];
  nO[7] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.E[j] = t.O[j];
  t.E[8] = 0;
  ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = nO[j];
  t.fold = fold;
}

ARK_D uint32_t shl1(uint32_t hi, uint32_t lo) { return (hi << 1) | (lo >> 31); }

ARK_D void sqr512(uint32_t* r, const uint32_t* a) {
  MontAcc t;
  acc_zero(t);
  uint32_t s[16];
  sqr_row<0>(t, a); acc_shift_emit(t, s[0]);
  sqr_row<1>(t, a); acc_shift_emit(t, s[1]);
  sqr_row<2>(t, a); acc_shift_emit(t, s[2]);
  sqr_row<3>(t, a); acc_shift_emit(t, s[3]);
  sqr_row<4>(t, a); acc_shift_emit(t, s[4]);
  sqr_row<5>(t, a); acc_shift_emit(t, s[5]);
  sqr_row<6>(t, a); acc_shift_emit(t, s[6]);
  sqr_row<7>(t, a); acc_shift_emit(t, s[7]);
  fe8 hi;
  acc_collapse(hi, t);
  ARK_UNROLL for (int j = 0; j < 8; j++) s[8 + j] = hi.v[j];
  // 2S
  uint32_t d[16];
  d[0] = s[0] << 1;
  ARK_UNROLL for (int j = 1; j < 16; j++) d[j] = shl1(s[j], s[j - 1]);
  // + D: one 16-word chain
  r[0] = mad_lo_cc(a[0], a[0], d[0]);
  r[1] = madc_hi_cc(a[0], a[0], d[1]);
  ARK_UNROLL for (int i = 1; i < 8; i++) {
    r[2 * i] = madc_lo_cc(a[i], a[i], d[2 * i]);
    r[2 * i + 1] = madc_hi_cc(a[i], a[i], d[2 * i + 1]);
  }
  ARK_EMU_EXPECT_NO_CARRY();
}

// ----------------------------------------------------------------------------------------------
// 512-bit product by one Karatsuba level: a = a0 + a1 2^128, b = b0 + b1 2^128,
//   a b = z0 + (z0 + z2 + (a0 - a1)(b1 - b0)) 2^128 + z2 2^256,   z0 = a0 b0, z2 = a1 b1.
// Three 4 x 4-limb products (48 wide multiply-adds instead of 64) and, as compiled, ~170 more add / select instructions per field
// product.  MEASURED SLOWER (profiles/r02v_summary.txt: NTT 2^20 249 -> 293 us, batch inversion 97 -> 126 us): the wide multiplier
// is 85-90 % busy in these kernels, but every instruction also costs an issue slot behind a dependent predecessor, and eight extra
// ALU instructions per multiply saved do not pay.  Kept as ARKMPC_MUL=kara for the A/B, and as the proof that the remaining
// distance to the multiplier roof is not recoverable by trading multiplies for additions (the same happened to the dedicated
// square in round 1, curve.cuh).
// ----------------------------------------------------------------------------------------------
// r[0..7] = a[0..3] * b[0..3]: the even / odd accumulator of MontAcc, four limbs wide, one word emitted per row
ARK_D void mul128(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t E[5] = {0, 0, 0, 0, 0}, O[4] = {0, 0, 0, 0}, fold = 0;
  ARK_UNROLL for (int i = 0; i < 4; i++) {
    const uint32_t w = b[i];
    if (i > 0) {
      E[0] = add_cc(E[0], fold);  // its carry has weight 2^32 == O[0]
      O[0] = madc_lo_cc(a[1], w, O[0]);
    } else {
      O[0] = mad_lo_cc(a[1], w, O[0]);
    }
    O[1] = madc_hi_cc(a[1], w, O[1]);
    O[2] = madc_lo_cc(a[3], w, O[2]);
    O[3] = madc_hi_cc(a[3], w, O[3]);
    ARK_EMU_EXPECT_NO_CARRY();
    E[0] = mad_lo_cc(a[0], w, E[0]);
    E[1] = madc_hi_cc(a[0], w, E[1]);
    E[2] = madc_lo_cc(a[2], w, E[2]);
    E[3] = madc_hi_cc(a[2], w, E[3]);
    E[4] = addc(E[4], 0u);
    // emit the low word and shift down by one limb
    r[i] = E[0];
    fold = E[1];
    const uint32_t n0 = E[2], n1 = E[3], n2 = E[4];
    ARK_UNROLL for (int j = 0; j < 4; j++) E[j] = O[j];
    E[4] = 0;
    O[0] = n0; O[1] = n1; O[2] = n2; O[3] = 0;
  }
  r[4] = add_cc(E[0], fold);
  r[5] = addc_cc(E[1], O[0]);
  r[6] = addc_cc(E[2], O[1]);
  r[7] = addc_cc(E[3], O[2]);
  ARK_EMU_EXPECT_NO_CARRY();
}

// d = |x - y| (four limbs); returns 0xffffffff if x < y
ARK_D uint32_t absdiff128(uint32_t* d, const uint32_t* x, const uint32_t* y) {
  d[0] = sub_cc(x[0], y[0]);
  d[1] = subc_cc(x[1], y[1]);
  d[2] = subc_cc(x[2], y[2]);
  d[3] = subc_cc(x[3], y[3]);
  const uint32_t neg = subc(0u, 0u);
  d[0] = add_cc(d[0] ^ neg, neg & 1u);
  d[1] = addc_cc(d[1] ^ neg, 0u);
  d[2] = addc_cc(d[2] ^ neg, 0u);
  d[3] = addc(d[3] ^ neg, 0u);
  return neg;
}

ARK_D void kara512(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  uint32_t da[4], db[4], m[8];
  const uint32_t sa = absdiff128(da, a, a + 4);      // a0 - a1
  const uint32_t sb = absdiff128(db, b + 4, b);      // b1 - b0
  mul128(T, a, b);                                   // z0
  mul128(T + 8, a + 4, b + 4);                       // z2
  mul128(m, da, db);
  const uint32_t neg = sa ^ sb;                      // sign of (a0 - a1)(b1 - b0)
  // mid = z0 + z2 +- m, nine limbs (0 <= mid = a0 b1 + a1 b0 < 2^257)
  uint32_t mid[9];
  mid[0] = add_cc(T[0], T[8]);
  ARK_UNROLL for (int j = 1; j < 8; j++) mid[j] = addc_cc(T[j], T[8 + j]);
  mid[8] = addc(0u, 0u);
  mid[0] = add_cc(mid[0], neg & 1u);                 // two's complement of m when negative: ~m + 1, sign-extended
  ARK_UNROLL for (int j = 1; j < 8; j++) mid[j] = addc_cc(mid[j], 0u);
  mid[8] = addc(mid[8], 0u);
  mid[0] = add_cc(mid[0], m[0] ^ neg);
  ARK_UNROLL for (int j = 1; j < 8; j++) mid[j] = addc_cc(mid[j], m[j] ^ neg);
  mid[8] = addc(mid[8], neg);
  // T += mid 2^128
  T[4] = add_cc(T[4], mid[0]);
  ARK_UNROLL for (int j = 1; j < 9; j++) T[4 + j] = addc_cc(T[4 + j], mid[j]);
  T[13] = addc_cc(T[13], 0u);
  T[14] = addc_cc(T[14], 0u);
  T[15] = addc_cc(T[15], 0u);
  ARK_EMU_EXPECT_NO_CARRY();
}

// ----------------------------------------------------------------------------------------------
// Field operations
// ----------------------------------------------------------------------------------------------
template <class F>
struct Fp {
  ARK_DM static void load_p(uint32_t* p) {
    p[0] = F::P0; p[1] = F::P1; p[2] = F::P2; p[3] = F::P3; p[4] = F::P4; p[5] = F::P5; p[6] = F::P6; p[7] = F::P7;
  }

  // r = a - p if a >= p else a.  (one conditional subtraction; a < 2p on entry gives canonical output)
  ARK_DM static void csub_p(fe8& a) {
    uint32_t p[8];
    load_p(p);
    uint32_t t[8];
    t[0] = sub_cc(a.v[0], p[0]);
    ARK_UNROLL for (int j = 1; j < 8; j++) t[j] = subc_cc(a.v[j], p[j]);
    const uint32_t borrow = subc(0u, 0u);  // 0xffffffff if a < p
    ARK_UNROLL for (int j = 0; j < 8; j++) a.v[j] = borrow ? a.v[j] : t[j];
  }

  // plain 256-bit add, no reduction (caller guarantees no overflow)
  ARK_DM static void add_raw(fe8& r, const fe8& a, const fe8& b) {
    r.v[0] = add_cc(a.v[0], b.v[0]);
    ARK_UNROLL for (int j = 1; j < 7; j++) r.v[j] = addc_cc(a.v[j], b.v[j]);
    r.v[7] = addc(a.v[7], b.v[7]);
  }

  // canonical add: inputs canonical, output canonical (ark-ff add_assign semantics)
  ARK_DM static void add(fe8& r, const fe8& a, const fe8& b) {
    if (F::kBits <= 255) {
      add_raw(r, a, b);  // a + b < 2p < 2^256
      csub_p(r);
    }
  }

  // canonical sub: r = a - b mod p
  ARK_DM static void sub(fe8& r, const fe8& a, const fe8& b) {
    uint32_t p[8];
    load_p(p);
    r.v[0] = sub_cc(a.v[0], b.v[0]);
    ARK_UNROLL for (int j = 1; j < 8; j++) r.v[j] = subc_cc(a.v[j], b.v[j]);
    const uint32_t borrow = subc(0u, 0u);  // all ones if a < b
    r.v[0] = add_cc(r.v[0], p[0] & borrow);
    ARK_UNROLL for (int j = 1; j < 7; j++) r.v[j] = addc_cc(r.v[j], p[j] & borrow);
    r.v[7] = addc(r.v[7], p[7] & borrow);
  }

  // canonical negation, neg(0) = 0 (ark-ff neg semantics)
  ARK_DM static void neg(fe8& r, const fe8& a) {
    uint32_t p[8];
    load_p(p);
    uint32_t nz = 0;
    ARK_UNROLL for (int j = 0; j < 8; j++) nz |= a.v[j];
    const uint32_t mask = nz ? 0xffffffffu : 0u;
    r.v[0] = sub_cc(p[0] & mask, a.v[0]);
    ARK_UNROLL for (int j = 1; j < 7; j++) r.v[j] = subc_cc(p[j] & mask, a.v[j]);
    r.v[7] = subc(p[7] & mask, a.v[7]);
  }

  // Lazy Montgomery product: r = a*x/R mod p, with r < a*x/R + p (NOT canonical).
  // Needs a < 2^256 arbitrary limbs, x arbitrary; bound of the running value: 2^32 (a + p) < 2^288.
  ARK_DM static void mul_lazy(fe8& r, const fe8& a, const fe8& x) {
    MontAcc t;
    acc_zero(t);
    ARK_UNROLL for (int i = 0; i < 8; i++) {
      if (i == 0) acc_row(t, a.v, x.v[0]); else acc_row_first(t, a.v, x.v[i]);
      acc_reduce_shift<F>(t);
    }
    acc_collapse(r, t);
  }

  // canonical product of canonical inputs
  ARK_DM static void mul(fe8& r, const fe8& a, const fe8& x) {
    mul_lazy(r, a, x);  // < p*p/R + p < 2p
    csub_p(r);
  }

  // Lazy two-product accumulation: r = (a*x + b*y)/R mod p, r < (a*x + b*y)/R + p.
  // Requires a + b + p < 2^256 (running value < 2^32 (a+b+p) < 2^288): a, b canonical and kLazy2.
  ARK_DM static void mul2_lazy(fe8& r, const fe8& a, const fe8& x, const fe8& b, const fe8& y) {
    MontAcc t;
    acc_zero(t);
    ARK_UNROLL for (int i = 0; i < 8; i++) {
      if (i == 0) acc_row(t, a.v, x.v[0]); else acc_row_first(t, a.v, x.v[i]);
      acc_row(t, b.v, y.v[i]);
      acc_reduce_shift<F>(t);
    }
    acc_collapse(r, t);
  }

  // Lazy product with a batch-constant multiplier given as its table (CTab above): r = a*s/R mod p, r < p + 5p/2^32.
  // a: any 256-bit value.
  ARK_DM static void mul_ctab_lazy(fe8& r, const CTab& T, const fe8& a) {
    static_assert(F::kBits <= 254, "four unreduced rows need 4 * 2^32 p < 2^288");
    MontAcc t;
    acc_zero(t);
    acc_row(t, T.k[0], a.v[0]);
    acc_row(t, T.k[1], a.v[1]);
    acc_row(t, T.k[2], a.v[2]);
    acc_row(t, T.k[3], a.v[3]);
    acc_reduce_shift<F>(t);
    acc_row_first(t, T.k[4], a.v[4]);
    acc_row(t, T.k[5], a.v[5]);
    acc_row(t, T.k[6], a.v[6]);
    acc_row(t, T.k[7], a.v[7]);
    acc_reduce_shift<F>(t);
    acc_reduce_shift_pending<F>(t);
    acc_collapse(r, t);
  }
  // canonical
  ARK_DM static void mul_ctab(fe8& r, const CTab& T, const fe8& a) {
    mul_ctab_lazy(r, T, a);
    csub_p(r);
  }

  // canonical square of a canonical input: dedicated 512-bit square (36 wide multiply-adds), then a word-serial Montgomery
  // reduction of the 16-word value (64): 100 against 128 for mul(a, a)
  ARK_DM static void sqr(fe8& r, const fe8& a) {
    uint32_t T[16];
    sqr512(T, a.v);
    MontAcc t;
    ARK_UNROLL for (int j = 0; j < 8; j++) { t.E[j] = T[j]; t.O[j] = 0; }
    t.E[8] = 0;
    t.fold = 0;
    ARK_UNROLL for (int i = 0; i < 8; i++) {
      if (i > 0) {  // the word folded out by the previous shift; its carry has weight 2^32 == O[0]
        t.E[0] = add_cc(t.E[0], t.fold);
        ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = addc_cc(t.O[j], 0u);
        ARK_EMU_EXPECT_NO_CARRY();
        t.fold = 0;
      }
      acc_reduce_shift<F>(t);
      t.E[7] = add_cc(t.E[7], T[8 + i]);  // next word of the high half enters at relative position 7
      t.E[8] = addc(t.E[8], 0u);
    }
    acc_collapse(r, t);  // < T/R + p < 2p
    csub_p(r);
  }

  // word-serial Montgomery reduction of a 16-word value T < p 2^256: r = T / R mod p, r < T / R + p (NOT canonical)
  ARK_DM static void redc512_lazy(fe8& r, const uint32_t* T) {
    MontAcc t;
    ARK_UNROLL for (int j = 0; j < 8; j++) { t.E[j] = T[j]; t.O[j] = 0; }
    t.E[8] = 0;
    t.fold = 0;
    ARK_UNROLL for (int i = 0; i < 8; i++) {
      if (i > 0) {  // the word folded out by the previous shift; its carry has weight 2^32 == O[0]
        t.E[0] = add_cc(t.E[0], t.fold);
        ARK_UNROLL for (int j = 0; j < 8; j++) t.O[j] = addc_cc(t.O[j], 0u);
        ARK_EMU_EXPECT_NO_CARRY();
        t.fold = 0;
      }
      acc_reduce_shift<F>(t);
      t.E[7] = add_cc(t.E[7], T[8 + i]);  // next word of the high half enters at relative position 7
      t.E[8] = addc(t.E[8], 0u);
    }
    acc_collapse(r, t);
  }

  // canonical product of canonical inputs with the Karatsuba 512-bit product: 48 + 64 wide multiply-adds against 128 for mul
  ARK_DM static void mul_kara(fe8& r, const fe8& a, const fe8& x) {
    uint32_t T[16];
    kara512(T, a.v, x.v);
    redc512_lazy(r, T);  // < p*p/R + p < 2p
    csub_p(r);
  }

  // a^-1 mod p for 0 < a < p: binary extended Euclid on plain integers, BRANCH-FREE with a fixed 2 * kBits iterations, so the 32
  // lanes of a warp (32 different inputs) never diverge — a data-dependent version measured 127 k warp-instructions (274 us) for
  // 2048 inversions because every lane waits for the union of all lanes' paths (profiles/r02i_ntt_inverse_full.txt).  Invariants:
  // x1 * a = u, x2 * a = v (mod p), u odd.  Each iteration makes v even (subtracting the smaller odd value from the larger) and
  // halves it, so len(u) + len(v) drops by at least one per iteration: 2 * kBits iterations end with v = 0, u = 1, x1 = a^-1.
  // Shifts, adds and subtractions only: used where ONE inversion sits on the critical path (top of the batch-inversion tree), in
  // place of a Fermat chain whose ~320 dependent multiplications each hold the wide-multiplier pipe ~520 cycles.
  ARK_DM static void inv_plain(fe8& r, const fe8& a) {
    uint32_t p[8];
    load_p(p);
    uint32_t u[8], v[8], x1[8], x2[8];
    ARK_UNROLL for (int j = 0; j < 8; j++) { u[j] = p[j]; v[j] = a.v[j]; x1[j] = 0; x2[j] = 0; }
    x2[0] = 1;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int it = 0; it < 2 * F::kBits; it++) {
      const uint32_t odd = 0u - (v[0] & 1u);
      // v < u ?  (borrow of v - u; carry-chain primitives: one instruction per limb)
      (void)sub_cc(v[0], u[0]);
      ARK_UNROLL for (int j = 1; j < 8; j++) (void)subc_cc(v[j], u[j]);
      const uint32_t sw = odd & subc(0u, 0u);
      // v odd and v < u: swap (u, v) and (x1, x2), so that the odd value being reduced is the larger one
      ARK_UNROLL for (int j = 0; j < 8; j++) {
        const uint32_t tu = (u[j] ^ v[j]) & sw, tx = (x1[j] ^ x2[j]) & sw;
        u[j] ^= tu; v[j] ^= tu; x1[j] ^= tx; x2[j] ^= tx;
      }
      // v odd: v <- v - u (>= 0), x2 <- x2 - x1 mod p
      v[0] = sub_cc(v[0], u[0] & odd);
      ARK_UNROLL for (int j = 1; j < 7; j++) v[j] = subc_cc(v[j], u[j] & odd);
      v[7] = subc(v[7], u[7] & odd);
      x2[0] = sub_cc(x2[0], x1[0] & odd);
      ARK_UNROLL for (int j = 1; j < 8; j++) x2[j] = subc_cc(x2[j], x1[j] & odd);
      const uint32_t neg = subc(0u, 0u);
      x2[0] = add_cc(x2[0], p[0] & neg);
      ARK_UNROLL for (int j = 1; j < 7; j++) x2[j] = addc_cc(x2[j], p[j] & neg);
      x2[7] = addc(x2[7], p[7] & neg);
      // v is even now: v /= 2, x2 /= 2 mod p (x2 odd -> (x2 + p) / 2, the sum may carry into bit 256)
      ARK_UNROLL for (int j = 0; j < 7; j++) v[j] = (v[j] >> 1) | (v[j + 1] << 31);
      v[7] >>= 1;
      const uint32_t xo = 0u - (x2[0] & 1u);
      x2[0] = add_cc(x2[0], p[0] & xo);
      ARK_UNROLL for (int j = 1; j < 8; j++) x2[j] = addc_cc(x2[j], p[j] & xo);
      const uint32_t top = addc(0u, 0u);
      ARK_UNROLL for (int j = 0; j < 7; j++) x2[j] = (x2[j] >> 1) | (x2[j + 1] << 31);
      x2[7] = (x2[7] >> 1) | (top << 31);
    }
    ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = x1[j];
  }

  // a^-1 mod p for 0 < a < p by Bernstein-Yang division steps ("safegcd", half-delta variant; the published algorithm, arranged as
  // in the 32-bit modular inverse of Pieter Wuille's bitcoin-core implementation notes): 20 rounds of 30 division steps.  A round
  // runs its 30 steps on the LOW 32 bits of f and g only (about 20 integer instructions per step, nothing 256-bit), collects them
  // in a 2x2 matrix t with |entries| <= 2^30, and then applies t once to the full-width pairs (f, g) and (d, e), kept as nine
  // signed 30-bit limbs so every column sum fits a signed 64-bit accumulator (IMAD.WIDE, signed).  Invariants: d * a = f and
  // e * a = g (mod p); f, g shrink by exact division by 2^30 per round, (d, e) by Montgomery-style division mod p.  590 steps
  // suffice for any 256-bit odd modulus; 600 leave f = +-1, g = 0, d = +-a^-1.  Branch-free and fixed-length like inv_plain, with
  // a third of its instructions: 33 us for 16 k inversions spread one warp per scheduler (profiles/r02m_*), inv_plain took 81.
  static constexpr int32_t kM30 = 0x3fffffff;
  ARK_HDM static constexpr uint32_t p_word(int j) {
    return j == 0 ? F::P0 : j == 1 ? F::P1 : j == 2 ? F::P2 : j == 3 ? F::P3 : j == 4 ? F::P4 : j == 5 ? F::P5 : j == 6 ? F::P6 : j == 7 ? F::P7 : 0u;
  }
  // limb i of p in radix 2^30
  ARK_HDM static constexpr int32_t p_limb30(int i) {
    const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
    const uint64_t two = (uint64_t)p_word(w) | ((uint64_t)p_word(w + 1) << 32);
    return (int32_t)((two >> sh) & (uint64_t)kM30);
  }
  ARK_DM static void inv_safegcd(fe8& r, const fe8& a) {
    constexpr uint32_t kPInv30 = (0u - F::INV) & (uint32_t)kM30;  // p^-1 mod 2^30 (F::INV is -p^-1 mod 2^32)
    int32_t f[9], g[9], d[9], e[9], pl[9];
    ARK_UNROLL for (int i = 0; i < 9; i++) {
      const int bit = 30 * i, w = bit >> 5, sh = bit & 31;
      const uint32_t lo = a.v[w], hi = w + 1 < 8 ? a.v[w + 1] : 0u;
      g[i] = (int32_t)((sh ? (lo >> sh) | (hi << (32 - sh)) : lo) & (uint32_t)kM30);
      pl[i] = p_limb30(i);
      f[i] = pl[i];
      d[i] = 0;
      e[i] = 0;
    }
    e[0] = 1;
    int32_t zeta = -1;  // -(delta + 1/2), delta = 1/2
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int round = 0; round < 20; round++) {
      // 30 division steps on the low words; (u, v, q, r) accumulate the transition matrix scaled by 2^30
      uint32_t u = 1, v = 0, q = 0, rr = 1;
      uint32_t fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
      ARK_UNROLL for (int i = 0; i < 30; i++) {
        uint32_t c1 = (uint32_t)(zeta >> 31);   // zeta < 0  (delta > 0)
        const uint32_t c2 = 0u - (gl & 1u);     // g odd
        const uint32_t x = (fl ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // (f, u, v) negated when delta > 0
        gl += x & c2;
        q += y & c2;
        rr += z & c2;
        c1 &= c2;                               // delta > 0 and g odd: the step swaps
        zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
        fl += gl & c1;
        u += q & c1;
        v += rr & c1;
        gl >>= 1;
        u <<= 1;
        v <<= 1;
      }
      const int32_t tu = (int32_t)u, tv = (int32_t)v, tq = (int32_t)q, tr = (int32_t)rr;
      // (d, e) <- t * (d, e) / 2^30 mod p; a multiple of p chosen per row clears the low 30 bits; both stay in (-2p, p)
      {
        const int32_t sd = d[8] >> 31, se = e[8] >> 31;
        int32_t md = (tu & sd) + (tv & se), me = (tq & sd) + (tr & se);
        int64_t cd = (int64_t)tu * d[0] + (int64_t)tv * e[0];
        int64_t ce = (int64_t)tq * d[0] + (int64_t)tr * e[0];
        md -= (int32_t)((kPInv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)kM30);
        me -= (int32_t)((kPInv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)kM30);
        cd += (int64_t)pl[0] * md;
        ce += (int64_t)pl[0] * me;
        cd >>= 30;
        ce >>= 30;
        ARK_UNROLL for (int i = 1; i < 9; i++) {
          const int32_t di = d[i], ei = e[i];
          cd += (int64_t)tu * di + (int64_t)tv * ei + (int64_t)pl[i] * md;
          ce += (int64_t)tq * di + (int64_t)tr * ei + (int64_t)pl[i] * me;
          d[i - 1] = (int32_t)cd & kM30;
          e[i - 1] = (int32_t)ce & kM30;
          cd >>= 30;
          ce >>= 30;
        }
        d[8] = (int32_t)cd;
        e[8] = (int32_t)ce;
      }
      // (f, g) <- t * (f, g) / 2^30, exactly
      {
        int64_t cf = (int64_t)tu * f[0] + (int64_t)tv * g[0];
        int64_t cg = (int64_t)tq * f[0] + (int64_t)tr * g[0];
        cf >>= 30;
        cg >>= 30;
        ARK_UNROLL for (int i = 1; i < 9; i++) {
          const int32_t fi = f[i], gi = g[i];
          cf += (int64_t)tu * fi + (int64_t)tv * gi;
          cg += (int64_t)tq * fi + (int64_t)tr * gi;
          f[i - 1] = (int32_t)cf & kM30;
          g[i - 1] = (int32_t)cg & kM30;
          cf >>= 30;
          cg >>= 30;
        }
        f[8] = (int32_t)cf;
        g[8] = (int32_t)cg;
      }
    }
    // d = sign(f) * a^-1 in (-2p, p): add p if negative, negate if f = -1, add p again if still negative -> [0, p)
    {
      int32_t add = d[8] >> 31;
      const int32_t neg = f[8] >> 31;
      ARK_UNROLL for (int i = 0; i < 9; i++) d[i] = ((d[i] + (pl[i] & add)) ^ neg) - neg;
      ARK_UNROLL for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= kM30; }
      add = d[8] >> 31;
      ARK_UNROLL for (int i = 0; i < 9; i++) d[i] += pl[i] & add;
      ARK_UNROLL for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= kM30; }
    }
    ARK_UNROLL for (int j = 0; j < 8; j++) {  // radix 2^30 -> 2^32
      const int bit = 32 * j, i = bit / 30, sh = bit % 30;
      uint32_t w = (uint32_t)d[i] >> sh;
      w |= (uint32_t)d[i + 1] << (30 - sh);
      if (60 - sh < 32 && i + 2 < 9) w |= (uint32_t)d[i + 2] << (60 - sh);
      r.v[j] = w;
    }
  }

  // Montgomery image of a^-1 from the Montgomery image of a != 0: (aR)^-1 = a^-1 R^-1, then two multiplications by R^2
  ARK_DM static void inv_mont(fe8& r, const fe8& a_mont) {
    fe8 y, r2;
    inv_safegcd(y, a_mont);
    set_r2(r2);
    mul(y, y, r2);
    mul(r, y, r2);
  }

  ARK_DM static bool is_zero(const fe8& a) {
    uint32_t nz = 0;
    ARK_UNROLL for (int j = 0; j < 8; j++) nz |= a.v[j];
    return nz == 0;
  }
  ARK_DM static bool eq(const fe8& a, const fe8& b) {
    uint32_t d = 0;
    ARK_UNROLL for (int j = 0; j < 8; j++) d |= a.v[j] ^ b.v[j];
    return d == 0;
  }
  // a < p ?
  ARK_DM static bool is_canonical(const fe8& a) {
    uint32_t p[8];
    load_p(p);
    (void)sub_cc(a.v[0], p[0]);
    ARK_UNROLL for (int j = 1; j < 8; j++) (void)subc_cc(a.v[j], p[j]);
    return subc(0u, 0u) != 0u;
  }
  ARK_HDM static void set_one(fe8& r) {
    r.v[0] = F::R_0; r.v[1] = F::R_1; r.v[2] = F::R_2; r.v[3] = F::R_3; r.v[4] = F::R_4; r.v[5] = F::R_5; r.v[6] = F::R_6; r.v[7] = F::R_7;
  }
  ARK_HDM static void set_r2(fe8& r) {
    r.v[0] = F::R2_0; r.v[1] = F::R2_1; r.v[2] = F::R2_2; r.v[3] = F::R2_3; r.v[4] = F::R2_4; r.v[5] = F::R2_5; r.v[6] = F::R2_6; r.v[7] = F::R2_7;
  }
  ARK_HDM static void set_zero(fe8& r) { ARK_UNROLL for (int j = 0; j < 8; j++) r.v[j] = 0; }
};

}  // namespace ark
