// Curve-group arithmetic for the point gates: one point (or one point gate) per thread.
//
// What it replaces (paths under /root/reference/online-phase/src/algebra/curve): the plaintext group
// operations behind `CurvePoint<C>` (curve.rs:194-209 add, :403-409 scalar-mul, both forwarding to
// ark-ec 0.4 `CurveGroup`, which is NOT vendored) for the two supported groups:
//   * BN254 G1  — short Weierstrass y^2 = x^3 + 3 over Fq, Jacobian (X,Y,Z), identity Z = 0
//                 (the memory image of ark-bn254 `G1Projective`: x,y,z, 3 x 32 B Montgomery residues)
//   * Curve25519 Edwards — twisted Edwards -x^2 + y^2 = 1 + d x^2 y^2 over 2^255-19, extended (X,Y,T,Z),
//                 identity (0,1,0,1) (ark-curve25519 `EdwardsProjective`: x,y,t,z, 4 x 32 B)
// Projective representatives are not unique (curve.rs:46 compares through arkworks' projective equality), so
// parity with the reference is defined on the affine form; `normalize` produces it.
//
// Scalar multiplication: the reference does a per-element MSB-first double-and-add (`self.0 * rhs.0`).  Here:
//   variable base  4-bit fixed windows over a per-thread table {1..15}P kept in local memory (252 doublings + <=64 adds)
//   fixed base G   table {w * 256^j * G : j < 32, 1 <= w <= 255} in global memory, built once per context
//                  (no doublings, <= 32 mixed additions)
// and the callers regroup sums of products that share a base into one pass (see curve_kernels.cuh).
//
// Dual-target like fp256.cuh: compiles under g++ for the host-emulation tests.
#pragma once
#include "f25519.cuh"
#include "fp256.cuh"

namespace ark {

// ----------------------------------------------------------------------------------------------
// Base-field helpers (canonical Montgomery residues throughout; in-place aliasing is allowed)
// ----------------------------------------------------------------------------------------------
template <class Q>
struct Fq {
  using F = Fp<Q>;
  ARK_FQ_MUL static void mul(fe8& r, const fe8& a, const fe8& b) { F::mul(r, a, b); }
  // Fp::sqr (dedicated 512-bit square + separate word-serial reduction, 100 multiply-adds) measured SLOWER than the
  // interleaved CIOS product on B200 (BN254 point Beaver 3.95 M -> 3.66 M mults/s: the un-interleaved reduction's carry ripples
  // cost more than the 28 multiplies saved), so the Montgomery fields square with mul; F25519 keeps its dedicated square (+6 %).
  ARK_FQ_MUL static void sqr(fe8& r, const fe8& a) { F::mul(r, a, a); }
  ARK_DM static void add(fe8& r, const fe8& a, const fe8& b) { F::add(r, a, b); }
  ARK_DM static void sub(fe8& r, const fe8& a, const fe8& b) { F::sub(r, a, b); }
  ARK_DM static void dbl(fe8& r, const fe8& a) { F::add(r, a, a); }
  ARK_DM static void neg(fe8& r, const fe8& a) { F::neg(r, a); }
  ARK_DM static bool is_zero(const fe8& a) { return F::is_zero(a); }
  ARK_DM static bool eq(const fe8& a, const fe8& b) { return F::eq(a, b); }
  ARK_DM static void one(fe8& r) { F::set_one(r); }
  ARK_DM static void zero(fe8& r) { F::set_zero(r); }
  // Montgomery image of a^-1 by division steps (Fp::inv_safegcd: about the instructions of 110 multiplications, against
  // 256 squarings + ~128 multiplications for a Fermat chain); inv(0) = 0
  ARK_DM static void inv(fe8& r, const fe8& a) {
    if (is_zero(a)) { zero(r); return; }
    F::inv_mont(r, a);
  }
};

// Scalar-field element (Montgomery image) -> plain integer limbs, for window extraction.
template <class R>
ARK_D void scalar_to_plain(uint32_t* k, const fe8& s_mont) {
  fe8 one, r;
  Fp<R>::set_zero(one);
  one.v[0] = 1;
  Fp<R>::mul(r, s_mont, one);
  ARK_UNROLL for (int j = 0; j < 8; j++) k[j] = r.v[j];
}
ARK_D uint32_t window4(const uint32_t* k, int i) { return (k[i >> 3] >> ((i & 7) * 4)) & 15u; }

// ----------------------------------------------------------------------------------------------
// BN254 G1: Jacobian coordinates, a = 0
// ----------------------------------------------------------------------------------------------
struct PtSW { fe8 X, Y, Z; };
struct AffSW { fe8 x, y; };

struct Bn254G1 {
  static constexpr int kId = 0;
  using Q = Bn254Fq;
  using R = Bn254Fr;
  using Pt = PtSW;
  using Cached = PtSW;     // variable-base table entry
  using Aff = AffSW;       // fixed-base table entry (never the identity)
  using K = Fq<Q>;
  static constexpr int kCoords = 3;
  static constexpr int kPointBytes = 96;
  // resident 128-thread blocks per SM the scalar-multiplication kernels are compiled for: the Montgomery field needs ~220
  // registers, capping them spills and measured slower (profiles/r01f_pt_minblocks_ab.txt)
  static constexpr int kMinBlocks = 1;
  static constexpr bool kGlv = true;         // variable-base multiplications use the endomorphism (x, y) -> (beta x, y), see var_mul_glv
  // The recombination's two variable-base passes in lock-step were +5 % in round 1 (3.94 -> 4.14 M mults/s), but a lock-step
  // loop holds twice the inlined additions; with the single-copy loops of round 2 the sequential passes win (profiles/r02f_*).
  static constexpr bool kDualChain = false;
  static constexpr int kSplitParts = 2;      // tables of the two-pass multiplications: P and 2^68 P (with the endomorphism: four sub-scalars)
  static constexpr bool kAffineTables = true;  // window tables normalised with one shared inversion: mixed additions in the loops
  static constexpr int kAffWords = 16;  // u32 words per fixed-table entry
  // threads per block of the two-pass kernels (recombine, mul_authenticated): ONE block per SM whose warps move through the
  // loop body together (ARK_PHASE_SYNC); 230 registers per thread allow 256 threads
  static constexpr int kTwoPassBlock = 256;

  // reference memory image <-> internal representation: the same (canonical Montgomery residues)
  ARK_DM static void from_image(Pt&) {}
  ARK_DM static void to_image(Pt&) {}
  ARK_DM static void set_identity(Pt& p) { K::one(p.X); K::one(p.Y); K::zero(p.Z); }  // ark-ec: (1,1,0)
  ARK_DM static bool is_identity(const Pt& p) { return K::is_zero(p.Z); }
  ARK_DM static void set_generator(Pt& p) {  // (1, 2)
    K::one(p.X);
    K::add(p.Y, p.X, p.X);
    K::one(p.Z);
  }
  ARK_DM static void neg(Pt& p) { K::neg(p.Y, p.Y); }
  ARK_DM static void cache(Cached& c, const Pt& p) { c = p; }
  ARK_DM static void cached_neg(Cached& c) { K::neg(c.Y, c.Y); }
  // Y^2 = X^3 + 3 Z^6 (Jacobian); the identity (Z = 0) is on the curve.  G1 has cofactor 1: on-curve is in-subgroup.
  ARK_DM static bool on_curve(const Pt& p) {
    if (is_identity(p)) return true;
    fe8 y2, x3, z2, z6, b, t;
    K::sqr(y2, p.Y);
    K::sqr(x3, p.X);
    K::mul(x3, x3, p.X);
    K::sqr(z2, p.Z);
    K::sqr(z6, z2);
    K::mul(z6, z6, z2);
    K::one(b);
    K::add(t, b, b);
    K::add(b, t, b);  // 3
    K::mul(z6, z6, b);
    K::add(x3, x3, z6);
    return K::eq(y2, x3);
  }
  static constexpr bool kNeedsSubgroupCheck = false;

  // dbl-2009-l (2M + 5S); Z = 0 stays Z = 0
  ARK_DM static void dbl(Pt& p, bool = true) {
    fe8 A, B, C, D, E, F, t;
    K::sqr(A, p.X);
    K::sqr(B, p.Y);
    K::sqr(C, B);
    K::add(t, p.X, B);
    K::sqr(t, t);
    K::sub(t, t, A);
    K::sub(t, t, C);
    K::dbl(D, t);
    K::dbl(E, A);
    K::add(E, E, A);
    K::sqr(F, E);
    K::mul(p.Z, p.Y, p.Z);
    K::dbl(p.Z, p.Z);
    K::dbl(t, D);
    K::sub(p.X, F, t);
    K::sub(t, D, p.X);
    K::mul(t, E, t);
    K::dbl(C, C);
    K::dbl(C, C);
    K::dbl(C, C);
    K::sub(p.Y, t, C);
  }

  // add-2007-bl (11M + 5S) with the exceptional cases handled
  ARK_DM static void add(Pt& p, const Pt& q) {
    if (is_identity(q)) return;
    if (is_identity(p)) { p = q; return; }
    fe8 Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, r, V, t;
    K::sqr(Z1Z1, p.Z);
    K::sqr(Z2Z2, q.Z);
    K::mul(U1, p.X, Z2Z2);
    K::mul(U2, q.X, Z1Z1);
    K::mul(S1, p.Y, q.Z);
    K::mul(S1, S1, Z2Z2);
    K::mul(S2, q.Y, p.Z);
    K::mul(S2, S2, Z1Z1);
    K::sub(H, U2, U1);
    K::sub(r, S2, S1);
    if (K::is_zero(H)) {
      if (K::is_zero(r)) { dbl(p); return; }
      set_identity(p);
      return;
    }
    K::dbl(r, r);
    K::dbl(I, H);
    K::sqr(I, I);
    K::mul(J, H, I);
    K::mul(V, U1, I);
    K::add(t, p.Z, q.Z);
    K::sqr(t, t);
    K::sub(t, t, Z1Z1);
    K::sub(t, t, Z2Z2);
    K::mul(p.Z, t, H);
    K::sqr(p.X, r);
    K::sub(p.X, p.X, J);
    K::dbl(t, V);
    K::sub(p.X, p.X, t);
    K::sub(t, V, p.X);
    K::mul(t, r, t);
    K::mul(S1, S1, J);
    K::dbl(S1, S1);
    K::sub(p.Y, t, S1);
  }
  ARK_DM static void add_cached(Pt& p, const Cached& q) { add(p, q); }

  // madd-2007-bl (7M + 4S), q affine and not the identity
  ARK_DM static void madd(Pt& p, const Aff& q) {
    if (is_identity(p)) { p.X = q.x; p.Y = q.y; K::one(p.Z); return; }
    fe8 Z1Z1, U2, S2, H, HH, I, J, r, V, t;
    K::sqr(Z1Z1, p.Z);
    K::mul(U2, q.x, Z1Z1);
    K::mul(S2, q.y, p.Z);
    K::mul(S2, S2, Z1Z1);
    K::sub(H, U2, p.X);
    K::sub(r, S2, p.Y);
    if (K::is_zero(H)) {
      if (K::is_zero(r)) { dbl(p); return; }
      set_identity(p);
      return;
    }
    K::dbl(r, r);
    K::sqr(HH, H);
    K::dbl(I, HH);
    K::dbl(I, I);
    K::mul(J, H, I);
    K::mul(V, p.X, I);
    K::add(t, p.Z, H);
    K::sqr(t, t);
    K::sub(t, t, Z1Z1);
    K::sub(p.Z, t, HH);
    K::sqr(p.X, r);
    K::sub(p.X, p.X, J);
    K::dbl(t, V);
    K::sub(p.X, p.X, t);
    K::sub(t, V, p.X);
    K::mul(t, r, t);
    K::mul(J, p.Y, J);
    K::dbl(J, J);
    K::sub(p.Y, t, J);
  }

  // affine (x, y); the identity maps to (0, 0), which is not on the curve
  ARK_DM static void normalize(fe8& x, fe8& y, const Pt& p) {
    if (is_identity(p)) { K::zero(x); K::zero(y); return; }
    fe8 zi, zi2;
    K::inv(zi, p.Z);
    K::sqr(zi2, zi);
    K::mul(x, p.X, zi2);
    K::mul(zi2, zi2, zi);
    K::mul(y, p.Y, zi2);
  }
  ARK_DM static void to_aff(Aff& a, const Pt& p) { normalize(a.x, a.y, p); }
  // inverse of normalize: canonical affine image (Montgomery residues) -> projective point with Z = 1; (0, 0) is the identity
  ARK_DM static void from_affine(Pt& p, const fe8& x, const fe8& y) {
    if (K::is_zero(x) && K::is_zero(y)) { set_identity(p); return; }
    p.X = x;
    p.Y = y;
    K::one(p.Z);
  }
};

// ----------------------------------------------------------------------------------------------
// Curve25519 in twisted-Edwards form (a = -1): extended coordinates, complete unified addition
// ----------------------------------------------------------------------------------------------
struct PtTE { fe8 X, Y, T, Z; };
struct CachedTE { fe8 YpX, YmX, Z2, T2d; };  // (Y+X, Y-X, 2Z, 2d*T)
struct NielsTE { fe8 ypx, ymx, xy2d; };      // affine: (y+x, y-x, 2d*x*y)

struct Ed25519 {
  static constexpr int kId = 1;
  using Q = Curve25519Fq;
  using R = Curve25519Fr;
  using Pt = PtTE;
  using Cached = CachedTE;
  using Aff = NielsTE;
  using K = F25519;  // plain residues mod 2p with special-form reduction (f25519.cuh); images are converted on load / store
  static constexpr int kCoords = 4;
  static constexpr int kPointBytes = 128;
  static constexpr bool kGlv = false;        // no efficient endomorphism on Curve25519
  static constexpr bool kDualChain = false;  // measured slower here (8.68 -> 7.2-7.9 M mults/s): the 128-register build already keeps 16 warps busy
  static constexpr int kSplitParts = 4;      // tables of the two-pass multiplications: P, 2^64 P, 2^128 P, 2^192 P (see var_mul_split)
  static constexpr bool kAffineTables = false;  // the extended-coordinates addition gains one multiplication from Z = 1: not worth an inversion
  static constexpr int kMinBlocks = 4;  // 128 registers, <= 216 B of spills, +5 % over the unconstrained 230-register build
  static constexpr int kAffWords = 24;
  static constexpr int kTwoPassBlock = 512;  // one 16-warp block per SM at 128 registers per thread (see Bn254G1::kTwoPassBlock)

  ARK_DM static void from_image(Pt& p) { K::from_image(p.X, p.X); K::from_image(p.Y, p.Y); K::from_image(p.T, p.T); K::from_image(p.Z, p.Z); }
  ARK_DM static void to_image(Pt& p) { K::to_image(p.X, p.X); K::to_image(p.Y, p.Y); K::to_image(p.T, p.T); K::to_image(p.Z, p.Z); }
  ARK_DM static void set_2d(fe8& r) {  // 2d, plain
    r.v[0] = 0x26b2f159u; r.v[1] = 0xebd69b94u; r.v[2] = 0x8283b156u; r.v[3] = 0x00e0149au;
    r.v[4] = 0xeef3d130u; r.v[5] = 0x198e80f2u; r.v[6] = 0x56dffce7u; r.v[7] = 0x2406d9dcu;
  }
  ARK_DM static void set_identity(Pt& p) { K::zero(p.X); K::one(p.Y); K::zero(p.T); K::one(p.Z); }
  ARK_DM static bool is_identity(const Pt& p) { return K::is_zero(p.X) && K::eq(p.Y, p.Z); }
  ARK_DM static void set_generator(Pt& p) {  // RFC 8032 base point, plain residues
    p.X.v[0] = 0x8f25d51au; p.X.v[1] = 0xc9562d60u; p.X.v[2] = 0x9525a7b2u; p.X.v[3] = 0x692cc760u;
    p.X.v[4] = 0xfdd6dc5cu; p.X.v[5] = 0xc0a4e231u; p.X.v[6] = 0xcd6e53feu; p.X.v[7] = 0x216936d3u;
    p.Y.v[0] = 0x66666658u;
    ARK_UNROLL for (int j = 1; j < 8; j++) p.Y.v[j] = 0x66666666u;
    K::one(p.Z);
    K::mul(p.T, p.X, p.Y);
  }
  ARK_DM static void neg(Pt& p) { K::neg(p.X, p.X); K::neg(p.T, p.T); }
  ARK_DM static void cache(Cached& c, const Pt& p) {
    fe8 k;
    set_2d(k);
    K::add(c.YpX, p.Y, p.X);
    K::sub(c.YmX, p.Y, p.X);
    K::dbl(c.Z2, p.Z);
    K::mul(c.T2d, p.T, k);
  }
  // extended twisted Edwards, a = -1: Z != 0, T Z = X Y and 2 (Y^2 - X^2) Z^2 = 2 Z^4 + (2d) X^2 Y^2.  The group has cofactor 8:
  // membership of the prime-order subgroup is checked separately ([l]P = identity).
  ARK_DM static bool on_curve(const Pt& p) {
    if (K::is_zero(p.Z)) return false;
    fe8 x2, y2, z2, l, r, t, k;
    K::mul(l, p.T, p.Z);
    K::mul(r, p.X, p.Y);
    if (!K::eq(l, r)) return false;
    K::sqr(x2, p.X);
    K::sqr(y2, p.Y);
    K::sqr(z2, p.Z);
    K::sub(l, y2, x2);
    K::mul(l, l, z2);
    K::dbl(l, l);
    K::sqr(r, z2);
    K::dbl(r, r);
    set_2d(k);
    K::mul(t, x2, y2);
    K::mul(t, t, k);
    K::add(r, r, t);
    return K::eq(l, r);
  }
  static constexpr bool kNeedsSubgroupCheck = true;
  ARK_DM static void cached_neg(Cached& c) {  // -(X, Y, T, Z) = (-X, Y, -T, Z): swap Y+X and Y-X, negate 2dT
    const fe8 t = c.YpX;
    c.YpX = c.YmX;
    c.YmX = t;
    K::neg(c.T2d, c.T2d);
  }

  // dbl-2008-hwcd with a = -1 (4M + 4S).  T is consumed only by additions, so inside a run of doublings it is computed
  // for the last one only (`need_t`), saving one multiplication per skipped T.
  ARK_DM static void dbl(Pt& p, bool need_t = true) {
    fe8 A, B, C, E, G, F, H;
    K::sqr(A, p.X);
    K::sqr(B, p.Y);
    K::sqr(C, p.Z);
    K::dbl(C, C);
    K::add(E, p.X, p.Y);
    K::sqr(E, E);
    K::sub(E, E, A);
    K::sub(E, E, B);   // E = 2XY
    K::sub(G, B, A);   // D + B with D = -A
    K::sub(F, G, C);
    K::add(H, A, B);
    K::neg(H, H);      // D - B = -(A + B)
    K::mul(p.X, E, F);
    K::mul(p.Y, G, H);
    if (need_t) K::mul(p.T, E, H);
    K::mul(p.Z, F, G);
  }

  // add-2008-hwcd-3 against a cached operand (8M)
  ARK_DM static void add_cached(Pt& p, const Cached& q) {
    fe8 A, B, C, D, E, F, G, H;
    K::sub(A, p.Y, p.X);
    K::mul(A, A, q.YmX);
    K::add(B, p.Y, p.X);
    K::mul(B, B, q.YpX);
    K::mul(C, p.T, q.T2d);
    K::mul(D, p.Z, q.Z2);
    K::sub(E, B, A);
    K::sub(F, D, C);
    K::add(G, D, C);
    K::add(H, B, A);
    K::mul(p.X, E, F);
    K::mul(p.Y, G, H);
    K::mul(p.T, E, H);
    K::mul(p.Z, F, G);
  }
  ARK_DM static void add(Pt& p, const Pt& q) {
    Cached c;
    cache(c, q);
    add_cached(p, c);
  }
  // mixed addition with an affine Niels operand (7M)
  ARK_DM static void madd(Pt& p, const Aff& q) {
    fe8 A, B, C, D, E, F, G, H;
    K::sub(A, p.Y, p.X);
    K::mul(A, A, q.ymx);
    K::add(B, p.Y, p.X);
    K::mul(B, B, q.ypx);
    K::mul(C, p.T, q.xy2d);
    K::dbl(D, p.Z);
    K::sub(E, B, A);
    K::sub(F, D, C);
    K::add(G, D, C);
    K::add(H, B, A);
    K::mul(p.X, E, F);
    K::mul(p.Y, G, H);
    K::mul(p.T, E, H);
    K::mul(p.Z, F, G);
  }

  // affine (x, y) as internal residues
  ARK_DM static void affine(fe8& x, fe8& y, const Pt& p) {
    fe8 zi;
    K::inv(zi, p.Z);
    K::mul(x, p.X, zi);
    K::mul(y, p.Y, zi);
  }
  // canonical affine form in the reference's image (Montgomery residues)
  ARK_DM static void normalize(fe8& x, fe8& y, const Pt& p) {
    affine(x, y, p);
    K::to_image(x, x);
    K::to_image(y, y);
  }
  // inverse of normalize: canonical affine image (Montgomery residues) -> extended point (x, y, xy, 1), internal residues
  ARK_DM static void from_affine(Pt& p, const fe8& x, const fe8& y) {
    K::from_image(p.X, x);
    K::from_image(p.Y, y);
    K::mul(p.T, p.X, p.Y);
    K::one(p.Z);
  }
  ARK_DM static void to_aff(Aff& a, const Pt& p) {
    fe8 x, y, k;
    affine(x, y, p);
    set_2d(k);
    K::add(a.ypx, y, x);
    K::sub(a.ymx, y, x);
    K::mul(a.xy2d, x, y);
    K::mul(a.xy2d, a.xy2d, k);
  }
};

// ----------------------------------------------------------------------------------------------
// GLV decomposition for BN254 G1 (Gallant-Lambert-Vanstone).  phi(x, y) = (beta x, y) is the multiplication by lambda, with
//   beta   = 2203960485148121921418603742825762020974279258880205651966             (a cube root of unity in Fq)
//   lambda = 4407920970296243842393367215006156084916469457145843978461             (a cube root of unity in Fr)
// and (a1, b1) = (9931322734385697763, -147946756881789319000765030803803410728), (a2, b2) = (147946756881789319010696353538189108491,
// 9931322734385697763) is a reduced basis of the lattice {(x, y) : x + y lambda = 0 mod r} (determinant r).  For any integers c1, c2
//   k1 = k - c1 a1 - c2 a2,  k2 = -c1 b1 - c2 b2   satisfy   k1 + k2 lambda = k (mod r);
// with c1 = floor(k g1 / 2^256), c2 = floor(k g2 / 2^256), g1 = floor(2^256 b2 / r), g2 = floor(2^256 |b1| / r) both halves stay below
// 2^128 in magnitude (2^127 observed over 2*10^5 scalars incl. 0, 1, r-1, lambda), so k P = k1 P + k2 phi(P) needs half the doublings.
// Plain 64-bit C arithmetic: ~100 multiplications per scalar, irrelevant next to the ~2000 field multiplications it saves.
// ----------------------------------------------------------------------------------------------
ARK_D void limbs_mul(uint32_t* out, int nout, const uint32_t* a, int na, const uint32_t* b, int nb) {
  for (int i = 0; i < nout; i++) out[i] = 0;
  for (int i = 0; i < na; i++) {
    uint64_t carry = 0;
    for (int j = 0; j < nb && i + j < nout; j++) {
      const uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
      out[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    for (int p = i + nb; carry && p < nout; p++) {
      const uint64_t t = (uint64_t)out[p] + carry;
      out[p] = (uint32_t)t;
      carry = t >> 32;
    }
  }
}
// r = a - b (mod 2^(32 n)), b given with nb <= n limbs (zero-extended)
ARK_D void limbs_sub(uint32_t* r, const uint32_t* a, const uint32_t* b, int n, int nb) {
  uint64_t borrow = 0;
  for (int i = 0; i < n; i++) {
    const uint64_t t = (uint64_t)a[i] - (i < nb ? b[i] : 0u) - borrow;
    r[i] = (uint32_t)t;
    borrow = (t >> 32) & 1u;
  }
}
// two's-complement magnitude: if the top bit of the n-limb value is set, negate it and report the sign
ARK_D bool limbs_abs(uint32_t* v, int n) {
  if (!(v[n - 1] >> 31)) return false;
  uint64_t carry = 1;
  for (int i = 0; i < n; i++) {
    const uint64_t t = (uint64_t)(~v[i]) + carry;
    v[i] = (uint32_t)t;
    carry = t >> 32;
  }
  return true;
}

// k (plain integer < r, 8 limbs) -> |k1|, |k2| (5 limbs each, < 2^132) and their signs
ARK_D void glv_decompose_bn254(const uint32_t* k, uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2) {
  const uint32_t g1[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x00000002u};
  const uint32_t g2[5] = {0x391eb18du, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x00000002u};
  const uint32_t a1[2] = {0x94d213e3u, 0x89d32568u};                            // = b2
  const uint32_t b1n[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};  // |b1|
  const uint32_t a2[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};
  uint32_t t[13], c1[3], c2[5];
  limbs_mul(t, 11, k, 8, g1, 3);
  for (int i = 0; i < 3; i++) c1[i] = t[8 + i];
  limbs_mul(t, 13, k, 8, g2, 5);
  for (int i = 0; i < 5; i++) c2[i] = t[8 + i];
  // k1 = k - c1 a1 - c2 a2 over 9 limbs (two's complement)
  uint32_t p[9], acc[9];
  for (int i = 0; i < 8; i++) acc[i] = k[i];
  acc[8] = 0;
  limbs_mul(p, 9, c1, 3, a1, 2);
  limbs_sub(acc, acc, p, 9, 9);
  limbs_mul(p, 9, c2, 5, a2, 4);
  limbs_sub(acc, acc, p, 9, 9);
  neg1 = limbs_abs(acc, 9);
  for (int i = 0; i < 5; i++) k1[i] = acc[i];
  // k2 = c1 |b1| - c2 b2 over 8 limbs (b2 = a1)
  uint32_t u[8], v[8];
  limbs_mul(u, 8, c1, 3, b1n, 4);
  limbs_mul(v, 8, c2, 5, a1, 2);
  limbs_sub(u, u, v, 8, 8);
  neg2 = limbs_abs(u, 8);
  for (int i = 0; i < 5; i++) k2[i] = u[i];
}

// ----------------------------------------------------------------------------------------------
// Scalar multiplication building blocks
// ----------------------------------------------------------------------------------------------
// Variable base: SIGNED 4-bit windows.  With k' = k + 0x88...8 the digit of window i is nibble_i(k') - 8 in [-8, 7], so the
// table holds only 1P..8P (8 entries instead of 15; negation of a table entry is free) and a 256-bit scalar still costs 252
// doublings and at most 64 additions.  The table lives behind a small store interface (`put` / `get`) so that the same code
// runs on a per-thread array (host emulation, small kernels) and on the kernels' L2-resident scratch records (curve_kernels.cuh:
// a per-thread array indexed by a divergent window would be local memory, whose word-interleaved layout turns every entry
// read into 32 sectors per warp instruction — 8.7 GB of DRAM traffic per 2^17-element launch in round 1).
constexpr int kWindows = 64;      // 4-bit windows of a 256-bit scalar
constexpr int kTabEntries = 8;    // 1P .. 8P (entry w-1 holds w*P)

template <class C>
struct LocalTab {  // a view of kTabEntries entries; sub(p) is the p-th table behind it (the split multiplications keep up to four)
  static constexpr bool kAffine = C::kAffineTables;  // entry format once the tables are finished (finish_tables)
  typename C::Cached* t;
  fe8* aux;  // one spare field element per entry (normalize_tables)
  template <class T> ARK_DM void put(int idx, const T& c) const {
    static_assert(sizeof(T) <= sizeof(typename C::Cached), "entry too large");
    memcpy(static_cast<void*>(&t[idx]), &c, sizeof(T));
  }
  template <class T> ARK_DM void get(T& c, int idx) const { memcpy(&c, static_cast<const void*>(&t[idx]), sizeof(T)); }
  ARK_DM void put_aux(int idx, const fe8& v) const { aux[idx] = v; }
  ARK_DM void get_aux(fe8& v, int idx) const { v = aux[idx]; }
  ARK_DM void prefetch(int) const {}
  ARK_DM LocalTab sub(int p) const { return LocalTab{t + p * kTabEntries, aux + p * kTabEntries}; }
};
template <class C, int PARTS = 1>
struct LocalTabStore {
  typename C::Cached t[PARTS * kTabEntries];
  fe8 aux[PARTS * kTabEntries];
  ARK_DM LocalTab<C> view() { return LocalTab<C>{t, aux}; }
};

// store w*P for w = 1..8
template <class C, class Tab>
ARK_D void build_table(Tab& tab, const typename C::Pt& P) {
  typename C::Pt t = P;
  typename C::Cached c1, c;
  C::cache(c1, t);
  tab.put(0, c1);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int w = 2; w <= kTabEntries; w++) {
    C::add_cached(t, c1);
    C::cache(c, t);
    tab.put(w - 1, c);
  }
}

// Tab::kAffine (BN254, C::kAffineTables, wherever the kernel asks for it): the entries just built (Jacobian, `nparts` tables of kTabEntries behind `tab`) are brought to affine
// form with ONE shared inversion (Montgomery's trick over the Z coordinates: a prefix product per entry parked in the entry's
// spare 32 bytes, a safegcd inversion, a walk back), so that every addition of the multiplication loops is the mixed one — 7M + 4S
// instead of 11M + 5S, 132 times per two-pass gate — for ~7 multiplications per entry and an inversion that costs about as many
// instructions as 110 multiplications.  The identity (Z = 0) becomes (0, 0), which is not on the curve and is skipped by glv_add.
template <class C, class Tab>
ARK_D void normalize_tables(const Tab& tab, int nparts) {
  using K = typename C::K;
  const int n = nparts * kTabEntries;
  fe8 acc;
  K::one(acc);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int e = 0; e < n; e++) {
    typename C::Pt q;
    tab.get(q, e);
    tab.put_aux(e, acc);  // product of the (non-zero) Z before this entry
    if (!K::is_zero(q.Z)) K::mul(acc, acc, q.Z);
  }
  fe8 inv;
  Fp<typename C::Q>::inv_mont(inv, acc);  // acc != 0: zeros were skipped
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int e = n - 1; e >= 0; e--) {
    typename C::Pt q;
    typename C::Aff a;
    fe8 pre, zi, zi2;
    tab.get(q, e);
    tab.get_aux(pre, e);
    if (K::is_zero(q.Z)) {
      K::zero(a.x);
      K::zero(a.y);
    } else {
      K::mul(zi, inv, pre);     // 1 / Z_e
      K::mul(inv, inv, q.Z);    // drop Z_e from the running inverse
      K::sqr(zi2, zi);
      K::mul(a.x, q.X, zi2);
      K::mul(zi2, zi2, zi);
      K::mul(a.y, q.Y, zi2);
    }
    tab.put(e, a);
  }
}
// called once the tables of a multiplication are built
template <class C, class Tab>
ARK_D void finish_tables(const Tab& tab, int nparts) {
  if constexpr (Tab::kAffine) normalize_tables<C>(tab, nparts);
}

// k' = k + 0x8888...8 over `limbs` words; returns the carry out of the top word (an extra, non-negative top digit)
ARK_D uint32_t signed_recode(uint32_t* kk, const uint32_t* k, int limbs) {
  uint64_t c = 0;
  for (int j = 0; j < limbs; j++) {
    c += (uint64_t)k[j] + 0x88888888u;
    kk[j] = (uint32_t)c;
    c >>= 32;
  }
  return (uint32_t)c;
}

// request the table entry a signed digit will need (no-op for digit 0 and for array-backed tables)
template <class Tab>
ARK_D void prefetch_signed(const Tab& tab, int digit) {
  if (digit != 0) tab.prefetch((digit < 0 ? -digit : digit) - 1);
}

// acc += digit * P for a signed digit in [-8, 8]
template <class C, class Tab>
ARK_D void add_signed(typename C::Pt& acc, const Tab& tab, int digit) {
  if (digit == 0) return;
  typename C::Cached c;
  tab.get(c, (digit < 0 ? -digit : digit) - 1);
  if (digit < 0) C::cached_neg(c);
  C::add_cached(acc, c);
}

// acc = k * P from the table (acc must be the identity on entry)
template <class C, class Tab> ARK_D void var_mul_glv(typename C::Pt& acc, const Tab& tab, const uint32_t* k);
template <class C, class Tab>
ARK_D void var_mul2_glv(typename C::Pt& acc0, typename C::Pt& acc1, const Tab& tab, const uint32_t* ka, const uint32_t* kb);

template <class C, class Tab>
ARK_D void var_mul(typename C::Pt& acc, const Tab& tab, const uint32_t* k) {
  if constexpr (C::kGlv) {
    var_mul_glv<C>(acc, tab, k);
  } else {
    uint32_t kk[8];
    const uint32_t top = signed_recode(kk, k, 8);  // k < r < 2^254: the top nibble is <= 3, never carries out
    (void)top;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = kWindows - 1; i >= 0; i--) {
      const int digit = (int)window4(kk, i) - 8;
      prefetch_signed(tab, digit);
      if (i != kWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) C::dbl(acc, j == 3);  // every window ends in (possibly) an addition; the final T is part of the result
      }
      add_signed<C>(acc, tab, digit);
    }
  }
}

// (acc0, acc1) = (k0 * P, k1 * P): two accumulators advanced in lock-step over one table.  The two chains are independent, so the
// instruction scheduler can interleave them: the kernels are bound by the latency of dependent carry chains at 8-16 warps per
// SM, and a second chain per thread fills the bubbles (C::kDualChain selects it per curve, profiles/r01g_dual_chain_ab.txt).
template <class C, class Tab>
ARK_D void var_mul2(typename C::Pt& acc0, typename C::Pt& acc1, const Tab& tab, const uint32_t* k0, const uint32_t* k1) {
  if constexpr (C::kGlv) {
    var_mul2_glv<C>(acc0, acc1, tab, k0, k1);
  } else {
    uint32_t kk0[8], kk1[8];
    signed_recode(kk0, k0, 8);
    signed_recode(kk1, k1, 8);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = kWindows - 1; i >= 0; i--) {
      const int d0 = (int)window4(kk0, i) - 8, d1 = (int)window4(kk1, i) - 8;
      prefetch_signed(tab, d0);
      prefetch_signed(tab, d1);
      if (i != kWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) {
          C::dbl(acc0, j == 3);
          C::dbl(acc1, j == 3);
        }
      }
      add_signed<C>(acc0, tab, d0);
      add_signed<C>(acc1, tab, d1);
    }
  }
}

// GLV variants (C::kGlv): acc = k1 P + k2 phi(P) with |k1|, |k2| < 2^128: 33 signed windows, 128 doublings.  phi and the signs
// are applied to the looked-up table entry (one multiplication by beta, one negation), so both halves share the table of P.
constexpr int kGlvWindows = 33;

template <class C, class Tab>
ARK_D void glv_add(typename C::Pt& acc, const Tab& tab, int digit, bool endo, bool neg) {
  if (digit == 0) return;
  if constexpr (Tab::kAffine) {
    typename C::Aff a;
    tab.get(a, (digit < 0 ? -digit : digit) - 1);
    if (C::K::is_zero(a.x) && C::K::is_zero(a.y)) return;  // the identity (normalize_tables)
    if (endo) {
      fe8 beta;
      beta.v[0] = 0xd782e155u; beta.v[1] = 0x71930c11u; beta.v[2] = 0xffbe3323u; beta.v[3] = 0xa6bb947cu;
      beta.v[4] = 0xd4741444u; beta.v[5] = 0xaa303344u; beta.v[6] = 0x26594943u; beta.v[7] = 0x2c3b3f0du;
      C::K::mul(a.x, a.x, beta);
    }
    if (neg != (digit < 0)) C::K::neg(a.y, a.y);
    C::madd(acc, a);
    return;
  }
  typename C::Cached t;
  tab.get(t, (digit < 0 ? -digit : digit) - 1);
  if (endo) {
    fe8 beta;
    beta.v[0] = 0xd782e155u; beta.v[1] = 0x71930c11u; beta.v[2] = 0xffbe3323u; beta.v[3] = 0xa6bb947cu;
    beta.v[4] = 0xd4741444u; beta.v[5] = 0xaa303344u; beta.v[6] = 0x26594943u; beta.v[7] = 0x2c3b3f0du;
    C::K::mul(t.X, t.X, beta);
  }
  if (neg != (digit < 0)) C::cached_neg(t);
  C::add_cached(acc, t);
}

// digit of window i (0..32) of a recoded half-scalar
ARK_D int glv_digit(const uint32_t* kk, int i) { return (int)window4(kk, i) - 8; }
// k' = k + (0x8 in each of the 33 low nibbles).  |k| < 2^128 (glv_decompose_bn254), so k' < 2^128 + 0.54 * 2^132 < 2^132: the
// 33 nibbles hold all of k' and there is no carry digit.
ARK_D void glv_recode(uint32_t* kk, const uint32_t* k) {
  uint64_t c = 0;
  for (int j = 0; j < 4; j++) {
    c += (uint64_t)k[j] + 0x88888888u;
    kk[j] = (uint32_t)c;
    c >>= 32;
  }
  kk[4] = (uint32_t)(c + k[4] + 0x8u);
}

// CODE SIZE is the first-order concern in these loops: every field multiplication is ~110-300 inlined instructions, a point
// addition 900 (Edwards) to 3500 (Jacobian), and a hot loop that does not fit the instruction caches stalls on instruction
// fetch (ncu `stalled_no_instruction` was the top stall of every variant with several inlined additions per window;
// profiles/r02e_*).  Each loop below therefore contains exactly ONE copy of the doubling and ONE copy of the addition: the
// per-window additions are iterations of an inner `unroll 1` loop that only selects operands.
template <class C, class Tab>
ARK_D void var_mul_glv(typename C::Pt& acc, const Tab& tab, const uint32_t* k) {
  uint32_t k1[5], k2[5], r[2][5];
  bool n[2];
  glv_decompose_bn254(k, k1, n[0], k2, n[1]);
  glv_recode(r[0], k1);
  glv_recode(r[1], k2);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int i = kGlvWindows - 1; i >= 0; i--) {
    prefetch_signed(tab, glv_digit(r[0], i));
    prefetch_signed(tab, glv_digit(r[1], i));
    if (i != kGlvWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (int j = 0; j < 4; j++) C::dbl(acc, j == 3);
    }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int t = 0; t < 2; t++) glv_add<C>(acc, tab, glv_digit(r[t], i), t == 1, n[t]);
  }
}

template <class C, class Tab>
ARK_D void var_mul2_glv(typename C::Pt& acc0, typename C::Pt& acc1, const Tab& tab, const uint32_t* ka, const uint32_t* kb) {
  uint32_t a1[5], a2[5], b1[5], b2[5], t[5];
  bool na1, na2, nb1, nb2;
  glv_decompose_bn254(ka, a1, na1, a2, na2);
  glv_decompose_bn254(kb, b1, nb1, b2, nb2);
  glv_recode(t, a1); ARK_UNROLL for (int j = 0; j < 5; j++) a1[j] = t[j];
  glv_recode(t, a2); ARK_UNROLL for (int j = 0; j < 5; j++) a2[j] = t[j];
  glv_recode(t, b1); ARK_UNROLL for (int j = 0; j < 5; j++) b1[j] = t[j];
  glv_recode(t, b2); ARK_UNROLL for (int j = 0; j < 5; j++) b2[j] = t[j];
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int i = kGlvWindows - 1; i >= 0; i--) {
    const int da1 = glv_digit(a1, i), db1 = glv_digit(b1, i), da2 = glv_digit(a2, i), db2 = glv_digit(b2, i);
    prefetch_signed(tab, da1);
    prefetch_signed(tab, db1);
    prefetch_signed(tab, da2);
    prefetch_signed(tab, db2);
    if (i != kGlvWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (int j = 0; j < 4; j++) {
        C::dbl(acc0, j == 3);
        C::dbl(acc1, j == 3);
      }
    }
    glv_add<C>(acc0, tab, da1, false, na1);
    glv_add<C>(acc1, tab, db1, false, nb1);
    glv_add<C>(acc0, tab, da2, true, na2);
    glv_add<C>(acc1, tab, db2, true, nb2);
  }
}

// ----------------------------------------------------------------------------------------------
// Two multiplications of the SAME point (share and MAC passes of a gate): cut every scalar into C::kSplitParts pieces of W
// windows, k = sum_p 2^(4 W p) k_p, and give piece p its own table of P_p = 2^(4 W p) P.  The doublings that produce the P_p
// are paid once, and each pass then runs over W windows with one addition per piece per window.  With D scalar bits and s
// pieces that is D (s + 1) / s doublings for the two passes instead of 2 D, and 7 s more additions for the tables:
//   Curve25519 (D = 256, no endomorphism): s = 2: 128 + 2 x 124 doublings, s = 4: 192 + 2 x 60 — 8 % fewer field
//   multiplications for the gate than s = 2 (s = 8 gives the saving back to its eight tables);
//   BN254 (33-window GLV halves, four sub-scalars already): s = 2: 68 + 2 x 64 instead of 2 x 128; a further cut gains nothing.
// ----------------------------------------------------------------------------------------------
// ARK_PHASE_SYNC: the warps of a block re-converge after every point operation of the two-pass loops.  The loop body (one
// doubling + one addition, 35 KB of SASS on Curve25519, 110 KB on BN254) is larger than the SM's 32 KB instruction cache and is
// walked cyclically, the worst case for an LRU cache: every line misses, for every warp.  Keeping the block's warps within one
// operation of each other makes one fetch from L2 serve all of them.  The callers make the trip counts uniform per block.
#if defined(__CUDA_ARCH__) && !defined(ARK_NO_PHASE_SYNC)
#define ARK_PHASE_SYNC() __syncthreads()
#else
#define ARK_PHASE_SYNC() do { } while (0)
#endif

constexpr int kMaxSplitParts = 4;
constexpr int kGlvSplitWindows = (kGlvWindows + 1) / 2;  // 17 windows = 68 bits

// P' = 2^(4 * windows) P
template <class C>
ARK_D void shift_windows(typename C::Pt& p, int windows) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int j = 4 * windows - 1; j >= 0; j--) C::dbl(p, j == 0);
}
template <class C>
ARK_D int split_shift_windows() { return C::kGlv ? kGlvSplitWindows : kWindows / C::kSplitParts; }

// acc = k * P given the tables of the P_p (tab.sub(p)); acc must be the identity on entry
template <class C, class Tab>
ARK_D void var_mul_split(typename C::Pt& acc, const Tab& tab, const uint32_t* k) {
  if constexpr (C::kGlv) {
    const Tab lo = tab, hi = tab.sub(1);
    uint32_t k1[5], k2[5], r[2][5];
    bool n[2];
    glv_decompose_bn254(k, k1, n[0], k2, n[1]);
    glv_recode(r[0], k1);
    glv_recode(r[1], k2);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = kGlvSplitWindows - 1; i >= 0; i--) {
      const int ih = i + kGlvSplitWindows;  // the top half has one window fewer: window 33 does not exist
      prefetch_signed(lo, glv_digit(r[0], i));
      prefetch_signed(lo, glv_digit(r[1], i));
      if (ih < kGlvWindows) {
        prefetch_signed(hi, glv_digit(r[0], ih));
        prefetch_signed(hi, glv_digit(r[1], ih));
      }
      if (i != kGlvSplitWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) {
          ARK_PHASE_SYNC();
          C::dbl(acc, j == 3);
        }
      }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (int t = 0; t < 4; t++) {  // (k1, lo) (k2, lo) (k1, hi) (k2, hi): one inlined addition serves all four
        const int w = (t & 2) ? ih : i;
        const int digit = w < kGlvWindows ? glv_digit(r[t & 1], w) : 0;
        ARK_PHASE_SYNC();
        glv_add<C>(acc, (t & 2) ? hi : lo, digit, (t & 1) != 0, n[t & 1]);
      }
    }
  } else {
    constexpr int P = C::kSplitParts, W = kWindows / P;
    uint32_t kk[8];
    signed_recode(kk, k, 8);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = W - 1; i >= 0; i--) {
      ARK_UNROLL for (int t = 0; t < P; t++) prefetch_signed(tab.sub(t), (int)window4(kk, i + t * W) - 8);
      if (i != W - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) {
          ARK_PHASE_SYNC();
          C::dbl(acc, j == 3);
        }
      }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (int t = 0; t < P; t++) {  // one inlined addition serves every piece
        ARK_PHASE_SYNC();
        add_signed<C>(acc, tab.sub(t), (int)window4(kk, i + t * W) - 8);
      }
    }
  }
}

// (acc0, acc1) = (ka * P, kb * P) in lock-step (C::kDualChain), same tables
template <class C, class Tab>
ARK_D void var_mul2_split(typename C::Pt& acc0, typename C::Pt& acc1, const Tab& tab, const uint32_t* ka, const uint32_t* kb) {
  if constexpr (C::kGlv) {
    const Tab lo = tab, hi = tab.sub(1);
    uint32_t a1[5], a2[5], b1[5], b2[5], t[5];
    bool na1, na2, nb1, nb2;
    glv_decompose_bn254(ka, a1, na1, a2, na2);
    glv_decompose_bn254(kb, b1, nb1, b2, nb2);
    glv_recode(t, a1); ARK_UNROLL for (int j = 0; j < 5; j++) a1[j] = t[j];
    glv_recode(t, a2); ARK_UNROLL for (int j = 0; j < 5; j++) a2[j] = t[j];
    glv_recode(t, b1); ARK_UNROLL for (int j = 0; j < 5; j++) b1[j] = t[j];
    glv_recode(t, b2); ARK_UNROLL for (int j = 0; j < 5; j++) b2[j] = t[j];
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = kGlvSplitWindows - 1; i >= 0; i--) {
      const bool has_hi = i + kGlvSplitWindows < kGlvWindows;
      const int ih = i + kGlvSplitWindows;
      const int la1 = glv_digit(a1, i), la2 = glv_digit(a2, i), lb1 = glv_digit(b1, i), lb2 = glv_digit(b2, i);
      const int ha1 = has_hi ? glv_digit(a1, ih) : 0, ha2 = has_hi ? glv_digit(a2, ih) : 0;
      const int hb1 = has_hi ? glv_digit(b1, ih) : 0, hb2 = has_hi ? glv_digit(b2, ih) : 0;
      prefetch_signed(lo, la1); prefetch_signed(lo, lb1); prefetch_signed(lo, la2); prefetch_signed(lo, lb2);
      prefetch_signed(hi, ha1); prefetch_signed(hi, hb1); prefetch_signed(hi, ha2); prefetch_signed(hi, hb2);
      if (i != kGlvSplitWindows - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) {
          C::dbl(acc0, j == 3);
          C::dbl(acc1, j == 3);
        }
      }
      glv_add<C>(acc0, lo, la1, false, na1);
      glv_add<C>(acc1, lo, lb1, false, nb1);
      glv_add<C>(acc0, lo, la2, true, na2);
      glv_add<C>(acc1, lo, lb2, true, nb2);
      glv_add<C>(acc0, hi, ha1, false, na1);
      glv_add<C>(acc1, hi, hb1, false, nb1);
      glv_add<C>(acc0, hi, ha2, true, na2);
      glv_add<C>(acc1, hi, hb2, true, nb2);
    }
  } else {
    constexpr int P = C::kSplitParts, W = kWindows / P;
    uint32_t kk0[8], kk1[8];
    signed_recode(kk0, ka, 8);
    signed_recode(kk1, kb, 8);
#if defined(__CUDACC__)
#pragma unroll 1
#endif
    for (int i = W - 1; i >= 0; i--) {
      ARK_UNROLL for (int t = 0; t < P; t++) {
        prefetch_signed(tab.sub(t), (int)window4(kk0, i + t * W) - 8);
        prefetch_signed(tab.sub(t), (int)window4(kk1, i + t * W) - 8);
      }
      if (i != W - 1) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
        for (int j = 0; j < 4; j++) {
          C::dbl(acc0, j == 3);
          C::dbl(acc1, j == 3);
        }
      }
#if defined(__CUDACC__)
#pragma unroll 1
#endif
      for (int t = 0; t < P; t++) {
        add_signed<C>(acc0, tab.sub(t), (int)window4(kk0, i + t * W) - 8);
        add_signed<C>(acc1, tab.sub(t), (int)window4(kk1, i + t * W) - 8);
      }
    }
  }
}

// Fixed base: 12-bit windows, gtab[j * 4096 + w] = w * 4096^j * G for j < 22, 1 <= w <= 4095 (entry 0 unused): no doublings
// and at most 22 mixed additions per multiplication (8-bit windows: 32; the two fixed-base multiplications of a point Beaver
// gate were 16 % of its field multiplications on BN254).  The table (22 x 4096 affine entries, 5.8 MB on BN254, 8.7 MB on
// Curve25519) lives in global memory, is built once per context and is served from L2.
constexpr int kFixBits = 12;
constexpr int kFixWindows = (256 + kFixBits - 1) / kFixBits;
constexpr int kFixEntries = 1 << kFixBits;
ARK_D uint32_t window_fix(const uint32_t* k, int j) {
  const int bit = kFixBits * j, w = bit >> 5, sh = bit & 31;
  uint32_t v = k[w] >> sh;
  if (sh + kFixBits > 32 && w + 1 < 8) v |= k[w + 1] << (32 - sh);
  return v & (uint32_t)(kFixEntries - 1);
}

// acc += k * G
template <class C>
ARK_D void fix_mul_acc(typename C::Pt& acc, const typename C::Aff* gtab, const uint32_t* k) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int j = 0; j < kFixWindows; j++) {
    const uint32_t w = window_fix(k, j);
    if (w) C::madd(acc, gtab[j * kFixEntries + w]);
  }
}

// One table entry: out = w * 2^(kFixBits j) * G (w != 0), by kFixBits * j doublings of G and a double-and-add over the bits of w.
template <class C>
ARK_D void build_gtab_entry(typename C::Aff& out, int j, uint32_t w) {
  typename C::Pt base, acc;
  C::set_generator(base);
  for (int i = 0; i < kFixBits * j; i++) C::dbl(base);
  C::set_identity(acc);
  for (int b = kFixBits - 1; b >= 0; b--) {
    C::dbl(acc);
    if ((w >> b) & 1u) C::add(acc, base);
  }
  C::to_aff(out, acc);
}

}  // namespace ark
