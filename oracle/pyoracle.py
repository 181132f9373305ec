"""CPU oracle (Python big-int) for the ark-mpc online-phase hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`ark_mpc_b200/`) may
import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may use `oracle/`.

What it restates (paths relative to /root/reference/online-phase/src):

* prime-field arithmetic of `Scalar<C>` (algebra/scalar/scalar.rs:210-286), which
  forwards to the third-party crate `ark-ff 0.4` (`Fp256<MontBackend<_,4>>`,
  online-phase/Cargo.toml:91; NOT vendored in /root/reference, no Cargo.lock).
  The published algorithm: elements are stored as canonical Montgomery residues
  a*2^256 mod p in four little-endian u64 limbs.  Because a canonical residue is
  unique, any exact implementation is bit-identical to arkworks once that
  representation is fixed; Python ints give the exact answer.
* `ScalarShare` algebra (algebra/scalar/share.rs:74-131),
* the authenticated Beaver multiplication `batch_mul`
  (algebra/scalar/authenticated_scalar.rs:848-879, single-gate form :799-843),
* `open_batch` (:129-172), the MAC check of `open_authenticated_batch`
  (:278-354, :201-220), the hash commitment (commitment.rs:63-89),
* input sharing `batch_share_scalar` (fabric.rs:578-600),
* the mock preprocessing source `PartyIDBeaverSource` (offline_prep.rs:88-170),
* curve-group arithmetic behind `CurvePoint<C>` (algebra/curve/curve.rs:194-517) and
  the point Beaver multiplication (algebra/curve/authenticated_curve.rs:682-714);
  group law from `ark-ec 0.4` (not vendored) restated in affine form, which is the
  canonical representation parity is defined on (projective coordinates are equal
  only up to scaling, curve.rs:46).

Pinning (see tests/test_oracle.py): arkworks' published BN254 Fr Montgomery
constants (R, R2, INV), the reference's fixed-value tests (PartyIDBeaverSource
2*3=6 with key 1, share-and-open 0/1, xor -> 0), EIP-196 2*G on BN254 G1 and the
RFC 8032 test-1 public key on Ed25519.  Curve25519 is never instantiated by any
reference test ("parity unpinned" against the reference for that curve; pinned
against the RFCs instead).
"""
from __future__ import annotations

import hashlib
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

# ---------------------------------------------------------------------------
# Fields
# ---------------------------------------------------------------------------

R_BITS = 256
R = 1 << R_BITS
MASK64 = (1 << 64) - 1


@dataclass(frozen=True)
class Field:
    """A 256-bit-limbed prime field in arkworks' Montgomery memory image."""

    name: str
    p: int

    @property
    def r(self) -> int:  # R mod p  (Montgomery form of 1)
        return R % self.p

    @property
    def r2(self) -> int:  # R^2 mod p
        return (R * R) % self.p

    @property
    def inv64(self) -> int:  # -p^{-1} mod 2^64
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        return self.inv64 & 0xFFFFFFFF

    @property
    def rinv(self) -> int:
        return pow(R, -1, self.p)

    @property
    def bits(self) -> int:
        return self.p.bit_length()

    @property
    def n_bytes(self) -> int:  # scalar.rs:118-127 `n_bytes_field`
        return (self.bits + 7) // 8

    # Montgomery memory image <-> integers
    def to_mont(self, x: int) -> int:
        return (x % self.p) * R % self.p

    def from_mont(self, xm: int) -> int:
        return xm * self.rinv % self.p

    def limbs(self, x: int) -> Tuple[int, int, int, int]:
        """Montgomery image of `x` as 4 LE u64 limbs (the Rust memory image)."""
        m = self.to_mont(x)
        return tuple((m >> (64 * i)) & MASK64 for i in range(4))

    def from_limbs(self, l: Sequence[int]) -> int:
        m = sum(int(v) << (64 * i) for i, v in enumerate(l))
        assert m < self.p, "non-canonical Montgomery residue"
        return self.from_mont(m)

    # scalar.rs:118-127 (`to_bytes_be`) / :107-111 (`from_be_bytes_mod_order`)
    def to_bytes_be(self, x: int) -> bytes:
        return (x % self.p).to_bytes(self.n_bytes, "big")

    def from_be_bytes_mod_order(self, b: bytes) -> int:
        return int.from_bytes(b, "big") % self.p


BN254_FR = Field("bn254_fr", 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001)
BN254_FQ = Field("bn254_fq", 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47)
CURVE25519_FR = Field("curve25519_fr", (1 << 252) + 27742317777372353535851937790883648493)
CURVE25519_FQ = Field("curve25519_fq", (1 << 255) - 19)

FIELDS = {f.name: f for f in (BN254_FR, CURVE25519_FR, BN254_FQ, CURVE25519_FQ)}

PARTY0, PARTY1 = 0, 1

# ---------------------------------------------------------------------------
# ScalarShare algebra  (share.rs:32-131).  Values are plain ints mod p.
# ---------------------------------------------------------------------------

Share = Tuple[int, int]  # (share, mac)


def share_add(F: Field, a: Share, b: Share) -> Share:  # share.rs:85-91
    return ((a[0] + b[0]) % F.p, (a[1] + b[1]) % F.p)


def share_neg(F: Field, a: Share) -> Share:  # share.rs:115-121
    return ((-a[0]) % F.p, (-a[1]) % F.p)


def share_sub(F: Field, a: Share, b: Share) -> Share:  # share.rs:95-101: self + (-rhs)
    return share_add(F, a, share_neg(F, b))


def share_mul_public(F: Field, a: Share, s: int) -> Share:  # share.rs:125-131
    return (a[0] * s % F.p, a[1] * s % F.p)


def share_add_public(F: Field, a: Share, v: int, mac_key: int, party: int) -> Share:
    """share.rs:74-77: only party 0 adds to the share; both update the MAC."""
    s = (a[0] + v) % F.p if party == PARTY0 else a[0]
    return (s, (a[1] + mac_key * v) % F.p)


def share_sub_public(F: Field, a: Share, v: int, mac_key: int, party: int) -> Share:
    return share_add_public(F, a, (-v) % F.p, mac_key, party)  # share.rs:80-82


def share_sum(F: Field, xs: Sequence[Share]) -> Share:  # share.rs:104-111
    return (sum(x[0] for x in xs) % F.p, sum(x[1] for x in xs) % F.p)


# ---------------------------------------------------------------------------
# Beaver multiplication, per party  (authenticated_scalar.rs:848-879)
# ---------------------------------------------------------------------------


def beaver_mask(F: Field, x: Sequence[Share], y: Sequence[Share], a: Sequence[Share],
                b: Sequence[Share]) -> Tuple[List[int], List[int]]:
    """Own share components of d=[x-a], e=[y-b] (what `open_batch` sends, :141-145).

    The reference's `batch_sub` (:662-688) also computes the MAC halves of the
    masked values; `open_batch` never reads them, so they do not appear here.
    """
    d = [share_sub(F, xi, ai)[0] for xi, ai in zip(x, a)]
    e = [share_sub(F, yi, bi)[0] for yi, bi in zip(y, b)]
    return d, e


def open_add(F: Field, mine: Sequence[int], peer: Sequence[int]) -> List[int]:
    return [(m + q) % F.p for m, q in zip(mine, peer)]  # :166-168


def beaver_recombine(F: Field, party: int, mac_key: int, d: Sequence[int], e: Sequence[int],
                     a: Sequence[Share], b: Sequence[Share], c: Sequence[Share]) -> List[Share]:
    """[x*y] = de + d[b] + e[a] + [c] in the reference's unfused op order (:871-878)."""
    out = []
    for di, ei, ai, bi, ci in zip(d, e, a, b, c):
        de = di * ei % F.p                                  # scalar_result.rs:257-278
        db = share_mul_public(F, bi, di)                    # :872
        ea = share_mul_public(F, ai, ei)                    # :873
        de_plus_db = share_add_public(F, db, de, mac_key, party)  # :876
        ea_plus_c = share_add(F, ea, ci)                    # :877
        out.append(share_add(F, de_plus_db, ea_plus_c))     # :878
    return out


def mac_check_shares(F: Field, mac_key: int, opened: Sequence[int], shares: Sequence[Share]) -> List[int]:
    return [(mac_key * v - s[1]) % F.p for v, s in zip(opened, shares)]  # :299-311


def hash_commit(F: Field, values: Sequence[int], blinder: int) -> int:
    """commitment.rs:63-89: SHA3-256(values BE || blinder BE) reduced BE mod p."""
    h = hashlib.sha3_256()
    for v in values:
        h.update(F.to_bytes_be(v))
    h.update(F.to_bytes_be(blinder))
    return F.from_be_bytes_mod_order(h.digest())


def batch_verify_mac_check(F: Field, mine: Sequence[int], peer: Sequence[int], peer_blinder: int,
                           peer_commitment: int) -> bool:  # :201-220
    if hash_commit(F, peer, peer_blinder) != peer_commitment:
        return False
    return all((m + q) % F.p == 0 for m, q in zip(mine, peer))


# ---------------------------------------------------------------------------
# Preprocessing sources
# ---------------------------------------------------------------------------


class PartyIDBeaverSource:
    """offline_prep.rs:88-170: a=2, b=3, c=6; [a]=(1,1) [b]=(3,0) [c]=(2,4); key share = party id."""

    def __init__(self, F: Field, party: int):
        assert party in (0, 1)
        self.F, self.party = F, party

    def get_mac_key_share(self) -> int:
        return self.party

    def next_triplet_batch(self, n: int):
        key = self.party
        if self.party == 0:
            a, b, c = 1, 3, 2
        else:
            a, b, c = 1, 0, 4
        return ([(a, key * 2)] * n, [(b, key * 3)] * n, [(c, key * 6)] * n)

    def next_local_input_mask_batch(self, n: int):
        v = 3
        return [v] * n, [(self.party * v, self.party * v)] * n

    def next_counterparty_input_mask_batch(self, n: int):
        v = 3 * self.party
        return [(v, self.party * v)] * n

    def next_shared_bit_batch(self, n: int):
        return [(self.party, self.party)] * n


def split(F: Field, v: int, rng: random.Random) -> Tuple[int, int]:
    s0 = rng.randrange(F.p)
    return s0, (v - s0) % F.p


def authenticated_split(F: Field, v: int, key: int, rng: random.Random) -> Tuple[Share, Share]:
    """Additively share v and key*v (what a correct SPDZ offline phase hands out)."""
    s0, s1 = split(F, v, rng)
    m0, m1 = split(F, key * v % F.p, rng)
    return (s0, m0), (s1, m1)


class RandomBeaverSource:
    """Correct random triples under a random MAC key, as offline-phase `mock_lowgear_with_triples`
    (/root/reference/offline-phase/src/lib.rs:157-179) fabricates them from plaintext."""

    def __init__(self, F: Field, seed: int):
        self.F = F
        self.rng = random.Random(seed)
        self.key_shares = (self.rng.randrange(F.p), self.rng.randrange(F.p))
        self.key = sum(self.key_shares) % F.p

    def triples(self, n: int):
        F, rng = self.F, self.rng
        out = ([[], [], []], [[], [], []])
        for _ in range(n):
            a, b = rng.randrange(F.p), rng.randrange(F.p)
            for k, v in enumerate((a, b, a * b % F.p)):
                s0, s1 = authenticated_split(F, v, self.key, rng)
                out[0][k].append(s0)
                out[1][k].append(s1)
        return out

    def share_values(self, vals: Sequence[int]):
        p0, p1 = [], []
        for v in vals:
            s0, s1 = authenticated_split(self.F, v, self.key, self.rng)
            p0.append(s0)
            p1.append(s1)
        return p0, p1


# ---------------------------------------------------------------------------
# Two-party drivers (the `execute_mock_mpc` shape, lib.rs:116-201)
# ---------------------------------------------------------------------------


def two_party_batch_mul(F: Field, keys: Tuple[int, int], x, y, trip):
    """x, y: ([shares p0], [shares p1]); trip: per party (a, b, c) share lists.

    Returns per-party output shares and the opened (d, e)."""
    de = [beaver_mask(F, x[p], y[p], trip[p][0], trip[p][1]) for p in (0, 1)]
    d = open_add(F, de[0][0], de[1][0])
    e = open_add(F, de[0][1], de[1][1])
    outs = [beaver_recombine(F, p, keys[p], d, e, *trip[p]) for p in (0, 1)]
    return outs, (d, e), de


def open_shares(F: Field, s0: Sequence[Share], s1: Sequence[Share]) -> List[int]:
    return [(u[0] + v[0]) % F.p for u, v in zip(s0, s1)]


def two_party_open_authenticated(F: Field, keys, s0: Sequence[Share], s1: Sequence[Share],
                                 blinders=(11, 13)) -> Tuple[List[int], bool]:
    """:278-354 for both parties; returns (opened values, both MAC checks passed)."""
    opened = open_shares(F, s0, s1)
    chk = [mac_check_shares(F, keys[p], opened, s) for p, s in ((0, s0), (1, s1))]
    comm = [hash_commit(F, chk[p], blinders[p]) for p in (0, 1)]
    ok0 = batch_verify_mac_check(F, chk[0], chk[1], blinders[1], comm[1])
    ok1 = batch_verify_mac_check(F, chk[1], chk[0], blinders[0], comm[0])
    return opened, ok0 and ok1


def two_party_share_scalars(F: Field, vals: Sequence[int], sender: int, src0, src1, keys):
    """fabric.rs:578-600 for both parties with the given preprocessing sources."""
    n = len(vals)
    srcs = (src0, src1)
    masks, sender_shares = srcs[sender].next_local_input_mask_batch(n)
    other_shares = srcs[1 - sender].next_counterparty_input_mask_batch(n)
    masked = [(v - m) % F.p for v, m in zip(vals, masks)]
    mask_shares = [None, None]
    mask_shares[sender], mask_shares[1 - sender] = sender_shares, other_shares
    return tuple([share_add_public(F, s, mv, keys[p], p) for s, mv in zip(mask_shares[p], masked)]
                 for p in (0, 1))


# ---------------------------------------------------------------------------
# Curve groups (affine; None = identity).  ark-ec 0.4 group law, canonical form.
# ---------------------------------------------------------------------------


@dataclass(frozen=True)
class Curve:
    name: str
    fq: Field
    fr: Field
    kind: str            # "sw" (y^2 = x^3 + b) or "te" (a x^2 + y^2 = 1 + d x^2 y^2)
    a: int
    b_or_d: int
    gx: int
    gy: int

    @property
    def generator(self):
        return (self.gx, self.gy)

    @property
    def identity(self):
        return None if self.kind == "sw" else (0, 1)

    def is_on_curve(self, P) -> bool:
        q = self.fq.p
        if self.kind == "sw":
            if P is None:
                return True
            x, y = P
            return (y * y - x * x * x - self.a * x - self.b_or_d) % q == 0
        x, y = P
        return (self.a * x * x + y * y - 1 - self.b_or_d * x * x * y * y) % q == 0

    def neg(self, P):
        if self.kind == "sw":
            return None if P is None else (P[0], (-P[1]) % self.fq.p)
        return ((-P[0]) % self.fq.p, P[1])

    def add(self, P, Q):
        q = self.fq.p
        if self.kind == "sw":
            if P is None:
                return Q
            if Q is None:
                return P
            x1, y1 = P
            x2, y2 = Q
            if x1 == x2:
                if (y1 + y2) % q == 0:
                    return None
                lam = (3 * x1 * x1 + self.a) * pow(2 * y1, -1, q) % q
            else:
                lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
            x3 = (lam * lam - x1 - x2) % q
            return (x3, (lam * (x1 - x3) - y1) % q)
        # twisted Edwards, complete when a is a square and d a non-square (Ed25519: a=-1)
        x1, y1 = P
        x2, y2 = Q
        t = self.b_or_d * x1 * x2 * y1 * y2 % q
        x3 = (x1 * y2 + y1 * x2) * pow(1 + t, -1, q) % q
        y3 = (y1 * y2 - self.a * x1 * x2) * pow(1 - t, -1, q) % q
        return (x3, y3)

    def sub(self, P, Q):
        return self.add(P, self.neg(Q))

    def mul(self, P, k: int):
        """curve.rs:403-409 -> ark-ec `Projective * ScalarField` (double-and-add)."""
        k %= self.fr.p
        acc = self.identity
        while k:
            if k & 1:
                acc = self.add(acc, P)
            P = self.add(P, P)
            k >>= 1
        return acc


BN254_G1 = Curve("bn254_g1", BN254_FQ, BN254_FR, "sw", 0, 3, 1, 2)

_ED_D = (-121665 * pow(121666, -1, CURVE25519_FQ.p)) % CURVE25519_FQ.p
_ED_GY = 4 * pow(5, -1, CURVE25519_FQ.p) % CURVE25519_FQ.p
_ED_GX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
CURVE25519_EDWARDS = Curve("curve25519_edwards", CURVE25519_FQ, CURVE25519_FR, "te",
                           CURVE25519_FQ.p - 1, _ED_D, _ED_GX, _ED_GY)

CURVES = {c.name: c for c in (BN254_G1, CURVE25519_EDWARDS)}

PointShare = Tuple[object, object]  # (share point, mac point)  curve/share.rs:25-30


def pshare_add(C: Curve, a: PointShare, b: PointShare) -> PointShare:  # curve/share.rs:66-72
    return (C.add(a[0], b[0]), C.add(a[1], b[1]))


def pshare_neg(C: Curve, a: PointShare) -> PointShare:
    return (C.neg(a[0]), C.neg(a[1]))


def pshare_sub(C: Curve, a: PointShare, b: PointShare) -> PointShare:
    return pshare_add(C, a, pshare_neg(C, b))


def pshare_mul_public(C: Curve, a: PointShare, s: int) -> PointShare:  # curve/share.rs:107-113
    return (C.mul(a[0], s), C.mul(a[1], s))


def pshare_add_public(C: Curve, a: PointShare, P, mac_key: int, party: int) -> PointShare:
    s = C.add(a[0], P) if party == PARTY0 else a[0]  # curve/share.rs:57-60
    return (s, C.add(a[1], C.mul(P, mac_key)))


def scalar_share_mul_point(C: Curve, s: Share, P) -> PointShare:  # scalar/share.rs:135-141
    return (C.mul(P, s[0]), C.mul(P, s[1]))


def point_beaver_mask(C: Curve, x: Sequence[Share], P: Sequence[PointShare], a, b):
    """authenticated_curve.rs:696-700: own components of d=[x-a] and E=[P - bG]."""
    G = C.generator
    bG = [scalar_share_mul_point(C, bi, G) for bi in b]           # :697 batch_mul_generator
    d = [share_sub(C.fr, xi, ai)[0] for xi, ai in zip(x, a)]      # :699
    E = [pshare_sub(C, Pi, bGi)[0] for Pi, bGi in zip(P, bG)]     # :700
    return d, E


def point_beaver_recombine(C: Curve, party: int, mac_key: int, d, E, a, b, c) -> List[PointShare]:
    """[x*P] = d*E + d[bG] + [a]E + [c]G  (authenticated_curve.rs:704-713)."""
    G = C.generator
    out = []
    for di, Ei, ai, bi, ci in zip(d, E, a, b, c):
        deG = C.mul(Ei, di)                                       # curve.rs:459-479
        dbG = pshare_mul_public(C, scalar_share_mul_point(C, bi, G), di)   # :706
        aeG = scalar_share_mul_point(C, ai, Ei)                   # curve.rs:483-517
        cG = scalar_share_mul_point(C, ci, G)                     # :708
        de_db = pshare_add_public(C, dbG, deG, mac_key, party)    # :710
        ae_c = pshare_add(C, aeG, cG)                             # :711
        out.append(pshare_add(C, de_db, ae_c))                    # :713
    return out


# ---------------------------------------------------------------------------
# Deterministic synthetic data (shared with the CUDA generator and the C oracle)
# ---------------------------------------------------------------------------


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def synth_element(F: Field, seed: int, index: int) -> int:
    """Uniform element of [0,p) in the MONTGOMERY IMAGE domain from (seed, index).

    Counter-based: limb j of attempt t is splitmix64(seed ^ splitmix64(index*8 + j + 4*(t&1)) + t),
    top limb masked to the modulus' bit length, rejected while >= p.  Returns the raw
    residue m (the 4-limb memory image is m itself)."""
    top_mask = (1 << (F.bits - 192)) - 1
    t = 0
    while True:
        limbs = []
        for j in range(4):
            ctr = (index * 4 + j) & MASK64
            limbs.append(splitmix64((seed ^ splitmix64(ctr)) + (t * 0xD1342543DE82EF95 & MASK64) & MASK64))
        limbs[3] &= top_mask
        m = sum(v << (64 * i) for i, v in enumerate(limbs))
        if m < F.p:
            return m
        t += 1


# ---------------------------------------------------------------------------
# FFT on shares (share.rs:162-192 -> ark-poly Radix2EvaluationDomain) and batch inversion (scalar.rs:93-100)
# ---------------------------------------------------------------------------
BN254_FR_GENERATOR = 5       # ark-bn254 FrConfig (crate not vendored)
BN254_FR_TWO_ADICITY = 28


def root_of_unity(F: Field, n: int) -> int:
    """ark-ff `get_root_of_unity(n)`: TWO_ADIC_ROOT_OF_UNITY squared down to order n (n a power of two)."""
    assert F is BN254_FR and n & (n - 1) == 0 and n <= 1 << BN254_FR_TWO_ADICITY
    w = pow(BN254_FR_GENERATOR, (F.p - 1) >> BN254_FR_TWO_ADICITY, F.p)
    return pow(w, (1 << BN254_FR_TWO_ADICITY) // n, F.p)


def naive_dft(F: Field, xs: Sequence[int], inverse: bool = False) -> List[int]:
    """Definition of ark-poly's fft / ifft on a domain of size len(xs): X_j = sum_i x_i w^(ij); ifft uses w^-1 and n^-1."""
    n = len(xs)
    w = root_of_unity(F, n)
    if inverse:
        w = pow(w, -1, F.p)
    out = [sum(x * pow(w, i * j, F.p) for i, x in enumerate(xs)) % F.p for j in range(n)]
    if inverse:
        ninv = pow(n, -1, F.p)
        out = [v * ninv % F.p for v in out]
    return out


def batch_inverse(F: Field, xs: Sequence[int]) -> List[int]:
    """ark_ff::batch_inversion: zeros are left untouched."""
    return [pow(x, -1, F.p) if x % F.p else 0 for x in xs]
