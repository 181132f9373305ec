"""Pins the C restatement of the curve / point-Beaver path (oracle/ark_oracle.c) against the exact affine big-int
oracle (oracle/pyoracle.py) and against external known answers (EIP-196 BN254 doubling, RFC 8032 base point order)."""
import random

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyoracle as po
from tests.util_curve import CURVE_BY_ID, affine_ints, points_from_affine, TwoPartyPointData

CURVES = [0, 1]


@pytest.mark.parametrize("cv", CURVES)
def test_generator_and_group_law(cv):
    Cv = CURVE_BY_ID[cv]
    g = co.pt_generator(cv).reshape(1, -1)
    assert affine_ints(cv, g)[0] == Cv.generator
    rng = random.Random(5 + cv)
    pts = [Cv.mul(Cv.generator, rng.randrange(1, Cv.fr.p)) for _ in range(4)] + [Cv.identity]
    A = points_from_affine(cv, pts, rng)
    for i, P in enumerate(pts):
        for j, Q in enumerate(pts):
            got = affine_ints(cv, co.pt_add(cv, A[i:i + 1], A[j:j + 1]))[0]
            assert got == Cv.add(P, Q)
            got = affine_ints(cv, co.pt_add(cv, A[i:i + 1], A[j:j + 1], sub=True))[0]
            assert got == Cv.sub(P, Q)


def test_known_answers():
    # EIP-196 / py_ecc: 2 * (1, 2) on BN254 G1
    two_g = affine_ints(0, co.pt_mul_generator(0, co.to_mont(0, co.ints_to_limbs([2]))))[0]
    assert two_g == (1368015179489954701390400359078579693043519447331113978918064868415326638035,
                     9918110051302171585080402603319702774565515993150576347155970296011118125764)
    # RFC 8032: the Ed25519 base point has order l
    l = po.CURVE25519_FR.p
    assert affine_ints(1, co.pt_mul_generator(1, co.to_mont(1, co.ints_to_limbs([l - 1]))))[0] == po.CURVE25519_EDWARDS.neg(po.CURVE25519_EDWARDS.generator)
    # RFC 8032 section 7.1 test 1: public key = clamp(sha512(sk)[:32]) * B
    import hashlib
    sk = bytes.fromhex("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60")
    h = bytearray(hashlib.sha512(sk).digest()[:32])
    h[0] &= 248; h[31] &= 127; h[31] |= 64
    s = int.from_bytes(h, "little") % l
    x, y = affine_ints(1, co.pt_mul_generator(1, co.to_mont(1, co.ints_to_limbs([s]))))[0]
    enc = (y | ((x & 1) << 255)).to_bytes(32, "little")
    assert enc.hex() == "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"


# RFC 8032 section 7.1, tests 1-3 and the 1023-byte-message test: (secret key, public key).  The public key is the compressed
# encoding of clamp(SHA-512(sk)[:32]) * B, so each vector pins the fixed-base multiplication, the group law and the base point of
# the Edwards oracle at a full-size scalar.
RFC8032_KEYS = [
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60", "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb", "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7", "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025"),
    ("f5e5767cf153319517630f226876b86c8160cc583bc013744c6bf255f5cc0ee5", "278117fc144c72340f67d0f2316e8386ceffbf2b2428c9c51fef7c597f1d426e"),
]


def _rfc8032_scalar(sk_hex):
    import hashlib

    h = bytearray(hashlib.sha512(bytes.fromhex(sk_hex)).digest()[:32])
    h[0] &= 248
    h[31] &= 127
    h[31] |= 64
    return int.from_bytes(h, "little")


@pytest.mark.parametrize("sk,pk", RFC8032_KEYS)
def test_rfc8032_public_keys(sk, pk):
    l = po.CURVE25519_FR.p
    s = _rfc8032_scalar(sk)
    for where, (x, y) in (("python oracle", po.CURVE25519_EDWARDS.mul(po.CURVE25519_EDWARDS.generator, s % l)),
                          ("C oracle", affine_ints(1, co.pt_mul_generator(1, co.to_mont(1, co.ints_to_limbs([s % l]))))[0])):
        assert (y | ((x & 1) << 255)).to_bytes(32, "little").hex() == pk, where


def test_bn254_g1_known_multiples():
    """EIP-196 precompile vectors (go-ethereum / py_ecc): small multiples of (1, 2), and the group order as published in EIP-196:
    r * G is the point at infinity, (r - 1) * G = -G = (1, p - 2)."""
    Cv = po.BN254_G1
    r, q = po.BN254_FR.p, Cv.fq.p
    assert r == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert q == 21888242871839275222246405745257275088696311157297823662689037894645226208583
    mul = lambda k: affine_ints(0, co.pt_mul_generator(0, co.to_mont(0, co.ints_to_limbs([k % r]))))[0]
    two_g = (1368015179489954701390400359078579693043519447331113978918064868415326638035,
             9918110051302171585080402603319702774565515993150576347155970296011118125764)
    three_g = (3353031288059533942658390886683067124040920775575537747144343083137631628272,
               19321533766552368860946552437480515441416830039777911637913418824951667761761)
    assert mul(2) == two_g == Cv.mul(Cv.generator, 2)
    assert mul(3) == three_g == Cv.add(two_g, Cv.generator)
    assert mul(r - 1) == (1, q - 2)
    assert Cv.mul(Cv.generator, r) is None
    # on-curve: y^2 = x^3 + 3
    for x, y in (two_g, three_g):
        assert (y * y - x * x * x - 3) % q == 0


@pytest.mark.parametrize("cv", CURVES)
def test_scalar_mul_matches_python(cv):
    Cv = CURVE_BY_ID[cv]
    rng = random.Random(9 + cv)
    r = Cv.fr.p
    scal = [0, 1, 2, r - 1, rng.randrange(r), rng.randrange(r)]
    P = Cv.mul(Cv.generator, rng.randrange(1, r))
    A = points_from_affine(cv, [P] * len(scal), rng)
    S = co.to_mont(co.CURVE_FR[cv], co.ints_to_limbs(scal))
    got = affine_ints(cv, co.pt_mul(cv, S, A))
    assert got == [Cv.mul(P, s) for s in scal]
    got = affine_ints(cv, co.pt_mul_generator(cv, S))
    assert got == [Cv.mul(Cv.generator, s) for s in scal]


@pytest.mark.parametrize("cv", CURVES)
def test_two_party_point_mul_matches_python(cv):
    """authenticated_curve.rs:682-714, both parties: C restatement == affine Python restatement; result opens to x*P."""
    Cv = CURVE_BY_ID[cv]
    D = TwoPartyPointData(cv, 5, seed=77 + cv)
    out0, out1, d_open, E_open = D.oracle_point_mul(threads=2)
    F = Cv.fr
    fr = co.CURVE_FR[cv]
    ints = lambda a: [F.from_mont(v) for v in co.limbs_to_ints(a)]
    sh = lambda aos_arr: list(zip(ints(aos_arr[:, :4]), ints(aos_arr[:, 4:])))
    w = co.point_words(cv)
    psh = lambda ps: list(zip(affine_ints(cv, ps[:, :w]), affine_ints(cv, ps[:, w:])))
    keys = [F.from_mont(co.limbs_to_ints(k)[0]) for k in D.keys]
    x = [sh(D.x[p]) for p in (0, 1)]
    a = [sh(D.a[p]) for p in (0, 1)]
    b = [sh(D.b[p]) for p in (0, 1)]
    c = [sh(D.c[p]) for p in (0, 1)]
    P = [psh(D.P[p]) for p in (0, 1)]
    masks = [po.point_beaver_mask(Cv, x[p], P[p], a[p], b[p]) for p in (0, 1)]
    d = po.open_add(F, masks[0][0], masks[1][0])
    E = [Cv.add(u, v) for u, v in zip(masks[0][1], masks[1][1])]
    assert ints(d_open) == d and affine_ints(cv, E_open) == E
    for p, out in ((0, out0), (1, out1)):
        want = po.point_beaver_recombine(Cv, p, keys[p], d, E, a[p], b[p], c[p])
        assert psh(out) == want
    # opens to x * P with a valid MAC
    key = sum(keys) % F.p
    opened = [Cv.add(u[0], v[0]) for u, v in zip(psh(out0), psh(out1))]
    macs = [Cv.add(u[1], v[1]) for u, v in zip(psh(out0), psh(out1))]
    xv = ints(D.xv)
    Pv = affine_ints(cv, D.Pv)
    assert opened == [Cv.mul(Pi, xi) for Pi, xi in zip(Pv, xv)]
    assert macs == [Cv.mul(Oi, key) for Oi in opened]
