"""Pins the CPU oracle (oracle/pyoracle.py exact big-int, oracle/ark_oracle.c C restatement).

The reference holds NO golden vectors for this path; its fixed-value checks are
restated here, plus published constants of the un-vendored dependencies (arkworks 0.4 BN254 Fr
Montgomery constants, EIP-196 2*G on BN254 G1, RFC 8032 test 1 on Ed25519)."""
import hashlib
import random

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyoracle as po

FR = [("bn254_fr", 0), ("curve25519_fr", 1)]
ALL = [("bn254_fr", 0), ("curve25519_fr", 1), ("bn254_fq", 2), ("curve25519_fq", 3)]


def test_arkworks_bn254_fr_constants():
    # ark-bn254 0.4 `FrConfig`: MODULUS, R, R2, INV as published
    F = po.BN254_FR
    assert F.p == 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert F.r == 0x0E0A77C19A07DF2F666EA36F7879462E36FC76959F60CD29AC96341C4FFFFFFB
    assert F.r2 == 0x0216D0B17F4E44A58C49833D53BB808553FE3AB1E35C59E31BB8E645AE216DA7
    assert F.inv64 == 0xC2E1F593EFFFFFFF
    assert F.limbs(1) == (0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F)


def test_curve25519_constants():
    assert po.CURVE25519_FR.p == 2**252 + 27742317777372353535851937790883648493
    assert po.CURVE25519_FQ.p == 2**255 - 19
    assert po.CURVE25519_FQ.r == 38 and po.CURVE25519_FQ.r2 == 1444


@pytest.mark.parametrize("name,fid", ALL)
def test_c_constants_match_python(name, fid):
    F = po.FIELDS[name]
    k = co.field_constants(fid)
    tolimbs = lambda v: [(v >> (64 * i)) & po.MASK64 for i in range(4)]
    assert k["p"] == tolimbs(F.p) and k["r"] == tolimbs(F.r) and k["r2"] == tolimbs(F.r2) and k["inv"] == F.inv64


def _edge_and_random(F, rng, n):
    edge = [0, 1, 2, F.p - 1, F.p - 2, (F.p - 1) // 2, (F.p + 1) // 2, F.r, F.rinv, (1 << 64) - 1, 1 << 64,
            (1 << 128) - 1, (1 << 192) + 5, F.p >> 1]
    return [e % F.p for e in edge] + [rng.randrange(F.p) for _ in range(n)]


@pytest.mark.parametrize("name,fid", ALL)
def test_c_field_ops_match_bigint(name, fid):
    F = po.FIELDS[name]
    rng = random.Random(fid + 7)
    a = _edge_and_random(F, rng, 300)
    b = list(reversed(_edge_and_random(F, rng, 300)))
    am = co.to_mont(fid, co.ints_to_limbs(a))
    bm = co.to_mont(fid, co.ints_to_limbs(b))
    assert co.limbs_to_ints(am) == [F.to_mont(v) for v in a]          # memory image = a*R mod p
    assert co.limbs_to_ints(co.from_mont(fid, am)) == a
    for op, ref in ((co.scalar_add, lambda u, v: (u + v) % F.p), (co.scalar_sub, lambda u, v: (u - v) % F.p),
                    (co.scalar_mul, lambda u, v: u * v % F.p)):
        got = co.limbs_to_ints(co.from_mont(fid, op(fid, am, bm)))
        assert got == [ref(u, v) for u, v in zip(a, b)]


@pytest.mark.parametrize("name,fid", FR)
def test_synth_generator_c_equals_python(name, fid):
    F = po.FIELDS[name]
    got = co.limbs_to_ints(co.synth(fid, 0xA11CE, 5, 64))
    want = [po.synth_element(F, 0xA11CE, 5 + i) for i in range(64)]
    assert got == want and all(v < F.p for v in got)
    assert len(set(got)) == 64


def _shares_to_aos(fid, F, shares):
    flat = []
    for s, m in shares:
        flat += [s, m]
    return co.to_mont(fid, co.ints_to_limbs(flat)).reshape(-1, 8)


def _aos_to_shares(fid, arr):
    v = co.limbs_to_ints(co.from_mont(fid, np.ascontiguousarray(arr).reshape(-1, 4)))
    return list(zip(v[0::2], v[1::2]))


@pytest.mark.parametrize("name,fid", FR)
@pytest.mark.parametrize("n", [1, 2, 33, 100])
def test_batch_mul_two_party_c_vs_python_and_plaintext(name, fid, n):
    """authenticated_scalar.rs test `test_batch_mul` shape: open(batch_mul(a,b)) == a*b, plus every
    per-party output limb of the C restatement equals the exact big-int oracle."""
    F = po.FIELDS[name]
    rng = random.Random(1000 * fid + n)
    src = po.RandomBeaverSource(F, seed=n)
    xv = [rng.randrange(F.p) for _ in range(n)]
    yv = [rng.randrange(F.p) for _ in range(n)]
    x, y = src.share_values(xv), src.share_values(yv)
    trip = src.triples(n)
    outs, (d, e), _ = po.two_party_batch_mul(F, src.key_shares, x, y, trip)
    assert po.open_shares(F, outs[0], outs[1]) == [u * v % F.p for u, v in zip(xv, yv)]
    assert [(s0[1] + s1[1]) % F.p for s0, s1 in zip(*outs)] == [src.key * u * v % F.p for u, v in zip(xv, yv)]
    aos = lambda sh: _shares_to_aos(fid, F, sh)
    keys = [co.to_mont(fid, co.ints_to_limbs([k]))[0] for k in src.key_shares]
    for threads in (1, 3):
        o0, o1, dc, ec = co.two_party_batch_mul(
            fid, threads, keys, (aos(x[0]), aos(x[1])), (aos(y[0]), aos(y[1])),
            (aos(trip[0][0]), aos(trip[1][0])), (aos(trip[0][1]), aos(trip[1][1])), (aos(trip[0][2]), aos(trip[1][2])))
        assert _aos_to_shares(fid, o0) == outs[0] and _aos_to_shares(fid, o1) == outs[1]
        assert co.limbs_to_ints(co.from_mont(fid, dc)) == d and co.limbs_to_ints(co.from_mont(fid, ec)) == e


@pytest.mark.parametrize("name,fid", FR)
def test_party_id_beaver_source_kat(name, fid):
    """offline_prep.rs:137-158 + integration/src/lowgear.rs:37-45: key=1, triple 2*3=6; share x,y via the
    mock input masks (fabric.rs:578-600), multiply, open authenticated."""
    F = po.FIELDS[name]
    s0, s1 = po.PartyIDBeaverSource(F, 0), po.PartyIDBeaverSource(F, 1)
    keys = (s0.get_mac_key_share(), s1.get_mac_key_share())
    assert sum(keys) % F.p == 1
    t0, t1 = s0.next_triplet_batch(4), s1.next_triplet_batch(4)
    for k, v in enumerate((2, 3, 6)):
        assert (t0[k][0][0] + t1[k][0][0]) % F.p == v and (t0[k][0][1] + t1[k][0][1]) % F.p == v
    xv, yv = [5, 0, F.p - 1, 12345], [7, 9, F.p - 1, 0]
    x = po.two_party_share_scalars(F, xv, 0, s0, s1, keys)
    y = po.two_party_share_scalars(F, yv, 1, s0, s1, keys)
    assert po.open_shares(F, *x) == xv and po.open_shares(F, *y) == yv
    outs, _, _ = po.two_party_batch_mul(F, keys, x, y, (t0, t1))
    opened, ok = po.two_party_open_authenticated(F, keys, outs[0], outs[1])
    assert ok and opened == [u * v % F.p for u, v in zip(xv, yv)]


def test_share_and_open_zero_one_and_xor():
    # integration/src/fabric.rs:15-32 (0 and 1) and authenticated_scalar.rs:1676-1688 (a xor a == 0)
    F = po.BN254_FR
    s0, s1 = po.PartyIDBeaverSource(F, 0), po.PartyIDBeaverSource(F, 1)
    keys = (0, 1)
    sh = po.two_party_share_scalars(F, [0, 1], 0, s0, s1, keys)
    assert po.open_shares(F, *sh) == [0, 1]
    a = po.two_party_share_scalars(F, [1], 0, s0, s1, keys)
    t = (s0.next_triplet_batch(1), s1.next_triplet_batch(1))
    ab, _, _ = po.two_party_batch_mul(F, keys, a, a, t)
    # xor = a + b - 2ab
    xor = [po.share_sub(F, po.share_add(F, a[p][0], a[p][0]), po.share_mul_public(F, ab[p][0], 2)) for p in (0, 1)]
    assert (xor[0][0] + xor[1][0]) % F.p == 0


@pytest.mark.parametrize("corrupt", ["mac", "share"])
def test_open_authenticated_detects_corruption(corrupt):
    # integration/src/authenticated_scalar.rs:49-75
    F = po.BN254_FR
    src = po.RandomBeaverSource(F, seed=3)
    s0, s1 = src.share_values([42, 43, 44])
    _, ok = po.two_party_open_authenticated(F, src.key_shares, s0, s1)
    assert ok
    bad = list(s0)
    bad[1] = (bad[1][0], (bad[1][1] + 1) % F.p) if corrupt == "mac" else ((bad[1][0] + 1) % F.p, bad[1][1])
    _, ok = po.two_party_open_authenticated(F, src.key_shares, bad, s1)
    assert not ok


@pytest.mark.parametrize("name,fid", FR)
def test_c_linear_gates_match_python(name, fid):
    F = po.FIELDS[name]
    rng = random.Random(99 + fid)
    n = 50
    rs = lambda: [(rng.randrange(F.p), rng.randrange(F.p)) for _ in range(n)]
    a, b = rs(), rs()
    a[0], b[0] = (0, 0), (0, F.p - 1)
    v = [rng.randrange(F.p) for _ in range(n)]
    v[1] = 0
    key = rng.randrange(F.p)
    A, B = _shares_to_aos(fid, F, a), _shares_to_aos(fid, F, b)
    V = co.to_mont(fid, co.ints_to_limbs(v))
    K = co.to_mont(fid, co.ints_to_limbs([key]))[0]
    assert _aos_to_shares(fid, co.batch_add(fid, A, B)) == [po.share_add(F, s, t) for s, t in zip(a, b)]
    assert _aos_to_shares(fid, co.batch_sub(fid, A, B)) == [po.share_sub(F, s, t) for s, t in zip(a, b)]
    assert _aos_to_shares(fid, co.batch_neg(fid, A)) == [po.share_neg(F, s) for s in a]
    assert _aos_to_shares(fid, co.batch_mul_public(fid, A, V)) == [po.share_mul_public(F, s, t) for s, t in zip(a, v)]
    for party in (0, 1):
        assert _aos_to_shares(fid, co.batch_add_public(fid, party, K, A, V)) == \
            [po.share_add_public(F, s, t, key, party) for s, t in zip(a, v)]
        assert _aos_to_shares(fid, co.batch_add_public(fid, party, K, A, V, sub=True)) == \
            [po.share_sub_public(F, s, t, key, party) for s, t in zip(a, v)]
    assert co.limbs_to_ints(co.from_mont(fid, co.mac_check(fid, K, V, A))) == po.mac_check_shares(F, key, v, a)
    assert _aos_to_shares(fid, co.share_sum(fid, A).reshape(1, 8)) == [po.share_sum(F, a)]


def test_hash_commitment_restatement():
    # commitment.rs:63-89: SHA3-256 over 32-byte BE values || blinder, reduced BE mod p
    F = po.BN254_FR
    vals, blinder = [1, F.p - 1, 2**200 + 3], 77
    raw = b"".join(v.to_bytes(32, "big") for v in vals) + blinder.to_bytes(32, "big")
    assert po.hash_commit(F, vals, blinder) == int.from_bytes(hashlib.sha3_256(raw).digest(), "big") % F.p


def test_bn254_g1_eip196_double():
    C = po.BN254_G1
    assert C.is_on_curve(C.generator)
    g2 = C.mul(C.generator, 2)
    assert g2 == (0x030644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD3,
                  0x15ED738C0E0A7C92E7845F96B2AE9C0A68A6A449E3538FC7FF3EBF7A5A18A2C4)
    assert C.mul(C.generator, C.fr.p) is None and C.mul(C.generator, C.fr.p - 1) == C.neg(C.generator)


def test_ed25519_rfc8032_test1_public_key():
    E = po.CURVE25519_EDWARDS
    assert E.is_on_curve(E.generator) and E.mul(E.generator, E.fr.p) == (0, 1)
    sk = bytes.fromhex("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60")
    a = bytearray(hashlib.sha512(sk).digest()[:32])
    a[0] &= 248; a[31] &= 127; a[31] |= 64
    A = E.mul(E.generator, int.from_bytes(a, "little"))
    enc = (A[1] | ((A[0] & 1) << 255)).to_bytes(32, "little").hex()
    assert enc == "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"


@pytest.mark.parametrize("cname", ["bn254_g1", "curve25519_edwards"])
def test_point_beaver_mul_identity(cname):
    """authenticated_curve.rs `test_multiplication`-shape: open([x]*[P]) == x*P, MACs consistent."""
    C = po.CURVES[cname]
    F = C.fr
    rng = random.Random(5)
    src = po.RandomBeaverSource(F, seed=8)
    n = 3
    xv = [rng.randrange(F.p) for _ in range(n)]
    Pv = [C.mul(C.generator, rng.randrange(F.p)) for _ in range(n)]
    x = src.share_values(xv)
    P = ([], [])
    for Pt in Pv:  # additive split of P and key*P
        r0 = C.mul(C.generator, rng.randrange(F.p))
        m0 = C.mul(C.generator, rng.randrange(F.p))
        P[0].append((r0, m0))
        P[1].append((C.sub(Pt, r0), C.sub(C.mul(Pt, src.key), m0)))
    trip = src.triples(n)
    masks = [po.point_beaver_mask(C, x[p], P[p], trip[p][0], trip[p][1]) for p in (0, 1)]
    d = po.open_add(F, masks[0][0], masks[1][0])
    E = [C.add(u, v) for u, v in zip(masks[0][1], masks[1][1])]
    outs = [po.point_beaver_recombine(C, p, src.key_shares[p], d, E, *trip[p]) for p in (0, 1)]
    for i in range(n):
        want = C.mul(Pv[i], xv[i])
        assert C.add(outs[0][i][0], outs[1][i][0]) == want
        assert C.add(outs[0][i][1], outs[1][i][1]) == C.mul(want, src.key)
