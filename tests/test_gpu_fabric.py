"""Two-party tests through the host-side mirror of the reference's operator surface (ark_mpc_b200/fabric.py), written
the way the reference's own tests are (`execute_mock_mpc`, /root/reference/online-phase/src/lib.rs:116-201; cases from
algebra/scalar/authenticated_scalar.rs:1131-1715, algebra/curve/authenticated_curve.rs:882-1295,
integration/src/authenticated_scalar.rs:49-75, integration/src/circuits.rs:22-50): share inputs, run the gates on the
device, open with the MAC check, compare with the same expression on plaintext values."""
import random

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util_curve import CURVE_BY_ID, points_from_affine, xy_to_affine

pytestmark = pytest.mark.gpu

FIELDS = ["bn254_fr", "curve25519_fr"]
CURVE_ID = {"bn254_fr": 0, "curve25519_fr": 1}
N = 100


def fab():
    from ark_mpc_b200 import fabric

    return fabric


def rand_vals(field, seed, n=N):
    rng = random.Random(seed)
    p = po.FIELDS[field].p
    return [rng.randrange(p) for _ in range(n)]


def run_binary(field, op, expect, source=None):
    """Party 0 shares a, party 1 shares b; returns the opened result of op(a, b) checked on both parties."""
    F = fab()
    p = po.FIELDS[field].p
    a, b = rand_vals(field, 1), rand_vals(field, 2)

    def party(fabric):
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else N, 0)
        B = fabric.batch_share_scalar(b if fabric.party_id() == 1 else N, 1)
        res = op(F, fabric, A, B, a, b)
        return F.AuthenticatedScalarResult.open_authenticated_batch(res).result().to_ints()

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=source)
    want = [expect(x, y) % p for x, y in zip(a, b)]
    assert r0 == want and r1 == want


def random_source(pid, engine):
    return fab().DeviceTripleSource(pid, engine, seed=0xBEEF)


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("source", [None, random_source], ids=["party_id_source", "random_triples"])
def test_batch_mul(field, source):
    run_binary(field, lambda F, f, A, B, a, b: F.AuthenticatedScalarResult.batch_mul(A, B), lambda x, y: x * y, source)


@pytest.mark.parametrize("field", FIELDS)
def test_linear_gates(field):
    S = lambda F: F.AuthenticatedScalarResult
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_add(A, B), lambda x, y: x + y)
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_sub(A, B), lambda x, y: x - y)
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_neg(A), lambda x, y: -x)
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_mul_constant(A, 12345), lambda x, y: 12345 * x)
    # public operands: both parties know b
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_add_public(A, f.allocate_scalars(b)), lambda x, y: x + y)
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_sub_public(A, f.allocate_scalars(b)), lambda x, y: x - y)
    run_binary(field, lambda F, f, A, B, a, b: S(F).batch_mul_public(A, f.allocate_scalars(b)), lambda x, y: x * y, random_source)


@pytest.mark.parametrize("field", FIELDS)
def test_party_id_beaver_source_kat(field):
    """offline_prep.rs:137-158 / integration/src/lowgear.rs:37-45: the mock triple opens to 2 * 3 = 6 under key 1."""
    F = fab()

    def party(fabric):
        a, b, c = fabric.next_triple_batch(4)
        S = F.AuthenticatedScalarResult
        return [S.open_authenticated_batch(v).result().to_ints() for v in (a, b, c)]

    r0, r1 = F.execute_mock_mpc(party, field=field)
    assert r0 == r1 == [[2] * 4, [3] * 4, [6] * 4]


@pytest.mark.parametrize("field", FIELDS)
def test_share_and_open_and_empty(field):
    """integration/src/fabric.rs:15-32 (share-and-open of 0 and 1) and the empty batch (:854-856)."""
    F = fab()

    def party(fabric):
        v = fabric.batch_share_scalar([0, 1] if fabric.party_id() == 0 else 2, 0)
        S = F.AuthenticatedScalarResult
        opened = S.open_authenticated_batch(v).result().to_ints()
        e = fabric.allocate_scalar_shares(np.zeros((0, 8), dtype=np.uint64))
        empty = S.batch_mul(e, e)
        return opened, len(empty), len(S.open_batch(empty))

    r0, r1 = F.execute_mock_mpc(party, field=field)
    assert r0 == r1 == ([0, 1], 0, 0)


@pytest.mark.parametrize("field", FIELDS)
def test_xor_circuit(field):
    """authenticated_scalar.rs:1676-1688: a ^ b = a + b - 2ab on shared bits; a = b = 1 gives 0."""
    F = fab()

    def party(fabric):
        S = F.AuthenticatedScalarResult
        a = fabric.batch_share_scalar([1, 0, 1, 0] if fabric.party_id() == 0 else 4, 0)
        b = fabric.batch_share_scalar([1, 1, 0, 0] if fabric.party_id() == 1 else 4, 1)
        ab = S.batch_mul(a, b)
        res = S.batch_sub(S.batch_add(a, b), S.batch_mul_constant(ab, 2))
        return S.open_authenticated_batch(res).result().to_ints()

    r0, r1 = F.execute_mock_mpc(party, field=field)
    assert r0 == r1 == [0, 1, 1, 0]


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("what", ["mac", "share"])
def test_open_authenticated_detects_corruption(field, what):
    """integration/src/authenticated_scalar.rs:49-75: a modified MAC or share must make open_authenticated fail."""
    F = fab()

    def party(fabric):
        S = F.AuthenticatedScalarResult
        v = fabric.batch_share_scalar(rand_vals(field, 5, 16) if fabric.party_id() == 0 else 16, 0)
        if fabric.party_id() == 0:
            (v.modify_mac if what == "mac" else v.modify_share)(3)
        res = S.open_authenticated_batch(v)
        try:
            res.result()
            return "ok"
        except F.AuthenticationError:
            return "auth_error"

    assert F.execute_mock_mpc(party, field=field) == ("auth_error", "auth_error")


@pytest.mark.parametrize("field", FIELDS)
def test_inner_product(field):
    """integration/src/circuits.rs:22-50 (n = 100) and BASELINE config 4's shape: batch_mul + Sum + open_authenticated."""
    F = fab()
    p = po.FIELDS[field].p
    a, b = rand_vals(field, 7, 1000), rand_vals(field, 8, 1000)

    def party(fabric):
        S = F.AuthenticatedScalarResult
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else len(a), 0)
        B = fabric.batch_share_scalar(b if fabric.party_id() == 1 else len(b), 1)
        unfused = S.open_authenticated_batch(S.batch_mul(A, B).sum()).result().to_ints()
        fused = S.open_authenticated_batch(S.batch_mul_sum(A, B)).result().to_ints()   # recombine + sum in one kernel
        return unfused, fused

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=random_source)
    want = [sum(x * y for x, y in zip(a, b)) % p]
    assert r0 == r1 == (want, want)


# ---- points -------------------------------------------------------------------------------------------------------
def point_inputs(field, n, seed):
    cv = CURVE_ID[field]
    Cv = CURVE_BY_ID[cv]
    rng = random.Random(seed)
    scal = [rng.randrange(Cv.fr.p) for _ in range(n)]
    pts = [Cv.mul(Cv.generator, rng.randrange(1, Cv.fr.p)) for _ in range(n)]
    return cv, Cv, scal, pts, points_from_affine(cv, pts, rng)


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("source", [None, random_source], ids=["party_id_source", "random_triples"])
def test_point_batch_mul_and_msm(field, source):
    """authenticated_curve.rs:1217-1295: [x] * [P] opens (with MAC check) to x * P; msm opens to sum x_i P_i."""
    F = fab()
    n = 20
    cv, Cv, scal, pts, images = point_inputs(field, n, 11)

    def party(fabric):
        P = F.AuthenticatedPointResult
        X = fabric.batch_share_scalar(scal if fabric.party_id() == 0 else n, 0)
        Pt = fabric.batch_share_point(fabric.engine.upload_points(images) if fabric.party_id() == 1 else n, 1)
        prod = P.open_authenticated_batch(P.batch_mul(X, Pt)).result().to_affine_limbs()
        msm = P.open_authenticated_batch(P.msm(X, Pt)).result().to_affine_limbs()
        return prod, msm

    (prod0, msm0), (prod1, msm1) = F.execute_mock_mpc(party, field=field, beaver=source)
    want = [Cv.mul(Pi, xi) for Pi, xi in zip(pts, scal)]
    assert xy_to_affine(cv, prod0) == want and xy_to_affine(cv, prod1) == want
    total = Cv.identity
    for w in want:
        total = Cv.add(total, w)
    assert xy_to_affine(cv, msm0) == [total] and xy_to_affine(cv, msm1) == [total]


@pytest.mark.parametrize("field", FIELDS)
def test_point_linear_and_public_gates(field):
    """authenticated_curve.rs:882-1215: add, sub, neg, add_public, mul_public, mul_generator."""
    F = fab()
    n = 12
    cv, Cv, scal, pts, images = point_inputs(field, n, 13)
    _, _, scal2, pts2, images2 = point_inputs(field, n, 17)

    def party(fabric):
        P = F.AuthenticatedPointResult
        A = fabric.batch_share_point(fabric.engine.upload_points(images) if fabric.party_id() == 0 else n, 0)
        B = fabric.batch_share_point(fabric.engine.upload_points(images2) if fabric.party_id() == 1 else n, 1)
        pub = F.CurvePointResult(fabric, fabric.engine.upload_points(images2))
        s_pub = fabric.allocate_scalars(scal)
        X = fabric.batch_share_scalar(scal2 if fabric.party_id() == 0 else n, 0)
        outs = [P.batch_add(A, B), P.batch_sub(A, B), P.batch_neg(A), P.batch_add_public(A, pub), P.batch_sub_public(A, pub),
                P.batch_mul_public(s_pub, A), P.batch_mul_generator(X)]
        return [P.open_authenticated_batch(o).result().to_affine_limbs() for o in outs]

    r0, r1 = F.execute_mock_mpc(party, field=field)
    want = [[Cv.add(a, b) for a, b in zip(pts, pts2)], [Cv.sub(a, b) for a, b in zip(pts, pts2)], [Cv.neg(a) for a in pts],
            [Cv.add(a, b) for a, b in zip(pts, pts2)], [Cv.sub(a, b) for a, b in zip(pts, pts2)],
            [Cv.mul(a, s) for a, s in zip(pts, scal)], [Cv.mul(Cv.generator, s) for s in scal2]]
    for got0, got1, w in zip(r0, r1, want):
        assert xy_to_affine(cv, got0) == w and xy_to_affine(cv, got1) == w


@pytest.mark.parametrize("field", FIELDS)
def test_point_open_authenticated_detects_corruption(field):
    F = fab()
    n = 6
    cv, Cv, scal, pts, images = point_inputs(field, n, 19)

    def party(fabric):
        P = F.AuthenticatedPointResult
        A = fabric.batch_share_point(fabric.engine.upload_points(images) if fabric.party_id() == 0 else n, 0)
        if fabric.party_id() == 1:
            A.modify_mac(2)
        try:
            P.open_authenticated_batch(A).result()
            return "ok"
        except F.AuthenticationError:
            return "auth_error"

    assert F.execute_mock_mpc(party, field=field) == ("auth_error", "auth_error")


@pytest.mark.parametrize("field", FIELDS)
@pytest.mark.parametrize("source", [None, random_source], ids=["party_id_source", "random_triples"])
def test_constant_wires_bits_and_inverse_pairs(field, source):
    """fabric.rs:497-546 (zero / one wires, party 0 holds 0 and party 1 holds 1 of the shared one), :943-978 (inverse pairs, shared bits)."""
    F = fab()
    p = po.FIELDS[field].p
    n = 17
    a = rand_vals(field, 5, n)

    def party(fabric):
        S = F.AuthenticatedScalarResult
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else n, 0)
        left, right = fabric.random_inverse_pairs(n)
        bits = fabric.random_shared_bits(n)
        outs = [fabric.zeros_authenticated(n), fabric.ones_authenticated(n), S.batch_add(A, fabric.ones_authenticated(n)),
                S.batch_mul(left, right), bits, S.batch_mul(bits, bits)]
        opened = [S.open_authenticated_batch(o).result().to_ints() for o in outs]
        return opened, fabric.zeros(n).to_ints(), fabric.ones(n).to_ints()

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=source)
    assert r0 == r1
    opened, zeros, ones = r0
    assert zeros == [0] * n and ones == [1] * n
    assert opened[0] == [0] * n and opened[1] == [1] * n
    assert opened[2] == [(x + 1) % p for x in a]
    assert opened[3] == [1] * n
    assert all(b in (0, 1) for b in opened[4]) and opened[5] == opened[4]
    if source is not None:
        assert 0 < sum(opened[4]) < n   # random bits, not a constant


@pytest.mark.parametrize("field", FIELDS)
def test_public_point_gates_identities_and_share_corruption(field):
    """curve.rs batch_sub / batch_neg on public points, the identity wires (fabric.rs:536-546), allocate_points, and a corrupted point
    share failing the MAC check."""
    F = fab()
    n = 9
    cv, Cv, scal, pts, images = point_inputs(field, n, 23)
    _, _, _, pts2, images2 = point_inputs(field, n, 29)

    def party(fabric):
        P, C = F.AuthenticatedPointResult, F.CurvePointResult
        A = fabric.batch_share_point(fabric.engine.upload_points(images) if fabric.party_id() == 0 else n, 0)
        pa, pb = fabric.allocate_points(fabric.engine.upload_points(images)), fabric.allocate_points(fabric.engine.upload_points(images2))
        pub = [C.batch_sub(pa, pb).to_affine_limbs(), C.batch_neg(pa).to_affine_limbs(),
               C.batch_add(pa, fabric.curve_identities(n)).to_affine_limbs()]
        same = P.open_authenticated_batch(P.batch_add(A, fabric.curve_identities_authenticated(n))).result().to_affine_limbs()
        if fabric.party_id() == 0:
            A.modify_share(1)
        try:
            P.open_authenticated_batch(A).result()
            verdict = "ok"
        except F.AuthenticationError:
            verdict = "auth_error"
        return pub, same, verdict

    r0, r1 = F.execute_mock_mpc(party, field=field)
    want = [[Cv.sub(a, b) for a, b in zip(pts, pts2)], [Cv.neg(a) for a in pts], list(pts)]
    for r in (r0, r1):
        for got, w in zip(r[0], want):
            assert xy_to_affine(cv, got) == w
        assert xy_to_affine(cv, r[1]) == list(pts)
        assert r[2] == "auth_error"


def test_public_fft_round_trip_and_padding():
    """scalar_result.rs fft / ifft on public values: ifft(fft(x)) = x zero-padded to the domain size; agrees with the transform of shares."""
    F = fab()
    n = 24
    a = rand_vals("bn254_fr", 31, n)

    def party(fabric):
        S, P = F.AuthenticatedScalarResult, F.ScalarResult
        pub = fabric.allocate_scalars(a)
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else n, 0)
        fa = P.fft(pub)
        return P.ifft(fa).to_ints(), fa.to_ints(), S.open_authenticated_batch(S.fft(A)).result().to_ints()

    r0, r1 = F.execute_mock_mpc(party, field="bn254_fr")
    assert r0 == r1
    back, fa, fa_shared = r0
    assert back == a + [0] * (32 - n)
    assert fa == fa_shared and len(fa) == 32
