"""GPU parity tests of the point gates (BN254 G1, Curve25519 Edwards): every call goes through the C ABI and is
compared with the CPU oracle in canonical affine form (projective representatives are not unique,
/root/reference/online-phase/src/algebra/curve/curve.rs:46).  Cases restate the reference's tests
(algebra/curve/authenticated_curve.rs:882-1295: add / sub / neg / mul_public / mul / mul_generator / open, N = 100)."""
import random

import numpy as np
import pytest
import torch

from oracle import coracle as co
from oracle import pyoracle as po
from tests.util import split_aos
from tests.util_curve import CURVE_BY_ID, CURVE_NAME, TwoPartyPointData, points_from_affine

pytestmark = pytest.mark.gpu

CURVES = [0, 1]
FIELD_OF = {0: "bn254_fr", 1: "curve25519_fr"}


@pytest.fixture(scope="module")
def engines():
    from ark_mpc_b200.engine import Engine

    out = {}
    for cv in CURVES:
        E = Engine(0, FIELD_OF[cv])
        E.bind_curve(CURVE_NAME[cv])
        out[cv] = E
    return out


def norm_gpu(E, t):
    return E.download(E.pt_normalize(t))


def norm_cpu(cv, a):
    w = co.point_words(cv)
    return co.pt_normalize(cv, np.ascontiguousarray(a).reshape(-1, w))


def rand_points(cv, seed, n, with_identity=True):
    """n points s_i*G as a randomised projective image (plus the identity and the generator up front)."""
    fr = co.CURVE_FR[cv]
    s = co.synth(fr, seed, 0, n)
    pts = co.pt_mul_generator(cv, s)
    if with_identity and n >= 3:
        Cv = CURVE_BY_ID[cv]
        pts[:2] = points_from_affine(cv, [Cv.identity, Cv.generator], random.Random(seed))
    return pts


@pytest.mark.parametrize("cv", CURVES)
@pytest.mark.parametrize("n", [1, 100, 777])
def test_point_linear_gates(engines, cv, n):
    E = engines[cv]
    a, b = rand_points(cv, 1, n), rand_points(cv, 2, n)
    if n >= 8:
        b[4] = a[4]                                         # P + P -> doubling path
        nb = co.pt_add(cv, points_from_affine(cv, [CURVE_BY_ID[cv].identity]), a[6:7], sub=True)[0]
        b[6] = nb                                           # P + (-P) -> identity
    A, B = E.upload_points(a), E.upload_points(b)
    assert np.array_equal(norm_gpu(E, E.pt_add(A, B)), norm_cpu(cv, co.pt_add(cv, a, b)))
    assert np.array_equal(norm_gpu(E, E.pt_sub(A, B)), norm_cpu(cv, co.pt_add(cv, a, b, sub=True)))
    ident = points_from_affine(cv, [CURVE_BY_ID[cv].identity] * n)
    assert np.array_equal(norm_gpu(E, E.pt_neg(A)), norm_cpu(cv, co.pt_add(cv, ident, a, sub=True)))
    # normalize itself against the oracle on non-trivial representatives
    assert np.array_equal(norm_gpu(E, A), norm_cpu(cv, a))


@pytest.mark.parametrize("cv", CURVES)
@pytest.mark.parametrize("n", [1, 100, 1500])
def test_scalar_multiplications(engines, cv, n):
    E = engines[cv]
    fr = co.CURVE_FR[cv]
    Cv = CURVE_BY_ID[cv]
    s = co.synth(fr, 3, 0, n)
    k = min(n, 4)
    s[:k] = co.to_mont(fr, co.ints_to_limbs([0, 1, Cv.fr.p - 1, 16]))[:k]
    pts = rand_points(cv, 4, n)
    S, P = E.upload(s), E.upload_points(pts)
    assert np.array_equal(norm_gpu(E, E.pt_mul(S, P)), norm_cpu(cv, co.pt_mul(cv, s, pts)))
    assert np.array_equal(norm_gpu(E, E.pt_mul_generator_public(S)), norm_cpu(cv, co.pt_mul_generator(cv, s)))
    # batch_mul_generator / batch_mul_authenticated on ScalarShares: (share*G, mac*G), (share*P, mac*P)
    m = co.synth(fr, 5, 0, n)
    M = E.upload(m)
    w = co.point_words(cv)
    got = E.download(E.pt_mul_generator((S, M)))
    assert np.array_equal(norm_cpu(cv, got[:, :w]), norm_cpu(cv, co.pt_mul_generator(cv, s)))
    assert np.array_equal(norm_cpu(cv, got[:, w:]), norm_cpu(cv, co.pt_mul_generator(cv, m)))
    got = E.download(E.pt_mul_authenticated((S, M), P))
    assert np.array_equal(norm_cpu(cv, got[:, :w]), norm_cpu(cv, co.pt_mul(cv, s, pts)))
    assert np.array_equal(norm_cpu(cv, got[:, w:]), norm_cpu(cv, co.pt_mul(cv, m, pts)))
    # batch_mul_public on PointShares: (s*share, s*mac)
    macs = rand_points(cv, 6, n, with_identity=False)
    ps = np.ascontiguousarray(np.concatenate([pts, macs], axis=1))
    got = E.download(E.pt_share_mul_public(S, E.upload_points(ps)))
    assert np.array_equal(norm_cpu(cv, got[:, :w]), norm_cpu(cv, co.pt_mul(cv, s, pts)))
    assert np.array_equal(norm_cpu(cv, got[:, w:]), norm_cpu(cv, co.pt_mul(cv, s, macs)))


@pytest.mark.parametrize("cv", CURVES)
@pytest.mark.parametrize("party", [0, 1])
def test_share_add_public_and_mac_check(engines, cv, party):
    E = engines[cv]
    n = 300
    fr = co.CURVE_FR[cv]
    w = co.point_words(cv)
    key = co.synth(fr, 9, 0, 1)[0]
    ps = np.ascontiguousarray(np.concatenate([rand_points(cv, 10, n), rand_points(cv, 11, n, False)], axis=1))
    pub = rand_points(cv, 12, n)
    PS, PUB = E.upload_points(ps), E.upload_points(pub)
    for sub in (False, True):
        got = E.download(E.pt_share_add_public(party, key, PS, PUB, sub=sub))
        want = co.pt_share_add_public(cv, party, key, ps, pub, sub=sub)
        assert np.array_equal(norm_cpu(cv, got), norm_cpu(cv, want))
    # mac_key * opened - mac  (authenticated_curve.rs:217-232)
    got = E.download(E.pt_mac_check(key, PUB, PS))
    kp = co.pt_mul(cv, np.tile(key, (n, 1)), pub)
    want = co.pt_add(cv, kp, np.ascontiguousarray(ps[:, w:]), sub=True)
    assert np.array_equal(norm_cpu(cv, got), norm_cpu(cv, want))
    # zero-sum verification (:128-131)
    neg = E.pt_neg(PUB)
    assert E.pt_sum_is_identity(PUB, neg)
    bad = E.download(neg).copy()
    bad[n // 2] = pub[1]   # PUB[n/2] + G is not the identity
    assert not E.pt_sum_is_identity(PUB, E.upload_points(bad))


@pytest.mark.parametrize("cv", CURVES)
@pytest.mark.parametrize("n", [1, 100, 2049])
def test_point_beaver_mul_matches_oracle(engines, cv, n):
    """AuthenticatedPointResult::batch_mul (authenticated_curve.rs:682-714) for both parties: the fused two-pass kernel against
    the unfused 10-scalar-multiplication restatement, per-party outputs compared in affine form, opened d and E included."""
    E = engines[cv]
    D = TwoPartyPointData(cv, n, seed=4242 + cv)
    out0, out1, d_open, E_open = D.oracle_point_mul(threads=8)
    planes = lambda t: tuple(E.upload(v) for v in split_aos(t))
    masks = []
    for p in (0, 1):
        xs = E.upload(split_aos(D.x[p])[0])
        a_s = E.upload(split_aos(D.a[p])[0])
        b_s = E.upload(split_aos(D.b[p])[0])
        masks.append(E.pt_beaver_mask(xs, E.upload_points(D.P[p]), a_s, b_s))
    w = co.point_words(cv)
    for p, want in ((0, out0), (1, out1)):
        out, (d_o, E_o) = E.pt_beaver_recombine(p, D.keys[p], masks[p][0], masks[1 - p][0], masks[p][1], masks[1 - p][1],
                                                planes(D.a[p]), planes(D.b[p]), planes(D.c[p]), want_open=True)
        got = E.download(out)
        assert np.array_equal(norm_cpu(cv, got), norm_cpu(cv, want)), f"party {p}: result PointShares differ from the oracle"
        assert np.array_equal(E.download(d_o), d_open)
        assert np.array_equal(norm_gpu(E, E_o), norm_cpu(cv, E_open))
    # without the optional opened outputs
    out, _ = E.pt_beaver_recombine(0, D.keys[0], masks[0][0], masks[1][0], masks[0][1], masks[1][1], planes(D.a[0]), planes(D.b[0]), planes(D.c[0]))
    assert np.array_equal(norm_cpu(cv, E.download(out)), norm_cpu(cv, out0))


@pytest.mark.parametrize("cv", CURVES)
def test_point_beaver_mul_opens_to_product_large(engines, cv):
    """Size-independent property at a size the unfused oracle cannot reach quickly: the two parties' outputs open to x*P
    (checked against the device's own single scalar multiplication) and the MAC shares open to key * (x*P)."""
    E = engines[cv]
    n = 1 << 14
    fr = co.CURVE_FR[cv]
    rnd = lambda s: E.random(s, 0, n)
    key0, key1 = co.synth(fr, 900, 0, 1)[0], co.synth(fr, 901, 0, 1)[0]
    key = co.scalar_add(fr, key0.reshape(1, 4), key1.reshape(1, 4))[0]

    def shared(seed, val=None):
        v = rnd(seed) if val is None else val
        s0, m0 = rnd(seed + 1), rnd(seed + 2)
        return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

    xv, x0, x1 = shared(10)
    sv, s0, s1 = shared(20)
    av, a0, a1 = shared(30)
    bv, b0, b1 = shared(40)
    _, c0, c1 = shared(50, E.mul(av, bv))
    P = [E.pt_mul_generator(s0), E.pt_mul_generator(s1)]
    parts = [dict(key=key0, x=x0, a=a0, b=b0, c=c0), dict(key=key1, x=x1, a=a1, b=b1, c=c1)]
    masks = [E.pt_beaver_mask(parts[p]["x"][0], P[p], parts[p]["a"][0], parts[p]["b"][0]) for p in (0, 1)]
    outs = [E.pt_beaver_recombine(p, parts[p]["key"], masks[p][0], masks[1 - p][0], masks[p][1], masks[1 - p][1], parts[p]["a"], parts[p]["b"],
                                  parts[p]["c"])[0] for p in (0, 1)]
    opened = E.pt_add(outs[0], outs[1])                        # (n, 2w): share and mac halves opened at once
    w = co.point_words(cv)
    xP = E.pt_mul_generator_public(E.mul(xv, sv))              # x * (s*G) = (x*s) * G
    got = E.download(E.pt_normalize(opened)).reshape(n, 2, 8)
    want_share = E.download(E.pt_normalize(xP))
    want_mac = E.download(E.pt_normalize(E.pt_mul_generator_public(E.scale(E.mul(xv, sv), key))))
    assert np.array_equal(got[:, 0, :], want_share)
    assert np.array_equal(got[:, 1, :], want_mac)
    # spot-check the device's reference values against the CPU oracle
    idx = [0, 1, n // 2, n - 1]
    xs = E.download(E.mul(xv, sv))[idx]
    assert np.array_equal(want_share[idx], co.pt_normalize(cv, co.pt_mul_generator(cv, np.ascontiguousarray(xs))))
    # and BOTH parties' PointShare outputs against the unfused oracle on a strided sample across the whole batch (every 256th gate)
    from tests.util import aos

    sel = torch.arange(0, n, 256, device=outs[0].device)
    m = sel.numel()
    cut = lambda pl: aos(E.download(pl[0][sel].contiguous()), E.download(pl[1][sel].contiguous()))
    ins = ((key0, key1), (cut(x0), cut(x1)), tuple(E.download(P[p][sel].contiguous()) for p in (0, 1)), (cut(a0), cut(a1)), (cut(b0), cut(b1)),
           (cut(c0), cut(c1)))
    o0, o1, _, _ = co.two_party_point_mul(cv, 4, *ins, want_open=False)
    for p, o in ((0, o0), (1, o1)):
        got_p = E.download(E.pt_normalize(outs[p][sel].contiguous()))
        assert np.array_equal(got_p, co.pt_normalize(cv, o.reshape(2 * m, -1))), f"party {p}: strided sample differs from the unfused oracle"


@pytest.mark.parametrize("cv", CURVES)
def test_point_empty_and_errors(engines, cv):
    import ark_mpc_b200._native as nat

    E = engines[cv]
    e = E.empty_points(0)
    assert E.pt_add(e, e).shape[0] == 0
    with pytest.raises(nat.ArkMpcError):
        E._call("arkmpc_pt_add", 7, 1, E._p(E.empty_points(1)), E._p(E.empty_points(1)), E._p(E.empty_points(1)))


@pytest.mark.parametrize("cv", CURVES)
@pytest.mark.parametrize("n", [0, 1, 33, 1000, 40000, 70001])
def test_point_sums_and_msm(engines, cv, n):
    """Sum of points / PointShares (authenticated_curve.rs:798-803) and the public MSM (curve.rs:549-560) against the oracle."""
    E = engines[cv]
    w = co.point_words(cv)
    Cv = CURVE_BY_ID[cv]
    ident = points_from_affine(cv, [Cv.identity])
    if n == 0:
        assert np.array_equal(norm_gpu(E, E.pt_sum(E.empty_points(0))), norm_cpu(cv, ident))
        return
    fr = co.CURVE_FR[cv]
    pts, macs = rand_points(cv, 41, n), rand_points(cv, 42, n, with_identity=False)
    acc_s, acc_m = ident.copy(), ident.copy()
    for i in range(0, n, 1):  # the oracle folds serially, as the reference's gate does
        acc_s = co.pt_add(cv, acc_s, pts[i:i + 1])
        acc_m = co.pt_add(cv, acc_m, macs[i:i + 1])
        if n > 2000 and i >= 1999:
            break
    m = min(n, 2000)
    P = E.upload_points(pts[:m])
    assert np.array_equal(norm_gpu(E, E.pt_sum(P)), norm_cpu(cv, acc_s))
    ps = np.ascontiguousarray(np.concatenate([pts[:m], macs[:m]], axis=1))
    got = E.download(E.pt_share_sum(E.upload_points(ps)))
    assert np.array_equal(norm_cpu(cv, got[:, :w]), norm_cpu(cv, acc_s)) and np.array_equal(norm_cpu(cv, got[:, w:]), norm_cpu(cv, acc_m))
    # MSM: sum_i s_i * P_i == (sum_i s_i * t_i) * G for P_i = t_i * G  (size-independent identity, checked at the full n)
    t = co.synth(fr, 43, 0, n)
    s = co.synth(fr, 44, 0, n)
    T, S = E.upload(t), E.upload(s)
    Pn = E.pt_mul_generator_public(T)
    lhs = E.pt_msm(S, Pn)
    rhs = E.pt_mul_generator_public(E.sum(E.mul(S, T)))
    assert np.array_equal(norm_gpu(E, lhs), norm_gpu(E, rhs))
    # msm_authenticated (curve.rs:619-642): two MSMs over the same points
    m2 = co.synth(fr, 45, 0, n)
    got2 = E.download(E.pt_msm_authenticated((S, E.upload(m2)), Pn))
    assert np.array_equal(norm_cpu(cv, got2[:, :w]), norm_gpu(E, lhs))
    assert np.array_equal(norm_cpu(cv, got2[:, w:]), norm_gpu(E, E.pt_mul_generator_public(E.sum(E.mul(E.upload(m2), T)))))
    if n <= 1000:
        want = ident.copy()
        prods = co.pt_mul(cv, s, co.pt_mul_generator(cv, t))
        for i in range(n):
            want = co.pt_add(cv, want, prods[i:i + 1])
        assert np.array_equal(norm_gpu(E, lhs), norm_cpu(cv, want))


def test_point_and_ntt_argument_errors(engines):
    """Error behaviour at the boundary: bad curve / party ids, misaligned or missing arrays and aliased FFT planes are reported as
    ARKMPC_ERR_INVALID with a message; nothing is computed."""
    import ctypes as C

    import ark_mpc_b200._native as nat

    E = engines[0]
    lib, ctx = E.lib, E.ctx
    pts = E.pt_mul_generator_public(E.random(1, 0, 4))
    s = E.random(2, 0, 4)
    key = E.key_limbs(co.synth(0, 3, 0, 1)[0])
    kp = key.ctypes.data_as(C.c_void_p)
    p = lambda t: C.c_void_p(t.data_ptr())
    assert lib.arkmpc_pt_mul(ctx, 9, 4, p(s), p(pts), p(pts)) == nat.ERR_INVALID                      # unknown curve
    assert lib.arkmpc_pt_mul(ctx, 0, 4, None, p(pts), p(pts)) == nat.ERR_INVALID                      # null array
    assert lib.arkmpc_pt_mul(ctx, 0, 4, p(s), C.c_void_p(pts.data_ptr() + 8), p(pts)) == nat.ERR_INVALID  # misaligned
    assert b"aligned" in lib.arkmpc_last_error(ctx)
    ps = E.pt_mul_generator((s, s))
    assert lib.arkmpc_pt_share_add_public(ctx, 0, 2, kp, 4, p(ps), p(pts), p(ps)) == nat.ERR_INVALID  # party id
    assert lib.arkmpc_pt_beaver_recombine(ctx, 0, 0, kp, 4, p(s), p(s), p(pts), p(pts), p(s), p(s), p(s), p(s), p(s), p(s), p(ps), p(s), None) == nat.ERR_INVALID
    assert lib.arkmpc_fr_fft(ctx, 0, 2, 0, p(s), p(s)) == nat.ERR_INVALID                             # in place is not supported
    assert lib.arkmpc_fr_fft(ctx, 0, 40, 0, p(s), p(pts)) == nat.ERR_INVALID                          # domain too large
    assert lib.arkmpc_pt_mul(ctx, 0, 0, None, None, None) == nat.OK                                   # empty batch is a no-op


@pytest.mark.parametrize("cv", [0, 1])
def test_point_validation_rejects_off_curve_and_torsion(cv):
    """arkmpc_pt_validate: on-curve and prime-order-subgroup membership of points received from the peer (curve.rs:105-135), the
    precondition of the regrouped recombination.  Curve25519 (cofactor 8): a point with a torsion component is on the curve and
    must be rejected."""
    from ark_mpc_b200.engine import Engine

    E = Engine(0, ["bn254_fr", "curve25519_fr"][cv])
    n = 300
    pts = E.pt_mul_generator_public(E.random(5, 0, n))
    assert E.pt_validate(pts)
    assert E.pt_validate(E.pt_mul_generator_public(E.upload(np.zeros((4, 4), dtype=np.uint64))))  # identities
    bad = pts.clone()
    bad[n - 3, 0] += 1  # X coordinate off by one (Montgomery image): off the curve
    assert not E.pt_validate(bad)
    if cv == 1:
        F = po.FIELDS["curve25519_fq"]
        q = F.p
        M = lambda v: [(F.to_mont(v) >> (64 * i)) & (2**64 - 1) for i in range(4)]
        # the order-2 point (0, -1) in extended coordinates (X, Y, T, Z) = (0, -1, 0, 1)
        t2 = np.array([M(0) + M(q - 1) + M(0) + M(1)], dtype=np.uint64)
        T2 = E.upload_points(t2)
        assert not E.pt_validate(T2)
        mixed = pts.clone()
        mixed[7:8] = E.pt_add(pts[7:8].contiguous(), T2)  # on the curve, outside the prime-order subgroup
        assert not E.pt_validate(mixed)
        assert E.pt_validate(E.pt_add(mixed[7:8].contiguous(), T2))  # adding the order-2 point twice removes it
    E.close()
