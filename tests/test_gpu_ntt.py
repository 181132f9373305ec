"""GPU parity tests of batch inversion and the FFT on shares through the C ABI, against the CPU oracle,
plus the two-party protocols built on them (batch_inverse, poly-mul shape fft -> batch_mul -> ifft) through the fabric mirror."""
import random

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from ark_mpc_b200.engine import Engine

    return {0: Engine(0, "bn254_fr"), 1: Engine(0, "curve25519_fr")}


@pytest.mark.parametrize("fid", [0, 1])
@pytest.mark.parametrize("n", [1, 7, 8, 9, 15, 16, 17, 1000, 100003, 131072, 131073, (1 << 20) + 5])
def test_batch_inverse(engines, fid, n):
    """Sizes on both sides of every level boundary of the product tree (groups of 8; <= 131072 elements go to the top kernel
    directly, 131073 needs one sweep, 2^20 + 5 one sweep with a ragged last group), zeros in every position class."""
    E = engines[fid]
    a = co.synth(fid, 21, 0, n)
    for k in (0, n // 2, n - 1):       # zeros stay zero (ark_ff::batch_inversion)
        if n > 2:
            a[k] = 0
    if n >= 1000:
        g = (n + 7) // 8
        a[3::g] = 0                    # a whole group of zeros (elements 3, 3 + g, ... form group 3 of the first level)
        a[5] = 0
        a[5 + g] = 0
    got = E.download(E.batch_inverse(E.upload(a)))
    assert np.array_equal(got, co.batch_inverse(fid, a))
    # x * x^-1 = 1 wherever x != 0
    one = co.to_mont(fid, co.ints_to_limbs([1]))[0]
    prod = co.scalar_mul(fid, a, got)
    nz = a.any(axis=1)
    assert np.array_equal(prod[nz], np.tile(one, (int(nz.sum()), 1)))


@pytest.mark.parametrize("log2n", [0, 1, 2, 5, 9, 10, 11, 13, 16])
def test_fft_matches_oracle(engines, log2n):
    E = engines[0]
    n = 1 << log2n
    a = co.synth(0, 31 + log2n, 0, n)
    A = E.upload(a)
    fwd = E.fft(A)
    assert np.array_equal(E.download(fwd), co.fft(0, a))
    inv = E.fft(A, inverse=True)
    assert np.array_equal(E.download(inv), co.fft(0, a, inverse=True))
    assert np.array_equal(E.download(E.fft(fwd, inverse=True)), a)
    # share planes, and a switch of domain size on the same context (twiddle cache)
    m = co.synth(0, 77, 0, n)
    s_out, m_out = E.share_fft((A, E.upload(m)))
    assert np.array_equal(E.download(s_out), co.fft(0, a)) and np.array_equal(E.download(m_out), co.fft(0, m))


def test_fft_large_properties(engines):
    """2^20: round trip and linearity (size-independent properties), spot-checked against the definition."""
    E = engines[0]
    n = 1 << 20
    a, b = E.random(5, 0, n), E.random(6, 0, n)
    fa, fb = E.fft(a), E.fft(b)
    assert np.array_equal(E.download(E.fft(fa, inverse=True)), E.download(a))
    assert np.array_equal(E.download(E.fft(E.add(a, b))), E.download(E.add(fa, fb)))
    # X_0 = sum x_i ; X_{n/2} = sum (-1)^i x_i
    x = E.download(a)
    F = po.BN254_FR
    assert np.array_equal(E.download(fa)[0], E.download(E.sum(a))[0])
    got = F.from_mont(co.limbs_to_ints(E.download(fa)[n // 2:n // 2 + 1])[0])
    ev, od = E.download(E.sum(a[0::2].contiguous()))[0], E.download(E.sum(a[1::2].contiguous()))[0]
    want = (F.from_mont(co.limbs_to_ints(ev)[0]) - F.from_mont(co.limbs_to_ints(od)[0])) % F.p
    assert got == want


def test_fft_unsupported_field_is_reported(engines):
    import ark_mpc_b200._native as nat

    E = engines[1]
    with pytest.raises(nat.ArkMpcError) as e:
        E.fft(E.random(1, 0, 8))
    assert e.value.status == nat.ERR_UNSUPPORTED


@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
@pytest.mark.parametrize("source", ["party_id", "random"])
def test_two_party_batch_inverse(field, source):
    """authenticated_scalar.rs:1596-1620 (test_batch_inverse shape): [x]^-1 opens to x^-1."""
    from ark_mpc_b200 import fabric as F

    p = po.FIELDS[field].p
    rng = random.Random(9)
    xs = [rng.randrange(1, p) for _ in range(50)]
    src = None if source == "party_id" else (lambda pid, eng: F.DeviceTripleSource(pid, eng, seed=0xFEED))

    def party(fabric):
        S = F.AuthenticatedScalarResult
        X = fabric.batch_share_scalar(xs if fabric.party_id() == 0 else len(xs), 0)
        return S.open_authenticated_batch(S.batch_inverse(X)).result().to_ints()

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=src)
    assert r0 == r1 == [pow(x, -1, p) for x in xs]


def test_two_party_polynomial_product_via_fft():
    """The poly-mul caller of the path (algebra/poly/authenticated_poly.rs:377-401): FFT -> batch_mul -> IFFT on shares."""
    from ark_mpc_b200 import fabric as F

    Fd = po.BN254_FR
    rng = random.Random(12)
    deg = 37
    pa = [rng.randrange(Fd.p) for _ in range(deg)]
    pb = [rng.randrange(Fd.p) for _ in range(deg)]
    size = 128                                   # >= 2*deg - 1
    pad = lambda c: c + [0] * (size - len(c))

    def party(fabric):
        S = F.AuthenticatedScalarResult
        A = fabric.batch_share_scalar(pad(pa) if fabric.party_id() == 0 else size, 0)
        B = fabric.batch_share_scalar(pad(pb) if fabric.party_id() == 1 else size, 1)
        prod = S.ifft(S.batch_mul(S.fft(A), S.fft(B)))
        return S.open_authenticated_batch(prod).result().to_ints()

    r0, r1 = F.execute_mock_mpc(party, field="bn254_fr", beaver=lambda pid, eng: F.DeviceTripleSource(pid, eng, seed=0xF00D))
    want = [0] * size
    for i, x in enumerate(pa):
        for j, y in enumerate(pb):
            want[i + j] = (want[i + j] + x * y) % Fd.p
    assert r0 == r1 == want


@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
def test_two_party_div_and_add_constant(field):
    """authenticated_scalar.rs tests `test_div` (a / b = a * b^-1, :1596-1620 region) and add/sub with a constant."""
    from ark_mpc_b200 import fabric as F

    p = po.FIELDS[field].p
    rng = random.Random(15)
    a = [rng.randrange(p) for _ in range(40)]
    b = [rng.randrange(1, p) for _ in range(40)]

    def party(fabric):
        S = F.AuthenticatedScalarResult
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else len(a), 0)
        B = fabric.batch_share_scalar(b if fabric.party_id() == 1 else len(b), 1)
        quot = S.batch_mul(A, S.batch_inverse(B))
        plus7 = S.batch_add_public(A, fabric.allocate_scalars([7] * len(a)))
        return (S.open_authenticated_batch(quot).result().to_ints(), S.open_authenticated_batch(plus7).result().to_ints())

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=lambda pid, eng: F.DeviceTripleSource(pid, eng, seed=0xD1F))
    want = ([x * pow(y, -1, p) % p for x, y in zip(a, b)], [(x + 7) % p for x in a])
    assert r0 == want and r1 == want


@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
def test_two_party_named_div_pow_constant_ops(field):
    """batch_div (:974-977), division by a public value (:953-958), pow (:86-101, test_pow :1622-1640 region) and
    batch_add_constant (:531-560) through their own entry points; public-side constant ops and pow (scalar_result.rs:26-39,119,205,281)."""
    from ark_mpc_b200 import fabric as F

    p = po.FIELDS[field].p
    rng = random.Random(21)
    n = 33
    a = [rng.randrange(p) for _ in range(n)]
    b = [rng.randrange(1, p) for _ in range(n)]
    c = [rng.randrange(1, p) for _ in range(n)]

    def party(fabric):
        S, P = F.AuthenticatedScalarResult, F.ScalarResult
        A = fabric.batch_share_scalar(a if fabric.party_id() == 0 else n, 0)
        B = fabric.batch_share_scalar(b if fabric.party_id() == 1 else n, 1)
        C = fabric.allocate_scalars(c)
        outs = [S.batch_div(A, B), S.batch_div_public(A, C), S.batch_pow(A, 13), S.batch_pow(A, 1), S.batch_pow(A, 0),
                S.batch_add_constant(A, c)]
        opened = [S.open_authenticated_batch(o).result().to_ints() for o in outs]
        pub = [P.batch_add_constant(C, a).to_ints(), P.batch_sub_constant(C, a).to_ints(), P.batch_mul_constant(C, a).to_ints(),
               P.batch_pow(C, 11).to_ints(), P.batch_pow(C, 0).to_ints()]
        return opened, pub

    r0, r1 = F.execute_mock_mpc(party, field=field, beaver=lambda pid, eng: F.DeviceTripleSource(pid, eng, seed=0xBEE))
    want = ([x * pow(y, -1, p) % p for x, y in zip(a, b)], [x * pow(y, -1, p) % p for x, y in zip(a, c)], [pow(x, 13, p) for x in a], a,
            [0] * n, [(x + y) % p for x, y in zip(a, c)])
    want_pub = ([(y + x) % p for x, y in zip(a, c)], [(y - x) % p for x, y in zip(a, c)], [y * x % p for x, y in zip(a, c)],
                [pow(y, 11, p) for y in c], [1] * n)
    for r in (r0, r1):
        assert tuple(r[0]) == want
        assert tuple(r[1]) == want_pub
