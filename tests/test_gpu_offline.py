"""Offline-phase batch algebra on the device: ValueMacBatch arithmetic, the Beaver multiplication and the
sacrifice check of LowGear (/root/reference/offline-phase/src/structs.rs:321-382, lowgear/multiplication.rs:13-39,
lowgear/triplets.rs:118-150, lowgear/mac_check.rs:14-48) as two-party runs over the mock network, compared with the oracle's
share algebra per party and with the plaintext identities."""
import numpy as np
import pytest

from oracle import coracle as co
from tests.util import aos

pytestmark = pytest.mark.gpu


def _run(fn, field="bn254_fr"):
    from ark_mpc_b200 import fabric as fb

    return fb.execute_mock_mpc(fn, field=field, beaver=lambda pid, E: fb.DeviceTripleSource(pid, E, seed=0xBEEF))


@pytest.mark.parametrize("field,fid", [("bn254_fr", 0), ("curve25519_fr", 1)])
def test_value_mac_batch_arithmetic_matches_the_oracle(field, fid):
    from ark_mpc_b200 import offline as off

    n = 1000

    def party(f):
        E, src = f.engine, f.offline_phase
        a, b, _ = src.next_triplet_batch(n)
        A, B = off.ValueMacBatch(f, *a), off.ValueMacBatch(f, *b)
        s = E.download(E.random(77, 0, 1))[0].copy()
        v = E.random(78, 0, n)
        dl = lambda x: aos(E.download(x.share), E.download(x.mac))
        lo, hi = (A + B).split_at(300)
        return dict(a=dl(A), b=dl(B), s=s, v=E.download(v), add=dl(A + B), sub=dl(A - B), muls=dl(A.mul_scalar(s)), mulv=dl(A.mul_elementwise(v)),
                    pub=dl(A.add_public_value(v)), key=f.mac_key().copy(), lens=(len(lo), len(hi)), vals=E.download(A.values()), macs=E.download(A.macs()))

    for pid, r in enumerate(_run(party, field)):
        assert np.array_equal(r["add"], co.batch_add(fid, r["a"], r["b"]))
        assert np.array_equal(r["sub"], co.batch_sub(fid, r["a"], r["b"]))
        assert np.array_equal(r["muls"], co.batch_mul_public(fid, r["a"], np.tile(r["s"], (n, 1))))
        assert np.array_equal(r["mulv"], co.batch_mul_public(fid, r["a"], r["v"]))
        assert np.array_equal(r["pub"], co.batch_add_public(fid, pid, r["key"], r["a"], r["v"]))
        assert r["lens"] == (300, 700)
        assert np.array_equal(r["vals"], r["a"][:, :4]) and np.array_equal(r["macs"], r["a"][:, 4:])


@pytest.mark.parametrize("field,fid", [("bn254_fr", 0), ("curve25519_fr", 1)])
def test_offline_beaver_mul_and_open_check(field, fid):
    from ark_mpc_b200 import offline as off

    n = 777

    def party(f):
        E, src = f.engine, f.offline_phase
        x = off.ValueMacBatch(f, *src.next_shared_value_batch(n))
        y = off.ValueMacBatch(f, *src.next_shared_value_batch(n))
        t = tuple(off.ValueMacBatch(f, *p) for p in src.next_triplet_batch(n))
        xy = off.beaver_mul(f, x, y, t)
        opened = off.open_and_check_macs(f, xy)           # the product's MACs verify
        want = E.mul(off.open_and_check_macs(f, x), off.open_and_check_macs(f, y))
        return E.download(opened), E.download(want)

    (o0, w0), (o1, w1) = _run(party, field)
    assert np.array_equal(o0, w0) and np.array_equal(o1, w1) and np.array_equal(o0, o1)


def test_open_and_check_macs_rejects_a_corrupted_mac():
    from ark_mpc_b200 import offline as off

    def party(f):
        E, src = f.engine, f.offline_phase
        x = off.ValueMacBatch(f, *src.next_shared_value_batch(64))
        if f.party_id() == 1:
            x.mac[5, 0] += 1  # corrupt one limb of one MAC share
        try:
            off.open_and_check_macs(f, x)
        except off.InvalidMac:
            return "invalid mac"
        return "accepted"

    assert _run(party) == ("invalid mac", "invalid mac")


@pytest.mark.parametrize("n", [1, 500])
def test_sacrifice_accepts_good_triples_and_rejects_bad_ones(n):
    """triplets.rs:118-150: triples (a, b, c) and (a, b', c') sharing a; the check passes iff c = ab and c' = ab'."""
    from ark_mpc_b200 import offline as off

    def make(f, corrupt):
        E, src = f.engine, f.offline_phase
        key = src.key
        a = E.random(1, 0, n)
        b, bp = E.random(2, 0, n), E.random(3, 0, n)
        c, cp = E.mul(a, b), E.mul(a, bp)
        if corrupt:
            one = E.upload(np.array([[1, 0, 0, 0]], dtype=np.uint64))
            c = c.clone()
            c[n // 2:n // 2 + 1] = E.add(c[n // 2:n // 2 + 1].contiguous(), one)  # c != ab at one index (MACs stay consistent with the wrong c)
        sh = lambda v, s: off.ValueMacBatch(f, *src.share_of(v, s))
        _ = key
        return sh(a, 100), sh(b, 200), sh(c, 300), sh(bp, 400), sh(cp, 500)

    def party(f):
        out = []
        for corrupt in (False, True):
            try:
                off.sacrifice(f, *make(f, corrupt))
                out.append("ok")
            except off.SacrificeError:
                out.append("sacrifice error")
        return tuple(out)

    assert _run(party) == (("ok", "sacrifice error"), ("ok", "sacrifice error"))
