"""Committed golden vectors (tests/golden/scalar_golden.json, made by tests/golden/make_golden.py from the Python
big-int oracle): the C oracle must reproduce them on CPU, the CUDA path must reproduce them on the GPU."""
import json
import os

import numpy as np
import pytest

from oracle import coracle as co

HERE = os.path.dirname(os.path.abspath(__file__))
FID = {"bn254_fr": 0, "curve25519_fr": 1}


def load():
    with open(os.path.join(HERE, "golden", "scalar_golden.json")) as f:
        return json.load(f)


def L(v):
    """nested hex-limb lists -> uint64 array with trailing dim 4"""
    a = np.array(v, dtype=object)
    flat = np.array([int(h, 16) for h in a.reshape(-1)], dtype=np.uint64)
    return flat.reshape(a.shape)


def S(v):
    """list of [share_limbs, mac_limbs] -> (n, 8) AoS"""
    return np.ascontiguousarray(L(v).reshape(-1, 8))


@pytest.mark.parametrize("idx", [0, 1])
def test_c_oracle_reproduces_golden(idx):
    g = load()["scalar"][idx]
    fid = FID[g["field"]]
    keys = L(g["key_shares"])
    P = g["party"]
    x, y, a, b, c = [(S(P[0][k]), S(P[1][k])) for k in "xyabc"]
    o0, o1, d, e = co.two_party_batch_mul(fid, 2, (keys[0], keys[1]), x, y, a, b, c)
    assert np.array_equal(o0, S(P[0]["batch_mul"])) and np.array_equal(o1, S(P[1]["batch_mul"]))
    assert np.array_equal(d, L(g["d_open"])) and np.array_equal(e, L(g["e_open"]))
    pub = L(g["public"])
    for p in (0, 1):
        assert np.array_equal(co.batch_add(fid, x[p], y[p]), S(P[p]["add"]))
        assert np.array_equal(co.batch_sub(fid, x[p], y[p]), S(P[p]["sub"]))
        assert np.array_equal(co.batch_neg(fid, x[p]), S(P[p]["neg"]))
        assert np.array_equal(co.batch_mul_public(fid, x[p], pub), S(P[p]["mul_public"]))
        assert np.array_equal(co.batch_add_public(fid, p, keys[p], x[p], pub), S(P[p]["add_public"]))
        assert np.array_equal(co.batch_add_public(fid, p, keys[p], x[p], pub, sub=True), S(P[p]["sub_public"]))
        assert np.array_equal(co.share_sum(fid, x[p]), L(P[p]["sum"]).reshape(8))
        assert np.array_equal(co.mac_check(fid, keys[p], L(g["product_open"]), S(P[p]["batch_mul"])), L(P[p]["mac_check"]))
        dm, em = co.beaver_mask(fid, x[p], y[p], a[p], b[p])
        assert np.array_equal(dm, L(P[p]["d_mine"])) and np.array_equal(em, L(P[p]["e_mine"]))


@pytest.mark.parametrize("idx", [0, 1])
def test_c_oracle_party_id_beaver_source_kat(idx):
    g = load()["party_id_beaver_source"][idx]
    fid = FID[g["field"]]
    keys = co.to_mont(fid, co.ints_to_limbs([0, 1]))  # key share = party id (offline_prep.rs:109-111)
    x, y = [(S(g[k][0]), S(g[k][1])) for k in "xy"]
    a, b, c = [(S(g["triples"][0][k]), S(g["triples"][1][k])) for k in range(3)]
    o0, o1, _, _ = co.two_party_batch_mul(fid, 1, (keys[0], keys[1]), x, y, a, b, c)
    assert np.array_equal(o0, S(g["batch_mul"][0])) and np.array_equal(o1, S(g["batch_mul"][1]))
    opened = co.scalar_add(fid, o0[:, :4], o1[:, :4])
    assert np.array_equal(opened, L(g["opened"]))
    assert np.array_equal(opened, co.scalar_mul(fid, L(g["x_plain"]), L(g["y_plain"])))


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_cuda_reproduces_golden(idx):
    from ark_mpc_b200.engine import Engine

    g = load()["scalar"][idx]
    E = Engine(0, g["field"])
    keys = L(g["key_shares"])
    P = g["party"]
    pl = lambda v: (E.upload(S(v)[:, :4]), E.upload(S(v)[:, 4:]))
    dn = lambda t: np.concatenate([E.download(t[0]), E.download(t[1])], axis=1)
    pub = E.upload(L(g["public"]))
    masks = []
    for p in (0, 1):
        x, y, a, b = (pl(P[p][k]) for k in "xyab")
        d, e = E.beaver_mask(x[0], y[0], a[0], b[0])
        assert np.array_equal(E.download(d), L(P[p]["d_mine"])) and np.array_equal(E.download(e), L(P[p]["e_mine"]))
        masks.append((d, e))
    for p in (0, 1):
        x, y, a, b, c = (pl(P[p][k]) for k in "xyabc")
        out, (do, eo) = E.beaver_recombine(p, keys[p], masks[p][0], masks[p][1], masks[1 - p][0], masks[1 - p][1], a, b, c, want_open=True)
        assert np.array_equal(dn(out), S(P[p]["batch_mul"]))
        assert np.array_equal(E.download(do), L(g["d_open"])) and np.array_equal(E.download(eo), L(g["e_open"]))
        assert np.array_equal(dn(E.share_add(x, y)), S(P[p]["add"]))
        assert np.array_equal(dn(E.share_sub(x, y)), S(P[p]["sub"]))
        assert np.array_equal(dn(E.share_neg(x)), S(P[p]["neg"]))
        assert np.array_equal(dn(E.share_mul_public(x, pub)), S(P[p]["mul_public"]))
        assert np.array_equal(dn(E.share_add_public(p, keys[p], x, pub)), S(P[p]["add_public"]))
        assert np.array_equal(dn(E.share_add_public(p, keys[p], x, pub, sub=True)), S(P[p]["sub_public"]))
        assert np.array_equal(dn(E.share_sum(x)).reshape(8), L(P[p]["sum"]).reshape(8))
        chk = E.mac_check(keys[p], E.upload(L(g["product_open"])), out[1])
        assert np.array_equal(E.download(chk), L(P[p]["mac_check"]))
    last = E.upload(L(g["x_plain"])[-1:])
    assert bytes(E.to_bytes_be(last).cpu().numpy().reshape(-1)).hex() == g["bytes_be_x0"]
    E.close()


def test_fixtures_say_who_generated_them():
    """The committed fixtures come from oracle/pyoracle.py (tests/golden/make_golden*.py), NOT from a run of the reference: the
    reference cannot be built in this image.  oracle/ref_recipe/ holds the recipe that produces a reference-run file."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for gen in ("make_golden.py", "make_golden_curve.py"):
        assert "oracle" in open(os.path.join(root, "tests", "golden", gen)).read()
    for f in ("regen.sh", "gen_golden.rs", "compare_reference.py", "README.md"):
        assert os.path.exists(os.path.join(root, "oracle", "ref_recipe", f))


def test_reference_run_fixture_when_present():
    """If a maintainer has run oracle/ref_recipe/regen.sh, tests/golden/reference_party_id.json holds the REFERENCE's own outputs
    for the PartyIDBeaverSource case: they must equal the oracle-generated fixture (that equality is what pins parity)."""
    import importlib.util

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = os.path.join(root, "tests", "golden", "reference_party_id.json")
    if not os.path.exists(ref):
        pytest.skip("parity unpinned: no reference-run fixture (the reference is Rust and cannot be built in this image; see oracle/ref_recipe)")
    spec = importlib.util.spec_from_file_location("compare_reference", os.path.join(root, "oracle", "ref_recipe", "compare_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.compare(ref) == []
