"""Committed golden vectors for the point Beaver multiplication, batch inversion and the FFT (tests/golden/curve_golden.json,
made by tests/golden/make_golden_curve.py from the affine Python big-int oracle): the C oracle reproduces them on CPU, the
CUDA path reproduces them on the GPU (points compared in canonical affine form)."""
import json
import os

import numpy as np
import pytest

from oracle import coracle as co
from tests.test_golden import L, S
from tests.util_curve import points_from_affine, CURVE_BY_ID

HERE = os.path.dirname(os.path.abspath(__file__))
CID = {"bn254_g1": 0, "curve25519_edwards": 1}


def load():
    with open(os.path.join(HERE, "golden", "curve_golden.json")) as f:
        return json.load(f)


def XY(v):
    """list of [x_limbs, y_limbs] -> (n, 8) affine Montgomery limbs"""
    return np.ascontiguousarray(L(v).reshape(-1, 8))


def proj(cv, xy):
    """(n, 8) affine Montgomery limbs -> (n, words) projective image with Z = 1 (BN254 (0,0) -> identity)."""
    Cv = CURVE_BY_ID[cv]
    fq = Cv.fq
    pts = []
    for x, y in zip(co.limbs_to_ints(xy[:, :4]), co.limbs_to_ints(xy[:, 4:])):
        P = (fq.from_mont(x), fq.from_mont(y))
        pts.append(None if (Cv.kind == "sw" and P == (0, 0)) else P)
    return points_from_affine(cv, pts)


def pshares(cv, v):
    """list of [[x,y],[x,y]] -> (n, 2*words) PointShare image"""
    a = L(v)  # (n, 2, 2, 4)
    share = proj(cv, np.ascontiguousarray(a[:, 0].reshape(-1, 8)))
    mac = proj(cv, np.ascontiguousarray(a[:, 1].reshape(-1, 8)))
    return np.ascontiguousarray(np.concatenate([share, mac], axis=1))


def pshares_xy(v):
    a = L(v)
    return np.ascontiguousarray(a.reshape(a.shape[0] * 2, 8))


@pytest.mark.parametrize("idx", [0, 1])
def test_c_oracle_reproduces_point_golden(idx):
    g = load()["point_beaver"][idx]
    cv = CID[g["curve"]]
    keys = L(g["key_shares"])
    P = g["party"]
    x, a, b, c = [(S(P[0][k]), S(P[1][k])) for k in "xabc"]
    Ps = (pshares(cv, P[0]["P"]), pshares(cv, P[1]["P"]))
    o0, o1, d, E = co.two_party_point_mul(cv, 2, (keys[0], keys[1]), x, Ps, a, b, c)
    w = co.point_words(cv)
    assert np.array_equal(d, L(g["d_open"]))
    assert np.array_equal(co.pt_normalize(cv, E), XY(g["E_open"]))
    for p, o in ((0, o0), (1, o1)):
        assert np.array_equal(co.pt_normalize(cv, o.reshape(-1, w)), pshares_xy(P[p]["batch_mul"]))
    opened = co.pt_add(cv, np.ascontiguousarray(o0[:, :w]), np.ascontiguousarray(o1[:, :w]))
    assert np.array_equal(co.pt_normalize(cv, opened), XY(g["product_open"]))


@pytest.mark.parametrize("idx", [0, 1])
def test_c_oracle_reproduces_ntt_golden(idx):
    g = load()["ntt"][idx]
    x = L(g["x"])
    assert np.array_equal(co.fft(0, x), L(g["fft"]))
    assert np.array_equal(co.fft(0, x, inverse=True), L(g["ifft"]))
    assert np.array_equal(co.batch_inverse(0, x), L(g["inverse"]))


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_cuda_reproduces_point_golden(idx):
    from ark_mpc_b200.engine import Engine

    g = load()["point_beaver"][idx]
    cv = CID[g["curve"]]
    E = Engine(0, {0: "bn254_fr", 1: "curve25519_fr"}[cv])
    keys = L(g["key_shares"])
    P = g["party"]
    pl = lambda v: (E.upload(S(v)[:, :4]), E.upload(S(v)[:, 4:]))
    masks = []
    for p in (0, 1):
        x, a, b = (pl(P[p][k]) for k in "xab")
        d, Em = E.pt_beaver_mask(x[0], E.upload_points(pshares(cv, P[p]["P"])), a[0], b[0])
        assert np.array_equal(E.download(d), L(P[p]["d_mine"]))
        assert np.array_equal(E.download(E.pt_normalize(Em)), XY(P[p]["E_mine"]))
        masks.append((d, Em))
    outs = []
    for p in (0, 1):
        a, b, c = (pl(P[p][k]) for k in "abc")
        out, (do, Eo) = E.pt_beaver_recombine(p, keys[p], masks[p][0], masks[1 - p][0], masks[p][1], masks[1 - p][1], a, b, c, want_open=True)
        assert np.array_equal(E.download(E.pt_normalize(out)), pshares_xy(P[p]["batch_mul"]))
        assert np.array_equal(E.download(do), L(g["d_open"]))
        assert np.array_equal(E.download(E.pt_normalize(Eo)), XY(g["E_open"]))
        outs.append(out)
    opened = E.download(E.pt_normalize(E.pt_add(outs[0], outs[1]))).reshape(-1, 2, 8)[:, 0, :]
    assert np.array_equal(opened, XY(g["product_open"]))
    E.close()


@pytest.mark.gpu
@pytest.mark.parametrize("idx", [0, 1])
def test_cuda_reproduces_ntt_golden(idx):
    from ark_mpc_b200.engine import Engine

    g = load()["ntt"][idx]
    E = Engine(0, "bn254_fr")
    x = E.upload(L(g["x"]))
    assert np.array_equal(E.download(E.fft(x)), L(g["fft"]))
    assert np.array_equal(E.download(E.fft(x, inverse=True)), L(g["ifft"]))
    assert np.array_equal(E.download(E.batch_inverse(x)), L(g["inverse"]))
    E.close()
