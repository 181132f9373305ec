"""Helpers for the point-gate tests: conversions between the affine big-int oracle and the projective limb images, and
synthetic two-party data for AuthenticatedPointResult::batch_mul built with the ORACLE (CPU)."""
from __future__ import annotations

import random

import numpy as np

from oracle import coracle as co
from oracle import pyoracle as po
from tests.util import aos, authenticated

CURVE_BY_ID = {0: po.BN254_G1, 1: po.CURVE25519_EDWARDS}
CURVE_NAME = {0: "bn254_g1", 1: "curve25519_edwards"}


def affine_ints(cv: int, pts: np.ndarray):
    """(n, words) projective limb image -> list of affine oracle points (None = BN254 identity)."""
    Cv = CURVE_BY_ID[cv]
    xy = co.pt_normalize(cv, pts)
    fq = Cv.fq
    out = []
    xs, ys = co.limbs_to_ints(xy[:, :4]), co.limbs_to_ints(xy[:, 4:])
    for x, y in zip(xs, ys):
        P = (fq.from_mont(x), fq.from_mont(y))
        out.append(None if (Cv.kind == "sw" and P == (0, 0)) else P)
    return out


def xy_to_affine(cv: int, xy: np.ndarray):
    """(n, 8) affine Montgomery limbs (e.g. from arkmpc_pt_normalize) -> oracle points."""
    Cv = CURVE_BY_ID[cv]
    xs, ys = co.limbs_to_ints(xy[:, :4]), co.limbs_to_ints(xy[:, 4:])
    out = []
    for x, y in zip(xs, ys):
        P = (Cv.fq.from_mont(x), Cv.fq.from_mont(y))
        out.append(None if (Cv.kind == "sw" and P == (0, 0)) else P)
    return out


def points_from_affine(cv: int, pts, rng: random.Random | None = None) -> np.ndarray:
    """Affine oracle points -> (n, words) projective image; with `rng` the representative is randomised (Z != 1)."""
    Cv = CURVE_BY_ID[cv]
    q = Cv.fq.p
    rows = []
    for P in pts:
        z = rng.randrange(1, q) if rng else 1
        if Cv.kind == "sw":
            c = [1, 1, 0] if P is None else [P[0] * z * z % q, P[1] * z * z * z % q, z]
        else:
            c = [P[0] * z % q, P[1] * z % q, P[0] * P[1] * z % q, z]
        rows.append([Cv.fq.to_mont(v) for v in c])
    flat = co.ints_to_limbs([v for row in rows for v in row])
    return np.ascontiguousarray(flat.reshape(len(pts), -1))


class TwoPartyPointData:
    """Everything both parties hold for one AuthenticatedPointResult::batch_mul: scalar shares x, PointShares of P = s*G,
    a triple (a, b, c = ab) and the MAC key shares.  Point shares are multiples of the generator, as in the reference's
    tests (lib.rs:48-54)."""

    def __init__(self, cv: int, n: int, seed: int = 0xC0FFEE):
        self.cv, self.n = cv, n
        fr = co.CURVE_FR[cv]
        self.fr = fr
        k0, k1 = co.synth(fr, seed + 100, 0, 1)[0], co.synth(fr, seed + 101, 0, 1)[0]
        self.keys = (k0, k1)
        self.key = co.scalar_add(fr, k0.reshape(1, 4), k1.reshape(1, 4))[0]
        self.xv = co.synth(fr, seed + 10, 0, n)
        sv = co.synth(fr, seed + 20, 0, n)
        av = co.synth(fr, seed + 30, 0, n)
        bv = co.synth(fr, seed + 40, 0, n)
        cv_ = co.scalar_mul(fr, av, bv)
        g = lambda t: (aos(*t[0]), aos(*t[1]))
        self.x = g(authenticated(fr, seed + 11, n, self.key, self.xv)[1:])
        self.a = g(authenticated(fr, seed + 31, n, self.key, av)[1:])
        self.b = g(authenticated(fr, seed + 41, n, self.key, bv)[1:])
        self.c = g(authenticated(fr, seed + 51, n, self.key, cv_)[1:])
        # P = s*G shared additively in the exponent: share_p = s_p*G, mac_p = m_p*G with s0+s1 = s, m0+m1 = key*s
        s_sh = authenticated(fr, seed + 21, n, self.key, sv)[1:]
        self.Pv = co.pt_mul_generator(cv, sv)
        self.P = tuple(np.ascontiguousarray(np.concatenate([co.pt_mul_generator(cv, s_sh[p][0]), co.pt_mul_generator(cv, s_sh[p][1])], axis=1))
                       for p in (0, 1))

    def oracle_point_mul(self, threads: int = 4):
        return co.two_party_point_mul(self.cv, threads, self.keys, self.x, self.P, self.a, self.b, self.c)
