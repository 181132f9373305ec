"""Shared helpers for the parity tests: synthetic two-party data built with the ORACLE (CPU)."""
from __future__ import annotations

import numpy as np

from oracle import coracle as co
from oracle import pyoracle as po

FIELD_BY_ID = {0: po.BN254_FR, 1: po.CURVE25519_FR}
FIELD_NAME = {0: "bn254_fr", 1: "curve25519_fr"}


def aos(share: np.ndarray, mac: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(np.concatenate([share, mac], axis=1))


def split_aos(a: np.ndarray):
    return np.ascontiguousarray(a[:, :4]), np.ascontiguousarray(a[:, 4:])


def mont_scalar(fid: int, v: int) -> np.ndarray:
    return co.to_mont(fid, co.ints_to_limbs([v]))[0]


def authenticated(fid: int, seed: int, n: int, key: np.ndarray, values: np.ndarray | None = None):
    """Additive sharing of `values` (default: synthetic) and of key*values.  Returns (values, (s0,m0), (s1,m1))."""
    v = co.synth(fid, seed, 0, n) if values is None else values
    s0 = co.synth(fid, seed + 1, 0, n)
    s1 = co.scalar_sub(fid, v, s0)
    kv = co.scalar_mul(fid, v, np.tile(key, (n, 1)))
    m0 = co.synth(fid, seed + 2, 0, n)
    m1 = co.scalar_sub(fid, kv, m0)
    return v, (s0, m0), (s1, m1)


class TwoPartyData:
    """Everything both parties hold for one batch_mul: x, y, triple (a,b,c=ab), MAC key shares."""

    def __init__(self, fid: int, n: int, seed: int = 0xA11CE, edge: bool = False):
        self.fid, self.n = fid, n
        F = FIELD_BY_ID[fid]
        k0, k1 = co.synth(fid, seed + 100, 0, 1)[0], co.synth(fid, seed + 101, 0, 1)[0]
        self.keys = (k0, k1)
        self.key = co.scalar_add(fid, k0.reshape(1, 4), k1.reshape(1, 4))[0]
        xv = co.synth(fid, seed + 10, 0, n)
        yv = co.synth(fid, seed + 20, 0, n)
        av = co.synth(fid, seed + 30, 0, n)
        bv = co.synth(fid, seed + 40, 0, n)
        if edge and n >= 6:
            special = co.to_mont(fid, co.ints_to_limbs([0, 1, F.p - 1, 0, F.p - 1, 2]))
            xv[:6] = special
            yv[:6] = special[::-1]
            av[:3] = special[:3]
            bv[3:6] = special[:3]
        cv = co.scalar_mul(fid, av, bv)
        self.xv, self.yv, self.av, self.bv, self.cv = xv, yv, av, bv, cv
        self.x = authenticated(fid, seed + 11, n, self.key, xv)[1:]
        self.y = authenticated(fid, seed + 21, n, self.key, yv)[1:]
        self.a = authenticated(fid, seed + 31, n, self.key, av)[1:]
        self.b = authenticated(fid, seed + 41, n, self.key, bv)[1:]
        self.c = authenticated(fid, seed + 51, n, self.key, cv)[1:]

    def party(self, p: int):
        return dict(key=self.keys[p], x=self.x[p], y=self.y[p], a=self.a[p], b=self.b[p], c=self.c[p])

    def oracle_batch_mul(self, threads: int = 4):
        """Reference (unfused, AoS) two-party batch_mul on the CPU oracle."""
        g = lambda t: (aos(*t[0]), aos(*t[1]))
        return co.two_party_batch_mul(self.fid, threads, self.keys, g(self.x), g(self.y), g(self.a), g(self.b), g(self.c))
