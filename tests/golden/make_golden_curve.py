#!/usr/bin/env python
"""Generate tests/golden/curve_golden.json from the Python big-int oracle (oracle/pyoracle.py): point Beaver multiplication
(AuthenticatedPointResult::batch_mul, authenticated_curve.rs:682-714) on both curves in canonical AFFINE form, batch inversion
and the FFT on BN254 Fr.  Scalars are Montgomery images (4 LE u64 limbs, hex); points are affine (x, y) Montgomery images,
the BN254 identity is (0, 0).  Run: python tests/golden/make_golden_curve.py  (deterministic)."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def limbs(F, v):
    return ["%016x" % l for l in F.limbs(v)]


def aff(C, P):
    if P is None:
        P = (0, 0)
    return [limbs(C.fq, P[0]), limbs(C.fq, P[1])]


def point_case(C, seed, n):
    rng = random.Random(seed)
    F, r, G = C.fr, C.fr.p, C.generator
    keys = (rng.randrange(r), rng.randrange(r))
    key = sum(keys) % r
    xs = [0, 1, r - 1] + [rng.randrange(r) for _ in range(n - 3)]
    ss = [rng.randrange(1, r) for _ in range(n)]
    sh = lambda v: po.authenticated_split(F, v, key, rng)
    x, a, b, c, P = ([], []), ([], []), ([], []), ([], []), ([], [])
    for i in range(n):
        av, bv = rng.randrange(r), rng.randrange(r)
        for dst, v in ((x, xs[i]), (a, av), (b, bv), (c, av * bv % r)):
            s0, s1 = sh(v)
            dst[0].append(s0)
            dst[1].append(s1)
        p0, m0 = rng.randrange(r), rng.randrange(r)
        P[0].append((C.mul(G, p0), C.mul(G, m0)))
        P[1].append((C.mul(G, (ss[i] - p0) % r), C.mul(G, (key * ss[i] - m0) % r)))
    masks = [po.point_beaver_mask(C, x[p], P[p], a[p], b[p]) for p in (0, 1)]
    d = po.open_add(F, masks[0][0], masks[1][0])
    E = [C.add(u, v) for u, v in zip(masks[0][1], masks[1][1])]
    outs = [po.point_beaver_recombine(C, p, keys[p], d, E, a[p], b[p], c[p]) for p in (0, 1)]
    opened = [C.add(u[0], v[0]) for u, v in zip(outs[0], outs[1])]
    assert opened == [C.mul(C.mul(G, s), xv) for s, xv in zip(ss, xs)]
    shl = lambda s: [limbs(F, s[0]), limbs(F, s[1])]
    case = {"curve": C.name, "n": n, "key_shares": [limbs(F, k) for k in keys], "d_open": [limbs(F, v) for v in d],
            "E_open": [aff(C, e) for e in E], "product_open": [aff(C, o) for o in opened], "party": []}
    for p in (0, 1):
        case["party"].append({
            "x": [shl(s) for s in x[p]], "a": [shl(s) for s in a[p]], "b": [shl(s) for s in b[p]], "c": [shl(s) for s in c[p]],
            "P": [[aff(C, s[0]), aff(C, s[1])] for s in P[p]],
            "d_mine": [limbs(F, v) for v in masks[p][0]], "E_mine": [aff(C, e) for e in masks[p][1]],
            "batch_mul": [[aff(C, s[0]), aff(C, s[1])] for s in outs[p]],
        })
    return case


def ntt_case(seed, n):
    F = po.BN254_FR
    rng = random.Random(seed)
    xs = [0, 1, F.p - 1] + [rng.randrange(F.p) for _ in range(n - 3)]
    return {"field": F.name, "n": n, "x": [limbs(F, v) for v in xs], "fft": [limbs(F, v) for v in po.naive_dft(F, xs)],
            "ifft": [limbs(F, v) for v in po.naive_dft(F, xs, inverse=True)], "inverse": [limbs(F, v) for v in po.batch_inverse(F, xs)]}


if __name__ == "__main__":
    out = {"point_beaver": [point_case(po.BN254_G1, 101, 6), point_case(po.CURVE25519_EDWARDS, 102, 6)],
           "ntt": [ntt_case(201, 16), ntt_case(202, 64)]}
    with open(os.path.join(HERE, "curve_golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "curve_golden.json"))
