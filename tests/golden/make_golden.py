#!/usr/bin/env python
"""Generate tests/golden/*.json from the Python big-int oracle (oracle/pyoracle.py).

The reference ships no golden vectors for this path and cannot be built here (Rust, un-vendored arkworks), so the
fixtures pin the C oracle and the CUDA kernels to the big-int restatement; the restatement itself is pinned in
tests/test_oracle.py.  Values are Montgomery memory images (4 LE u64 limbs as hex strings).
Run:  python tests/golden/make_golden.py   (deterministic; rewrites the files in place)"""
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


def share_l(F, s):
    return [limbs(F, s[0]), limbs(F, s[1])]


def scalar_case(F, seed, n):
    rng = random.Random(seed)
    src = po.RandomBeaverSource(F, seed)
    edge = [0, 1, F.p - 1, 2, F.p - 2, (F.p + 1) // 2]
    xs = edge + [rng.randrange(F.p) for _ in range(n - len(edge))]
    ys = edge[::-1] + [rng.randrange(F.p) for _ in range(n - len(edge))]
    x = src.share_values(xs)
    y = src.share_values(ys)
    trip = src.triples(n)
    outs, (d, e), de = po.two_party_batch_mul(F, src.key_shares, x, y, trip)
    opened = po.open_shares(F, outs[0], outs[1])
    assert opened == [a * b % F.p for a, b in zip(xs, ys)]
    pub = [rng.randrange(F.p) for _ in range(n)]
    case = {
        "field": F.name, "n": n, "key_shares": [limbs(F, k) for k in src.key_shares],
        "x_plain": [limbs(F, v) for v in xs], "y_plain": [limbs(F, v) for v in ys], "public": [limbs(F, v) for v in pub],
        "party": [],
        "d_open": [limbs(F, v) for v in d], "e_open": [limbs(F, v) for v in e],
        "product_open": [limbs(F, v) for v in opened],
    }
    for p in (0, 1):
        k = src.key_shares[p]
        case["party"].append({
            "x": [share_l(F, s) for s in x[p]], "y": [share_l(F, s) for s in y[p]],
            "a": [share_l(F, s) for s in trip[p][0]], "b": [share_l(F, s) for s in trip[p][1]], "c": [share_l(F, s) for s in trip[p][2]],
            "d_mine": [limbs(F, v) for v in de[p][0]], "e_mine": [limbs(F, v) for v in de[p][1]],
            "batch_mul": [share_l(F, s) for s in outs[p]],
            "add": [share_l(F, po.share_add(F, u, v)) for u, v in zip(x[p], y[p])],
            "sub": [share_l(F, po.share_sub(F, u, v)) for u, v in zip(x[p], y[p])],
            "neg": [share_l(F, po.share_neg(F, u)) for u in x[p]],
            "mul_public": [share_l(F, po.share_mul_public(F, u, v)) for u, v in zip(x[p], pub)],
            "add_public": [share_l(F, po.share_add_public(F, u, v, k, p)) for u, v in zip(x[p], pub)],
            "sub_public": [share_l(F, po.share_sub_public(F, u, v, k, p)) for u, v in zip(x[p], pub)],
            "sum": share_l(F, po.share_sum(F, x[p])),
            "mac_check": [limbs(F, v) for v in po.mac_check_shares(F, k, opened, outs[p])],
        })
    chk = [po.mac_check_shares(F, src.key_shares[p], opened, outs[p]) for p in (0, 1)]
    assert all((u + v) % F.p == 0 for u, v in zip(*chk))
    case["commit_blinder"] = limbs(F, 11)
    case["commit_party0"] = limbs(F, po.hash_commit(F, chk[0], 11))
    case["bytes_be_x0"] = F.to_bytes_be(xs[-1]).hex()
    return case


def party_id_kat(F):
    """offline_prep.rs:137-158: 2*3 = 6 under key 1 with the mock source."""
    srcs = [po.PartyIDBeaverSource(F, p) for p in (0, 1)]
    keys = (0, 1)
    n = 4
    vals_x, vals_y = [5, 0, F.p - 1, 7], [9, 3, F.p - 1, 0]
    x = po.two_party_share_scalars(F, vals_x, 0, srcs[0], srcs[1], keys)
    y = po.two_party_share_scalars(F, vals_y, 0, srcs[0], srcs[1], keys)
    trip = [s.next_triplet_batch(n) for s in srcs]
    outs, _, _ = po.two_party_batch_mul(F, keys, x, y, trip)
    opened, ok = po.two_party_open_authenticated(F, keys, outs[0], outs[1])
    assert ok and opened == [a * b % F.p for a, b in zip(vals_x, vals_y)]
    return {"field": F.name, "x_plain": [limbs(F, v) for v in vals_x], "y_plain": [limbs(F, v) for v in vals_y],
            "x": [[share_l(F, s) for s in x[p]] for p in (0, 1)], "y": [[share_l(F, s) for s in y[p]] for p in (0, 1)],
            "triples": [[[share_l(F, s) for s in t] for t in trip[p]] for p in (0, 1)],
            "batch_mul": [[share_l(F, s) for s in outs[p]] for p in (0, 1)], "opened": [limbs(F, v) for v in opened]}


def main():
    out = {"scalar": [scalar_case(po.BN254_FR, 20241, 24), scalar_case(po.CURVE25519_FR, 20242, 24)],
           "party_id_beaver_source": [party_id_kat(po.BN254_FR), party_id_kat(po.CURVE25519_FR)]}
    with open(os.path.join(HERE, "scalar_golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote scalar_golden.json", os.path.getsize(os.path.join(HERE, "scalar_golden.json")), "bytes")


if __name__ == "__main__":
    main()
