"""Bit-exact check of the per-element point gates (ark_mpc_b200/csrc/curve.cuh, curve_gates.cuh) on the CPU.

The device headers are compiled by g++ with the carry flag emulated (tests/host_emu/emu.cpp); projective outputs are
brought to affine form and compared with the exact affine big-int oracle (oracle/pyoracle.py), which is the
representation parity with the reference is defined on (projective representatives are not unique,
/root/reference/online-phase/src/algebra/curve/curve.rs:46)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CURVES = [(po.BN254_G1, 0), (po.CURVE25519_EDWARDS, 1)]


def to_proj(Cv, P, rng):
    """Affine oracle point -> the reference's projective memory image (ints, plain domain) with a random Z."""
    q = Cv.fq.p
    z = rng.randrange(1, q)
    if Cv.kind == "sw":
        if P is None:
            return [rng.randrange(1, q), rng.randrange(1, q), 0]
        return [P[0] * z * z % q, P[1] * z * z * z % q, z]
    x, y = P
    return [x * z % q, y * z % q, x * y * z % q, z]


def from_proj(Cv, c):
    q = Cv.fq.p
    if Cv.kind == "sw":
        X, Y, Z = c
        if Z == 0:
            return None
        zi = pow(Z, -1, q)
        return (X * zi * zi % q, Y * zi * zi * zi % q)
    X, Y, T, Z = c
    zi = pow(Z, -1, q)
    assert T * Z % q == X * Y % q, "extended coordinate T inconsistent"
    return (X * zi % q, Y * zi % q)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "host_emu", "libark_emu.so")
    csrc = os.path.join(ROOT, "ark_mpc_b200", "csrc")
    srcs = [os.path.join(HERE, "host_emu", "emu.cpp")] + [os.path.join(csrc, f) for f in ("fp256.cuh", "f25519.cuh", "beaver.cuh", "curve.cuh", "curve_gates.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_violations.restype = C.c_uint64

    def run(Cv, cid, op, scalars, points, n_out_fe, party=0):
        """scalars: Fr ints (plain) ; points: projective coordinate lists (plain Fq ints).  Everything is passed in Montgomery form."""
        vals = [Cv.fr.to_mont(s) for s in scalars] + [Cv.fq.to_mont(c) for P in points for c in P]
        buf = np.zeros(8 * max(len(vals), 1), dtype=np.uint32)
        for k, v in enumerate(vals):
            for j in range(8):
                buf[8 * k + j] = (v >> (32 * j)) & 0xFFFFFFFF
        out = np.zeros(8 * n_out_fe, dtype=np.uint32)
        assert lib.emu_curve(cid, op, party, buf.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert lib.emu_violations() == 0
        return [sum(int(out[8 * k + j]) << (32 * j) for j in range(8)) for k in range(n_out_fe)]

    return run


def fq_out(Cv, raw):
    assert all(r < Cv.fq.p for r in raw), "non-canonical coordinate"
    return [Cv.fq.from_mont(r) for r in raw]


def rand_point(Cv, rng):
    return Cv.mul(Cv.generator, rng.randrange(1, Cv.fr.p))


@pytest.mark.parametrize("Cv,cid", CURVES)
def test_constants_and_group_law(emu, Cv, cid):
    rng = random.Random(7 + cid)
    K = 3 if Cv.kind == "sw" else 4
    out = fq_out(Cv, emu(Cv, cid, 11, [], [], 2 * K))
    assert from_proj(Cv, out[:K]) == Cv.generator
    assert from_proj(Cv, out[K:]) == Cv.identity
    P, Q = rand_point(Cv, rng), rand_point(Cv, rng)
    ident = Cv.identity
    cases = [(P, Q), (P, P), (P, Cv.neg(P)), (ident, P), (P, ident), (ident, ident)]
    for A, B in cases:
        got = from_proj(Cv, fq_out(Cv, emu(Cv, cid, 0, [], [to_proj(Cv, A, rng), to_proj(Cv, B, rng)], K)))
        assert got == Cv.add(A, B)
    for A in (P, Q, ident):
        assert from_proj(Cv, fq_out(Cv, emu(Cv, cid, 1, [], [to_proj(Cv, A, rng)], K))) == Cv.add(A, A)
        assert from_proj(Cv, fq_out(Cv, emu(Cv, cid, 12, [], [to_proj(Cv, A, rng)], K))) == Cv.neg(A)
        x, y = fq_out(Cv, emu(Cv, cid, 2, [], [to_proj(Cv, A, rng)], 2))
        assert (x, y) == (A if A is not None else (0, 0))


@pytest.mark.parametrize("Cv,cid", CURVES)
def test_scalar_mul(emu, Cv, cid):
    rng = random.Random(11 + cid)
    K = 3 if Cv.kind == "sw" else 4
    r = Cv.fr.p
    P = rand_point(Cv, rng)
    glv_lambda = 4407920970296243842393367215006156084916469457145843978461  # BN254: exercises the GLV split k = k1 + k2*lambda
    for s in [0, 1, 2, 15, 16, 17, r - 1, r - 2, (1 << 252) - 1, glv_lambda, glv_lambda + 1, r - glv_lambda, (1 << 128) - 1, 1 << 128,
              rng.randrange(r), rng.randrange(r), rng.randrange(r), rng.randrange(r)]:
        s %= r
        got = from_proj(Cv, fq_out(Cv, emu(Cv, cid, 3, [s], [to_proj(Cv, P, rng)], K)))
        assert got == Cv.mul(P, s), f"var-base s={s:x}"
        got = from_proj(Cv, fq_out(Cv, emu(Cv, cid, 4, [s], [], K)))
        assert got == Cv.mul(Cv.generator, s), f"fixed-base s={s:x}"
    # identity base
    got = from_proj(Cv, fq_out(Cv, emu(Cv, cid, 3, [12345], [to_proj(Cv, Cv.identity, rng)], K)))
    assert got == Cv.identity
    # the two-pass form (tables of P and of 2^128 P / 2^68 P, signed windows): scalars around the split point and at both ends
    pairs = [(rng.randrange(r), rng.randrange(r)), (0, 1), (r - 1, r - 2), ((1 << 128) - 1, 1 << 128), ((1 << 127), (1 << 128) + 1),
             (8, 9), (0x8888888888888888, 0x7777777777777777), ((1 << 252) - 1, glv_lambda), (r - glv_lambda, glv_lambda + 1),
             ((1 << 68) - 1, 1 << 68), (int("8" * 63, 16) % r, int("7" * 63, 16) % r)]
    for s0, s1 in pairs:
        s0, s1 = s0 % r, s1 % r
        out = fq_out(Cv, emu(Cv, cid, 10, [s0, s1], [to_proj(Cv, P, rng)], 2 * K))
        assert from_proj(Cv, out[:K]) == Cv.mul(P, s0) and from_proj(Cv, out[K:]) == Cv.mul(P, s1), f"two-pass s0={s0:x} s1={s1:x}"
    out = fq_out(Cv, emu(Cv, cid, 10, [5, r - 5], [to_proj(Cv, Cv.identity, rng)], 2 * K))
    assert from_proj(Cv, out[:K]) == Cv.identity and from_proj(Cv, out[K:]) == Cv.identity


@pytest.mark.parametrize("Cv,cid", CURVES)
def test_point_beaver_gate(emu, Cv, cid):
    """authenticated_curve.rs:682-714 for both parties on one element at a time, against the unfused oracle."""
    rng = random.Random(23 + cid)
    K = 3 if Cv.kind == "sw" else 4
    F, r, G = Cv.fr, Cv.fr.p, Cv.generator
    for trial in range(3):
        keys = (rng.randrange(r), rng.randrange(r))
        key = sum(keys) % r
        xv, av, bv = (rng.randrange(r) for _ in range(3))
        if trial == 2:
            xv = av  # d = 0
        cv = av * bv % r
        sp = rng.randrange(r)
        Pv = Cv.mul(G, sp)
        sh = lambda v: po.authenticated_split(F, v, key, rng)
        x, a, b, c = sh(xv), sh(av), sh(bv), sh(cv)
        # additive point sharing of P and key*P
        p0, m0 = rng.randrange(r), rng.randrange(r)
        Psh = ((Cv.mul(G, p0), Cv.mul(G, m0)), (Cv.mul(G, (sp - p0) % r), Cv.mul(G, (key * sp - m0) % r)))
        masks = []
        for p in (0, 1):
            want_d, want_E = po.point_beaver_mask(Cv, [x[p]], [Psh[p]], [a[p]], [b[p]])
            out = emu(Cv, cid, 5, [x[p][0], a[p][0], b[p][0]], [to_proj(Cv, Psh[p][0], rng)], 1 + K)
            assert F.from_mont(out[0]) == want_d[0]
            Em = fq_out(Cv, out[1:])
            assert from_proj(Cv, Em) == want_E[0]
            masks.append((want_d[0], Em))
        d = (masks[0][0] + masks[1][0]) % r
        E = Cv.add(from_proj(Cv, masks[0][1]), from_proj(Cv, masks[1][1]))
        opened = []
        for p in (0, 1):
            want = po.point_beaver_recombine(Cv, p, keys[p], [d], [E], [a[p]], [b[p]], [c[p]])[0]
            sc = [keys[p], masks[p][0], masks[1 - p][0], a[p][0], a[p][1], b[p][0], b[p][1], c[p][0], c[p][1]]
            out = emu(Cv, cid, 6, sc, [masks[p][1], masks[1 - p][1]], 1 + 3 * K, party=p)
            out_dual = emu(Cv, cid, 13, sc, [masks[p][1], masks[1 - p][1]], 1 + 3 * K, party=p)
            assert [from_proj(Cv, fq_out(Cv, out_dual[1 + i * K:1 + (i + 1) * K])) for i in range(3)] == \
                [from_proj(Cv, fq_out(Cv, out[1 + i * K:1 + (i + 1) * K])) for i in range(3)], "dual-chain variant differs"
            assert F.from_mont(out[0]) == d
            pts = fq_out(Cv, out[1:])
            assert from_proj(Cv, pts[:K]) == E
            assert from_proj(Cv, pts[K:2 * K]) == want[0], f"share differs (party {p})"
            assert from_proj(Cv, pts[2 * K:]) == want[1], f"mac differs (party {p})"
            opened.append(want)
        # protocol sanity: opened product and MAC
        xP = Cv.mul(Pv, xv)
        assert Cv.add(opened[0][0], opened[1][0]) == xP
        assert Cv.add(opened[0][1], opened[1][1]) == Cv.mul(xP, key)


@pytest.mark.parametrize("Cv,cid", CURVES)
def test_point_share_public_gates(emu, Cv, cid):
    rng = random.Random(31 + cid)
    K = 3 if Cv.kind == "sw" else 4
    r = Cv.fr.p
    key = rng.randrange(r)
    S, M, P = rand_point(Cv, rng), rand_point(Cv, rng), rand_point(Cv, rng)
    for party in (0, 1):
        for op, fn in ((7, po.pshare_add_public), (8, lambda C_, a, Pt, k, pid: po.pshare_add_public(C_, a, C_.neg(Pt), k, pid))):
            out = fq_out(Cv, emu(Cv, cid, op, [key], [to_proj(Cv, S, rng), to_proj(Cv, M, rng), to_proj(Cv, P, rng)], 2 * K, party=party))
            want = fn(Cv, (S, M), P, key, party)
            assert (from_proj(Cv, out[:K]), from_proj(Cv, out[K:])) == want
    out = fq_out(Cv, emu(Cv, cid, 9, [key], [to_proj(Cv, P, rng), to_proj(Cv, M, rng)], K))
    assert from_proj(Cv, out) == Cv.sub(Cv.mul(P, key), M)


def test_f25519_special_form_field(emu):
    """f25519.cuh: arithmetic modulo 2p on loosely reduced 256-bit values, checked against big-int arithmetic mod p."""
    so = os.path.join(HERE, "host_emu", "libark_emu.so")
    lib = C.CDLL(so)
    lib.emu_violations.restype = C.c_uint64
    p = (1 << 255) - 19
    rng = random.Random(2519)

    def run(op, ins, n_out=1):
        buf = np.zeros(8 * max(len(ins), 1), dtype=np.uint32)
        for k, v in enumerate(ins):
            for j in range(8):
                buf[8 * k + j] = (v >> (32 * j)) & 0xFFFFFFFF
        out = np.zeros(8 * n_out, dtype=np.uint32)
        assert lib.emu_f25519(op, buf.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert lib.emu_violations() == 0
        return [sum(int(out[8 * k + j]) << (32 * j) for j in range(8)) for k in range(n_out)]

    top = (1 << 256) - 1
    edge = [0, 1, 18, 19, 37, 38, 39, p - 1, p, p + 1, 2 * p - 1, 2 * p, 2 * p + 1, top, top - 37, top - 38, 1 << 255, (1 << 255) - 1,
            (1 << 255) + 18, (1 << 255) + 19, (1 << 32) - 1, 1 << 224]
    vals = edge + [rng.randrange(1 << 256) for _ in range(60)]
    for a in vals:
        assert run(4, [a])[0] == a % p                                   # canon
        assert run(3, [a])[0] % p == (-a) % p                             # neg
        assert run(7, [a])[0] == a * 38 % p                               # to_image is canonical
        assert run(6, [a])[0] % p == a * pow(38, -1, p) % p               # from_image
        z = run(8, [a, a])[0]
        assert (z & 0xFFFFFFFF) == (1 if a % p == 0 else 0) and ((z >> 32) & 0xFFFFFFFF) == 1
    for _ in range(400):
        a, b = rng.choice(vals), rng.choice(vals)
        assert run(0, [a, b])[0] % p == (a + b) % p
        assert run(1, [a, b])[0] % p == (a - b) % p
        assert run(2, [a, b])[0] % p == (a * b) % p
        z = run(8, [a, b])[0]
        assert ((z >> 32) & 0xFFFFFFFF) == (1 if (a - b) % p == 0 else 0)
    hard = [top, top - 1, (1 << 256) - (1 << 32), int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), int("80000000" * 8, 16),
            int("7fffffff" * 8, 16), int("ffffffff" * 7 + "00000000", 16)]
    for a in vals + hard:
        lo, hi = run(10, [a], 2)
        assert lo + (hi << 256) == a * a                                   # the 512-bit square itself
        assert run(9, [a])[0] % p == a * a % p
    for a in vals[:12] + vals[-6:]:
        inv = run(5, [a])[0]
        assert inv % p == (pow(a, -1, p) if a % p else 0)


@pytest.mark.parametrize("Cv,cid", CURVES)
def test_point_validation_on_curve_and_subgroup(emu, Cv, cid):
    """pt_valid_elem (arkmpc_pt_validate): what arkworks' deserialisation checks on peer points (curve.rs:105-135) and what the
    regrouped recombination needs of E_peer.  Curve25519 has cofactor 8: a point with a torsion component is ON the curve but
    must be rejected; ((a + d) mod r) * E differs from a*E + d*E exactly for such points."""
    rng = random.Random(99 + cid)
    q, r = Cv.fq.p, Cv.fr.p
    valid = lambda c: emu(Cv, cid, 14, [], [c], 1)[0] & 1
    for _ in range(4):
        assert valid(to_proj(Cv, rand_point(Cv, rng), rng)) == 1
    assert valid(to_proj(Cv, Cv.identity, rng)) == 1
    assert valid(to_proj(Cv, Cv.generator, rng)) == 1
    # off the curve
    P = rand_point(Cv, rng)
    bad = to_proj(Cv, P, rng)
    bad[0] = (bad[0] + 1) % q
    if Cv.kind != "sw":
        bad[2] = bad[0] * bad[1] % q * pow(bad[3], -1, q) % q  # keep T consistent so that only the curve equation fails
    assert valid(bad) == 0
    if Cv.kind != "sw":
        # inconsistent T
        c = to_proj(Cv, P, rng)
        c[2] = (c[2] + 1) % q
        assert valid(c) == 0
        # torsion: (0, -1) has order 2, (sqrt(-1), 0) has order 4; P + T is on the curve and outside the prime-order subgroup
        T2 = (0, q - 1)
        i = pow(2, (q - 1) // 4, q)
        T4 = (i, 0)
        for T in (T2, T4):
            assert Cv.mul(T, 8) == Cv.identity and T != Cv.identity
            assert valid(to_proj(Cv, T, rng)) == 0
            PT = Cv.add(P, T)
            assert valid(to_proj(Cv, PT, rng)) == 0
            # the reason the guard exists: scalars combined mod r do not act on the torsion component like separate products
            a, d = r - 1, 2
            assert Cv.mul(PT, (a + d) % r) != Cv.add(Cv.mul(PT, a), Cv.mul(PT, d))
