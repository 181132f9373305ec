"""Bit-exact check of the per-element DEVICE functions (ark_mpc_b200/csrc/fp256.cuh, beaver.cuh) on the CPU.

The same headers are compiled by g++ with the PTX carry flag emulated in software
(tests/host_emu/emu.cpp); every result is compared with the exact big-int oracle.  This is what
lets the carry-chain Montgomery code be verified on a box with no GPU; the GPU tests then only
have to show that nvcc/ptxas compile the same sequence faithfully."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "host_emu", "libark_emu.so")
    srcs = [os.path.join(HERE, "host_emu", "emu.cpp"), os.path.join(ROOT, "ark_mpc_b200", "csrc", "fp256.cuh"),
            os.path.join(ROOT, "ark_mpc_b200", "csrc", "beaver.cuh"),
            os.path.join(ROOT, "ark_mpc_b200", "csrc", "curve.cuh"), os.path.join(ROOT, "ark_mpc_b200", "csrc", "curve_gates.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_violations.restype = C.c_uint64

    def run(field, op, ins, n_out, party=0):
        buf = np.zeros(8 * len(ins), dtype=np.uint32)
        for k, v in enumerate(ins):
            for j in range(8):
                buf[8 * k + j] = (v >> (32 * j)) & 0xFFFFFFFF
        out = np.zeros(8 * n_out, dtype=np.uint32)
        rc = lib.emu_run(field, op, party, buf.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert lib.emu_violations() == 0, "an accumulator chain carried out (bound analysis violated)"
        return [sum(int(out[8 * k + j]) << (32 * j) for j in range(8)) for k in range(n_out)]

    return run


FIELDS = [("bn254_fr", 0), ("curve25519_fr", 1), ("bn254_fq", 2), ("curve25519_fq", 3)]
LAZY = FIELDS[:3]


def samples(F, rng, n):
    edge = [0, 1, 2, F.p - 1, F.p - 2, F.r, (F.p - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 224) - 1, F.p - (1 << 32)]
    return [e % F.p for e in edge] + [rng.randrange(F.p) for _ in range(n)]


@pytest.mark.parametrize("name,fid", FIELDS)
def test_constants(emu, name, fid):
    F = po.FIELDS[name]
    one, r2 = emu(fid, 12, [], 2)
    assert one == F.r and r2 == F.r2


@pytest.mark.parametrize("name,fid", FIELDS)
def test_add_sub_neg_mul(emu, name, fid):
    F = po.FIELDS[name]
    rng = random.Random(fid)
    A = samples(F, rng, 200)
    B = list(reversed(samples(F, rng, 200)))
    for a, b in zip(A, B):
        assert emu(fid, 0, [a, b], 1)[0] == (a + b) % F.p
        assert emu(fid, 1, [a, b], 1)[0] == (a - b) % F.p
        assert emu(fid, 2, [a], 1)[0] == (-a) % F.p
        # Montgomery product of images: a*b/R
        assert emu(fid, 3, [a, b], 1)[0] == a * b * F.rinv % F.p


@pytest.mark.parametrize("name,fid", FIELDS)
def test_mul_lazy_bounds(emu, name, fid):
    """mul_lazy(a, x): a canonical (or < 2^256 - p), x ANY 256-bit value; result ≡ a*x/R and < a*x/R + p."""
    F = po.FIELDS[name]
    rng = random.Random(10 + fid)
    for _ in range(200):
        a = rng.randrange(F.p)
        x = rng.choice([rng.randrange(1 << 256), (1 << 256) - 1, 3 * F.p % (1 << 256), rng.randrange(F.p)])
        r = emu(fid, 4, [a, x], 1)[0]
        assert r % F.p == a * x * F.rinv % F.p
        assert r < (a * x >> 256) + F.p + 1 and r < (1 << 256)
    # worst-case row operand allowed for lazy inputs on the 4p<2^256 fields
    if name != "curve25519_fq":
        a = (1 << 256) - F.p - 1
        r = emu(fid, 4, [a, (1 << 256) - 1], 1)[0]
        assert r % F.p == a * ((1 << 256) - 1) * F.rinv % F.p


@pytest.mark.parametrize("name,fid", LAZY)
def test_mul2_lazy(emu, name, fid):
    F = po.FIELDS[name]
    rng = random.Random(20 + fid)
    cases = [(F.p - 1, 3 * F.p - 1, F.p - 1, F.p - 1), (F.p - 1, (1 << 256) - 1, F.p - 1, (1 << 256) - 1), (0, 0, 0, 0)]
    for _ in range(300):
        cases.append((rng.randrange(F.p), rng.randrange(3 * F.p), rng.randrange(F.p), rng.randrange(F.p)))
    for a, x, b, y in cases:
        r = emu(fid, 5, [a, x, b, y], 1)[0]
        assert r % F.p == (a * x + b * y) * F.rinv % F.p
        assert r < ((a * x + b * y) >> 256) + F.p + 1
        if x < 3 * F.p and y < F.p:
            assert r < 2 * F.p  # one conditional subtraction suffices (kLazy2)


@pytest.mark.parametrize("name,fid", LAZY)
def test_csub(emu, name, fid):
    F = po.FIELDS[name]
    for v in [0, 1, F.p - 1, F.p, F.p + 1, 2 * F.p - 1]:
        assert emu(fid, 11, [v], 1)[0] == v % F.p


@pytest.mark.parametrize("name,fid", LAZY[:2])
def test_beaver_elements_match_reference_sequence(emu, name, fid):
    """mask/recombine per element vs the reference's unfused sequence (pyoracle.beaver_*), Montgomery images."""
    F = po.FIELDS[name]
    rng = random.Random(30 + fid)
    M = F.to_mont
    for it in range(150):
        edge = it < 6
        pick = (lambda: rng.choice([0, 1, F.p - 1])) if edge else (lambda: rng.randrange(F.p))
        key = pick()
        xs, ys = (pick(), pick()), (pick(), pick())
        a, b, c = (pick(), pick()), (pick(), pick()), (pick(), pick())
        dp, ep = pick(), pick()
        d_mine, e_mine = po.beaver_mask(F, [xs], [ys], [a], [b])
        got = emu(fid, 6, [M(xs[0]), M(ys[0]), M(a[0]), M(b[0])], 2)
        assert got == [M(d_mine[0]), M(e_mine[0])]
        d = po.open_add(F, d_mine, [dp])
        e = po.open_add(F, e_mine, [ep])
        for party in (0, 1):
            want = po.beaver_recombine(F, party, key, d, e, [a], [b], [c])[0]
            ins = [M(key), M(d_mine[0]), M(e_mine[0]), M(dp), M(ep), M(a[0]), M(a[1]), M(b[0]), M(b[1]), M(c[0]), M(c[1])]
            out_s, out_m, dd, ee = emu(fid, 7, ins, 4, party)
            assert (out_s, out_m) == (M(want[0]), M(want[1]))
            assert (dd, ee) == (M(d[0]), M(e[0]))


@pytest.mark.parametrize("name,fid", LAZY[:2])
def test_linear_gate_elements(emu, name, fid):
    F = po.FIELDS[name]
    rng = random.Random(40 + fid)
    M = F.to_mont
    for _ in range(100):
        key, s, m, v = (rng.randrange(F.p) for _ in range(4))
        for party in (0, 1):
            w = po.share_add_public(F, (s, m), v, key, party)
            assert emu(fid, 8, [M(key), M(s), M(m), M(v)], 2, party) == [M(w[0]), M(w[1])]
            w = po.share_sub_public(F, (s, m), v, key, party)
            assert emu(fid, 9, [M(key), M(s), M(m), M(v)], 2, party) == [M(w[0]), M(w[1])]
        assert emu(fid, 10, [M(key), M(v), M(m)], 1)[0] == M((key * v - m) % F.p)


@pytest.mark.parametrize("name,fid", FIELDS)
def test_dedicated_square(emu, name, fid):
    """Fp::sqr (512-bit square with 36 multiply-adds + word-serial Montgomery reduction) == canonical a*a/R mod p."""
    F = po.FIELDS[name]
    rng = random.Random(40 + fid)
    hard = [F.p - 1, F.p - 2, (F.p - 1) // 2, int("ffffffff00000000" * 4, 16) % F.p, int("00000000ffffffff" * 4, 16) % F.p,
            int("ffffffff" * 7 + "00000000", 16) % F.p, int("80000000" * 8, 16) % F.p]
    for a in samples(F, rng, 300) + hard:
        assert emu(fid, 13, [a], 1)[0] == a * a * F.rinv % F.p


@pytest.mark.parametrize("name,fid", LAZY)
def test_constant_multiplier_table(emu, name, fid):
    """ctab.hpp builds k[i] = s*2^(32i-160) (i<4) / s*2^(32i-192) (i>=4) mod p; Fp::mul_ctab_lazy(T, a) == s*a/R mod p with
    the result < p + 5p/2^32 for ANY 256-bit a (three reduction steps instead of eight); mul_ctab is canonical."""
    F = po.FIELDS[name]
    rng = random.Random(50 + fid)
    inv2 = pow(2, -1, F.p)
    keys = [0, 1, F.p - 1, F.p - 2, F.r, F.r2, (F.p - 1) // 2] + [rng.randrange(F.p) for _ in range(40)]
    for s in keys:
        tab = emu(fid, 16, [s], 8)
        for i in range(8):
            e = 32 * i - 160 if i < 4 else 32 * i - 192
            want = s * (pow(2, e, F.p) if e >= 0 else pow(inv2, -e, F.p)) % F.p
            assert tab[i] == want
        xs = [0, 1, F.p - 1, (1 << 256) - 1, int("ffffffff00000000" * 4, 16), int("00000000ffffffff" * 4, 16), 3 * F.p % (1 << 256)]
        xs += [rng.randrange(1 << 256) for _ in range(6)] + [rng.randrange(F.p) for _ in range(6)]
        for a in xs:
            r = emu(fid, 14, [s, a], 1)[0]
            assert r % F.p == s * a * F.rinv % F.p
            assert r < F.p + (5 * F.p >> 32) + 1
            assert emu(fid, 15, [s, a], 1)[0] == s * a * F.rinv % F.p


@pytest.mark.parametrize("name,fid", FIELDS)
def test_binary_euclid_inversion(emu, name, fid):
    """Fp::inv_plain (binary extended Euclid), Fp::inv_safegcd (Bernstein-Yang division steps, radix 2^30) and Fp::inv_mont
    (Montgomery image in, Montgomery image out): the inversions at the top of the batch-inversion product tree (scalar.rs:93-100
    -> ark_ff::batch_inversion)."""
    F = po.FIELDS[name]
    rng = random.Random(60 + fid)
    vals = [1, 2, 3, F.p - 1, F.p - 2, (F.p - 1) // 2, (F.p + 1) // 2, 1 << 255 if F.p > 1 << 255 else 1 << 252, (1 << 128) - 1, F.r, F.r2]
    vals += [(1 << k) % F.p for k in range(1, 256, 7)] + [F.p - (1 << k) for k in range(0, 250, 11)] + [(1 << 30) - 1, 1 << 30, (1 << 60) + 1]
    vals += [rng.randrange(1, F.p) for _ in range(400)]
    for a in vals:
        a %= F.p
        if a == 0:
            continue
        assert emu(fid, 19, [a], 1)[0] == pow(a, -1, F.p)
        assert emu(fid, 17, [a], 1)[0] == pow(a, -1, F.p)
        assert emu(fid, 18, [F.to_mont(a)], 1)[0] == F.to_mont(pow(a, -1, F.p))


@pytest.mark.parametrize("name,fid", FIELDS)
def test_karatsuba_product(emu, name, fid):
    """kara512 (one Karatsuba level: three 4 x 4-limb products) is the exact 512-bit product for ANY 256-bit operands, and
    Fp::mul_kara (that product + word-serial Montgomery reduction) equals Fp::mul on canonical inputs."""
    F = po.FIELDS[name]
    rng = random.Random(90 + fid)
    full = (1 << 256) - 1
    halves = [0, 1, (1 << 128) - 1, 1 << 127, (1 << 64) + 1, (1 << 128) - (1 << 32)]
    raw = [lo | (hi << 128) for lo in halves for hi in halves] + [full, full - 1, 1 << 255] + [rng.getrandbits(256) for _ in range(300)]
    for a, b in zip(raw, reversed(raw)):
        lo, hi = emu(fid, 21, [a, b], 2)
        assert lo | (hi << 256) == a * b
    A = samples(F, rng, 300)
    B = list(reversed(samples(F, rng, 300)))
    for a, b in zip(A, B):
        assert emu(fid, 20, [a, b], 1)[0] == a * b * F.rinv % F.p
