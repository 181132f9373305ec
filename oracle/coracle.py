"""ctypes binding of oracle/libark_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY.

Arrays are numpy uint64: a vector of n scalars is shape (n, 4) (Montgomery image, LE limbs);
a vector of n ScalarShares is shape (n, 8) = AoS {share[4], mac[4]} — the reference's
memory image (/root/reference/online-phase/src/algebra/scalar/share.rs:32-37)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libark_oracle.so")

FIELD_IDS = {"bn254_fr": 0, "curve25519_fr": 1, "bn254_fq": 2, "curve25519_fq": 3}


def build() -> str:
    src = os.path.join(_HERE, "ark_oracle.c")
    if (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_field_inv.restype = C.c_uint64
        for name in ("orc_field_modulus", "orc_field_r", "orc_field_r2"):
            getattr(_lib, name).restype = C.POINTER(C.c_uint64)
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def field_constants(f: int):
    L = lib()
    g = lambda fn: [int(fn(f)[i]) for i in range(4)]
    return dict(p=g(L.orc_field_modulus), r=g(L.orc_field_r), r2=g(L.orc_field_r2), inv=int(L.orc_field_inv(f)))


def ints_to_limbs(vals) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            out[i, j] = (int(v) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(a: np.ndarray):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(a[i, j]) << (64 * j) for j in range(4)) for i in range(a.shape[0])]


def to_mont(f: int, plain: np.ndarray) -> np.ndarray:
    plain = _u64(plain)
    out = np.empty_like(plain)
    lib().orc_to_mont(f, C.c_size_t(plain.size // 4), _p(out), _p(plain))
    return out


def from_mont(f: int, mont: np.ndarray) -> np.ndarray:
    mont = _u64(mont)
    out = np.empty_like(mont)
    lib().orc_from_mont(f, C.c_size_t(mont.size // 4), _p(out), _p(mont))
    return out


def synth(f: int, seed: int, first: int, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_synth(f, C.c_uint64(seed), C.c_uint64(first), C.c_size_t(n), _p(out))
    return out


def _binary(name, f, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    getattr(lib(), name)(f, C.c_size_t(a.shape[0]), _p(out), _p(a), _p(b))
    return out


def scalar_add(f, a, b): return _binary("orc_scalar_batch_add", f, a, b)
def scalar_sub(f, a, b): return _binary("orc_scalar_batch_sub", f, a, b)
def scalar_mul(f, a, b): return _binary("orc_scalar_batch_mul", f, a, b)
def batch_add(f, a, b): return _binary("orc_batch_add", f, a, b)
def batch_sub(f, a, b): return _binary("orc_batch_sub", f, a, b)
def batch_mul_public(f, a, s): return _binary("orc_batch_mul_public", f, a, s)


def batch_neg(f, a):
    a = _u64(a)
    out = np.empty_like(a)
    lib().orc_batch_neg(f, C.c_size_t(a.shape[0]), _p(out), _p(a))
    return out


def batch_add_public(f, party, key, a, v, sub=False):
    a, v, key = _u64(a), _u64(v), _u64(key)
    out = np.empty_like(a)
    fn = lib().orc_batch_sub_public if sub else lib().orc_batch_add_public
    fn(f, party, _p(key), C.c_size_t(a.shape[0]), _p(out), _p(a), _p(v))
    return out


def beaver_mask(f, x, y, a, b):
    x, y, a, b = map(_u64, (x, y, a, b))
    n = x.shape[0]
    d = np.empty((n, 4), dtype=np.uint64)
    e = np.empty((n, 4), dtype=np.uint64)
    scratch = np.empty((2 * n, 8), dtype=np.uint64)
    lib().orc_beaver_mask(f, C.c_size_t(n), _p(x), _p(y), _p(a), _p(b), _p(d), _p(e), _p(scratch))
    return d, e


def beaver_recombine(f, party, key, d, e, a, b, c):
    key, d, e, a, b, c = map(_u64, (key, d, e, a, b, c))
    n = d.shape[0]
    out = np.empty((n, 8), dtype=np.uint64)
    scratch = np.empty(n * 4 + 5 * n * 8, dtype=np.uint64)
    lib().orc_beaver_recombine(f, party, _p(key), C.c_size_t(n), _p(d), _p(e), _p(a), _p(b), _p(c), _p(out), _p(scratch))
    return out


def mac_check(f, key, opened, shares):
    key, opened, shares = map(_u64, (key, opened, shares))
    out = np.empty_like(opened)
    lib().orc_mac_check(f, _p(key), C.c_size_t(opened.shape[0]), _p(opened), _p(shares), _p(out))
    return out


def share_sum(f, shares):
    shares = _u64(shares)
    out = np.empty(8, dtype=np.uint64)
    lib().orc_share_sum(f, C.c_size_t(shares.shape[0]), _p(shares), _p(out))
    return out


def two_party_batch_mul(f, threads, keys, x, y, a, b, c, want_open=True):
    """keys/x/y/a/b/c: pairs (party0, party1) of arrays.  Returns (out0, out1, d_open, e_open)."""
    n = x[0].shape[0]
    arrs = [_u64(v) for pair in (x, y, a, b, c) for v in pair]  # x0 x1 y0 y1 ...
    x0, x1, y0, y1, a0, a1, b0, b1, c0, c1 = arrs
    k0, k1 = _u64(keys[0]), _u64(keys[1])
    out0 = np.empty((n, 8), dtype=np.uint64)
    out1 = np.empty((n, 8), dtype=np.uint64)
    d = np.empty((n, 4), dtype=np.uint64) if want_open else None
    e = np.empty((n, 4), dtype=np.uint64) if want_open else None
    rc = lib().orc_two_party_batch_mul(
        f, C.c_size_t(n), threads, _p(k0), _p(k1), _p(x0), _p(y0), _p(a0), _p(b0), _p(c0),
        _p(x1), _p(y1), _p(a1), _p(b1), _p(c1), _p(out0), _p(out1),
        _p(d) if want_open else None, _p(e) if want_open else None)
    assert rc == 0
    return out0, out1, d, e


# ---------------------------------------------------------------------------------------------
# Curve groups (points as uint64 arrays in the reference's AoS projective image: (n, 12) for BN254 G1,
# (n, 16) for Curve25519 Edwards; PointShares are (n, 2*words) = {share, mac})
# ---------------------------------------------------------------------------------------------
CURVE_IDS = {"bn254_g1": 0, "curve25519_edwards": 1}
CURVE_FQ = {0: 2, 1: 3}  # base-field ids in FIELD_IDS
CURVE_FR = {0: 0, 1: 1}


def point_words(cv: int) -> int:
    return int(lib().orc_point_words(cv))


def pt_generator(cv: int) -> np.ndarray:
    out = np.zeros(point_words(cv), dtype=np.uint64)
    lib().orc_pt_generator(cv, _p(out))
    return out


def pt_normalize(cv: int, pts: np.ndarray) -> np.ndarray:
    """(n, words) projective -> (n, 8) affine (x, y) Montgomery limbs; BN254 identity -> zeros."""
    pts = _u64(pts).reshape(-1, point_words(cv))
    out = np.empty((pts.shape[0], 8), dtype=np.uint64)
    lib().orc_pt_normalize(cv, C.c_size_t(pts.shape[0]), _p(pts), _p(out))
    return out


def pt_mul(cv: int, scalars: np.ndarray, pts: np.ndarray) -> np.ndarray:
    scalars, pts = _u64(scalars), _u64(pts)
    out = np.empty_like(pts)
    lib().orc_pt_mul(cv, C.c_size_t(scalars.shape[0]), _p(scalars), _p(pts), _p(out))
    return out


def pt_mul_generator(cv: int, scalars: np.ndarray) -> np.ndarray:
    scalars = _u64(scalars)
    out = np.empty((scalars.shape[0], point_words(cv)), dtype=np.uint64)
    lib().orc_pt_mul_generator(cv, C.c_size_t(scalars.shape[0]), _p(scalars), _p(out))
    return out


def pt_add(cv: int, a: np.ndarray, b: np.ndarray, sub: bool = False) -> np.ndarray:
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    lib().orc_pt_add(cv, C.c_size_t(a.shape[0]), _p(a), _p(b), _p(out), int(sub))
    return out


def pt_share_add_public(cv: int, party: int, key: np.ndarray, a_ps: np.ndarray, pub: np.ndarray, sub: bool = False) -> np.ndarray:
    a_ps, pub, key = _u64(a_ps), _u64(pub), _u64(key)
    out = np.empty_like(a_ps)
    lib().orc_pt_share_add_public(cv, party, _p(key), C.c_size_t(pub.shape[0]), _p(a_ps), _p(pub), _p(out), int(sub))
    return out


def two_party_point_mul(cv, threads, keys, x, P, a, b, c, want_open=True):
    """x/a/b/c: pairs of (n,8) AoS ScalarShare arrays; P: pair of (n, 2*words) PointShare arrays.
    Returns (out0, out1, d_open, E_open) with out* PointShare arrays (projective)."""
    n = x[0].shape[0]
    w = point_words(cv)
    x0, x1, P0, P1, a0, a1, b0, b1, c0, c1 = [_u64(v) for pair in (x, P, a, b, c) for v in pair]
    k0, k1 = _u64(keys[0]), _u64(keys[1])
    out0 = np.empty((n, 2 * w), dtype=np.uint64)
    out1 = np.empty((n, 2 * w), dtype=np.uint64)
    d = np.empty((n, 4), dtype=np.uint64) if want_open else None
    E = np.empty((n, w), dtype=np.uint64) if want_open else None
    rc = lib().orc_two_party_point_mul(cv, C.c_size_t(n), threads, _p(k0), _p(k1), _p(x0), _p(P0), _p(a0), _p(b0), _p(c0),
                                       _p(x1), _p(P1), _p(a1), _p(b1), _p(c1), _p(out0), _p(out1),
                                       _p(d) if want_open else None, _p(E) if want_open else None)
    assert rc == 0
    return out0, out1, d, E


# ---------------------------------------------------------------------------------------------
# Batch inversion and FFT (BN254 Fr)
# ---------------------------------------------------------------------------------------------
def batch_inverse(f: int, a: np.ndarray) -> np.ndarray:
    a = _u64(a)
    out = np.empty_like(a)
    lib().orc_batch_inverse(f, C.c_size_t(a.shape[0]), _p(out), _p(a))
    return out


def fft(f: int, a: np.ndarray, inverse: bool = False) -> np.ndarray:
    a = _u64(a)
    n = a.shape[0]
    assert n & (n - 1) == 0 and n > 0
    out = np.empty_like(a)
    rc = lib().orc_fft(f, n.bit_length() - 1, int(inverse), _p(a), _p(out))
    assert rc == 0, "orc_fft: unsupported field or size"
    return out
