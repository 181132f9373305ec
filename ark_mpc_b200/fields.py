"""Field and curve parameters the host side needs (moduli, Montgomery images of small constants).

Host-side helpers only: the arithmetic on batches runs in the CUDA library.  Values cross the C ABI as the
reference's memory image — canonical Montgomery residues, R = 2^256, four little-endian u64 limbs
(`Scalar<C>`, /root/reference/online-phase/src/algebra/scalar/scalar.rs:46)."""
from __future__ import annotations

from typing import Iterable, List

import numpy as np

R = 1 << 256
MASK64 = (1 << 64) - 1

MODULUS = {
    "bn254_fr": 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
    "curve25519_fr": (1 << 252) + 27742317777372353535851937790883648493,
    "bn254_fq": 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47,
    "curve25519_fq": (1 << 255) - 19,
}
FIELD_OF_CURVE = {"bn254_g1": "bn254_fr", "curve25519_edwards": "curve25519_fr"}
BASE_FIELD_OF_CURVE = {"bn254_g1": "bn254_fq", "curve25519_edwards": "curve25519_fq"}
CURVE_OF_FIELD = {v: k for k, v in FIELD_OF_CURVE.items()}
POINT_WORDS = {"bn254_g1": 12, "curve25519_edwards": 16}
# base-field modulus by native curve id (arkmpc_curve): 0 BN254 G1, 1 Curve25519 Edwards
BASE_MODULUS = {0: MODULUS["bn254_fq"], 1: MODULUS["curve25519_fq"]}


def int_to_limbs(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & MASK64 for i in range(4)], dtype=np.uint64)


def limbs_to_int(l) -> int:
    return sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(l, dtype=np.uint64).reshape(4)))


def mont_limbs(field: str, v: int) -> np.ndarray:
    """Montgomery image of the integer v as 4 LE u64 limbs."""
    p = MODULUS[field]
    return int_to_limbs((v % p) * R % p)


def mont_limbs_batch(field: str, vals: Iterable[int]) -> np.ndarray:
    vals = list(vals)
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = mont_limbs(field, v)
    return out


def from_mont_limbs(field: str, limbs) -> int:
    p = MODULUS[field]
    return limbs_to_int(limbs) * pow(R, -1, p) % p


def from_mont_batch(field: str, arr: np.ndarray) -> List[int]:
    p = MODULUS[field]
    rinv = pow(R, -1, p)
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(a[i, j]) << (64 * j) for j in range(4)) * rinv % p for i in range(a.shape[0])]
