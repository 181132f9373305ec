"""Pins the C restatement of batch inversion and the radix-2 FFT (oracle/ark_oracle.c) against the definition in exact
big-int arithmetic (oracle/pyoracle.py) and against the published BN254 Fr root of unity."""
import random

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyoracle as po

F = po.BN254_FR


def test_root_of_unity_known_answer():
    # the 2^28-th root of unity of BN254 Fr used by arkworks, circom/snarkjs and gnark alike
    assert po.root_of_unity(F, 1 << 28) == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    w = po.root_of_unity(F, 1 << 10)
    assert pow(w, 1 << 10, F.p) == 1 and pow(w, 1 << 9, F.p) == F.p - 1


@pytest.mark.parametrize("n", [1, 2, 4, 8, 64, 256])
def test_c_fft_matches_definition(n):
    rng = random.Random(n)
    xs = [rng.randrange(F.p) for _ in range(n)]
    if n >= 4:
        xs[0], xs[1] = 0, F.p - 1
    a = co.to_mont(0, co.ints_to_limbs(xs))
    fwd = co.fft(0, a)
    assert [F.from_mont(v) for v in co.limbs_to_ints(fwd)] == po.naive_dft(F, xs)
    inv = co.fft(0, a, inverse=True)
    assert [F.from_mont(v) for v in co.limbs_to_ints(inv)] == po.naive_dft(F, xs, inverse=True)
    assert np.array_equal(co.fft(0, fwd, inverse=True), a)


@pytest.mark.parametrize("fid,name", [(0, "bn254_fr"), (1, "curve25519_fr")])
def test_c_batch_inverse(fid, name):
    Fd = po.FIELDS[name]
    rng = random.Random(3 + fid)
    xs = [0, 1, Fd.p - 1, 2] + [rng.randrange(Fd.p) for _ in range(60)] + [0]
    a = co.to_mont(fid, co.ints_to_limbs(xs))
    got = [Fd.from_mont(v) for v in co.limbs_to_ints(co.batch_inverse(fid, a))]
    assert got == po.batch_inverse(Fd, xs)
