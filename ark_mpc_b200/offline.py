"""Offline-phase batch algebra on the device (second half): `ValueMacBatch` and the two sub-protocols of
LowGear that are pure batch arithmetic over it — the Beaver multiplication and the SPDZ sacrifice — mirrored from
/root/reference/offline-phase/src:

  ValueMacBatch  Add / Sub / Mul<Scalar> / Mul<&[Scalar]>, values(), macs(), from_parts, split_at   structs.rs:185-382
  open_batch, open_and_check_macs (random linear combination of values and MACs, one MAC check)     lowgear/mac_check.rs:14-48
  beaver_mul      [xy] = de + d[b] + e[a] + [c], public de added by party 0 only                    lowgear/multiplication.rs:13-70
  sacrifice       rho = open(r b - b'), tau = open(r c - c' - rho a) must be all zero                lowgear/triplets.rs:118-150

Everything runs on the same scalar-field kernels as the online path (`arkmpc_fr_share_*`, `arkmpc_fr_mul`, `arkmpc_fr_sum`, ...);
a batch is two device planes.  What is NOT here is the BGV side of LowGear (ciphertext arithmetic lives in MP-SPDZ).  The
commit-reveal of single field elements stays on the host, as the hash commitments of the online phase do."""
from __future__ import annotations

import secrets
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import fields as fl
from .engine import Planes
from .fabric import MpcError, MpcFabric, _commit


class LowGearError(MpcError):
    pass


class InvalidMac(LowGearError):  # LowGearError::InvalidMac
    pass


class SacrificeError(LowGearError):  # LowGearError::SacrificeError
    pass


class ValueMacBatch:
    """structs.rs:185-382 — a vector of (value share, mac share) pairs; here: a share plane and a mac plane on the device."""

    def __init__(self, fabric: MpcFabric, share: torch.Tensor, mac: torch.Tensor):
        assert share.shape == mac.shape
        self.fabric, self.share, self.mac = fabric, share, mac

    @staticmethod
    def from_parts(fabric: MpcFabric, values: torch.Tensor, macs: torch.Tensor) -> "ValueMacBatch":  # :274-283
        return ValueMacBatch(fabric, values, macs)

    def __len__(self) -> int:
        return self.share.shape[0]

    def planes(self) -> Planes:
        return self.share, self.mac

    def values(self) -> torch.Tensor:  # :241-243
        return self.share

    def macs(self) -> torch.Tensor:  # :246-248
        return self.mac

    def split_at(self, i: int) -> Tuple["ValueMacBatch", "ValueMacBatch"]:  # :261-264
        f = self.fabric
        return (ValueMacBatch(f, self.share[:i].contiguous(), self.mac[:i].contiguous()),
                ValueMacBatch(f, self.share[i:].contiguous(), self.mac[i:].contiguous()))

    def _same(self, other: "ValueMacBatch") -> None:
        assert len(self) == len(other), "batch lengths differ"  # the reference's assert_eq!

    def __add__(self, other: "ValueMacBatch") -> "ValueMacBatch":  # :319-333
        self._same(other)
        return ValueMacBatch(self.fabric, *self.fabric.engine.share_add(self.planes(), other.planes()))

    def __sub__(self, other: "ValueMacBatch") -> "ValueMacBatch":  # :335-349
        self._same(other)
        return ValueMacBatch(self.fabric, *self.fabric.engine.share_sub(self.planes(), other.planes()))

    def mul_scalar(self, s: np.ndarray) -> "ValueMacBatch":  # Mul<Scalar<C>> :351-363 (s: 4 Montgomery limbs)
        E = self.fabric.engine
        return ValueMacBatch(self.fabric, E.scale(self.share, s), E.scale(self.mac, s))

    def mul_elementwise(self, v: torch.Tensor) -> "ValueMacBatch":  # Mul<&[Scalar<C>]> :366-378
        assert v.shape[0] == len(self)
        return ValueMacBatch(self.fabric, *self.fabric.engine.share_mul_public(self.planes(), v))

    def add_public_value(self, public: torch.Tensor) -> "ValueMacBatch":  # multiplication.rs:56-70
        f = self.fabric
        return ValueMacBatch(f, *f.engine.share_add_public(f.party_id(), f.mac_key(), self.planes(), public))


# -- sub-protocols -----------------------------------------------------------------------------------------------------
def _exchange_scalar(f: MpcFabric, v: int) -> int:
    E = f.engine
    t = torch.from_numpy(fl.int_to_limbs(v).view(np.int64).reshape(1, 4).copy()).to(E.tdev)
    return fl.limbs_to_int(E.download(f.exchange_tensor(t))[0])


def commit_reveal_single(f: MpcFabric, value: int) -> int:
    """lowgear/commit_reveal.rs: commit to one field element, exchange commitments, then open and verify the peer's."""
    blinder = secrets.randbelow(fl.MODULUS[f.field])
    mine = _commit([value.to_bytes(32, "big")], blinder.to_bytes(32, "big"), f.field)
    peer_comm = _exchange_scalar(f, mine)
    peer_value = _exchange_scalar(f, value)
    peer_blinder = _exchange_scalar(f, blinder)
    if _commit([peer_value.to_bytes(32, "big")], peer_blinder.to_bytes(32, "big"), f.field) != peer_comm:
        raise LowGearError("peer's commitment does not open")
    return peer_value


def get_shared_randomness_vec(f: MpcFabric, n: int) -> torch.Tensor:
    """lowgear/shared_random.rs: both parties contribute a seed through commit-reveal; the vector is expanded from the sum of the
    seeds with the library's counter-based generator (identical on both sides)."""
    p = fl.MODULUS[f.field]
    mine = secrets.randbelow(1 << 62)
    seed = (mine + commit_reveal_single(f, mine)) % (1 << 63)
    _ = p
    return f.engine.random(seed, 0, n)


def open_batch(f: MpcFabric, values: torch.Tensor) -> torch.Tensor:
    E = f.engine
    peer = f.exchange_tensor(values)
    if not E.validate(peer):
        raise LowGearError("peer sent a non-canonical field element")
    return E.add(values, peer)


def open_and_check_macs(f: MpcFabric, x: ValueMacBatch) -> torch.Tensor:
    """mac_check.rs:14-42: open, then check ONE MAC on a random linear combination of the opened values and of the MAC shares."""
    E = f.engine
    n = len(x)
    if n == 0:
        return E.empty(0)
    recovered = open_batch(f, x.values())
    r = get_shared_randomness_vec(f, n)
    combined_value = E.sum(E.mul(recovered, r))       # linear_combination, :45-48
    combined_mac = E.sum(E.mul(x.macs(), r))
    # mac_check = mac - mac_share * x  (:34); commit-reveal it; the two parties' values must cancel
    mine_t = E.neg(E.mac_check(f.mac_key(), combined_value, combined_mac))
    mine = fl.from_mont_limbs(f.field, E.download(mine_t)[0])
    theirs = commit_reveal_single(f, mine)
    if (mine + theirs) % fl.MODULUS[f.field] != 0:
        raise InvalidMac("MAC check failed on an opened batch")
    return recovered


def beaver_mul(f: MpcFabric, lhs: ValueMacBatch, rhs: ValueMacBatch, triples: Tuple[ValueMacBatch, ValueMacBatch, ValueMacBatch]) -> ValueMacBatch:
    """multiplication.rs:13-39 with the triples passed in (the reference pops them from `self.triples`)."""
    E = f.engine
    a, b, c = triples
    assert len(lhs) == len(rhs) == len(a), "Batch sizes must match"
    d = open_and_check_macs(f, lhs - a)
    e = open_and_check_macs(f, rhs - b)
    de = E.mul(d, e)
    db = b.mul_elementwise(d)
    ea = a.mul_elementwise(e)
    return ((db + ea) + c).add_public_value(de)


def sacrifice(f: MpcFabric, a: ValueMacBatch, b: ValueMacBatch, c: ValueMacBatch, b_prime: ValueMacBatch, c_prime: ValueMacBatch) -> None:
    """triplets.rs:118-150: raises SacrificeError unless c = a b and c' = a b' hold for every triple."""
    E = f.engine
    r = E.download(get_shared_randomness_vec(f, 1))[0].copy()
    rho = open_and_check_macs(f, b.mul_scalar(r) - b_prime)
    rho_a = a.mul_elementwise(rho)
    tau = open_and_check_macs(f, (c.mul_scalar(r) - c_prime) - rho_a)
    zero = torch.zeros_like(tau)
    if not torch.equal(tau, zero):
        raise SacrificeError("sacrifice: tau is not all zero")
