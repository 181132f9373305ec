"""Binary wire format for the batch payloads of the path (second half).

The reference frames every outbound message as `u64 LE length || serde_json(NetworkOutbound { result_id, payload })`
(/root/reference/online-phase/src/network/quic.rs:292-310); inside the JSON a scalar is the byte ARRAY of arkworks'
`serialize_uncompressed` (algebra/scalar/scalar.rs:187-202: the canonical integer, 32 bytes little-endian) spelled out as
decimal numbers — ~3.6 characters per byte, and a serde pass over every element.  With the arithmetic on the GPU that encoder
is the bottleneck, so batches travel in a fixed binary layout that a device buffer can be copied into and out of:

    frame  = u64 LE body length || body                                    (the length prefix quic.rs already writes)
    body   = u64 LE result_id || u8 tag || u8 field/curve id || u16 0 || u32 LE count || payload
    tag    = NetworkPayload's variant index (network.rs:45-60): 0 Bytes, 1 Scalar, 2 ScalarBatch, 3 ScalarShare,
             4 Point, 5 PointBatch, 6 PointShare
    scalar = arkworks `serialize_uncompressed`: canonical integer, 32 bytes LE          (what Scalar's Serialize impl emits)
    share  = scalar(share) || scalar(mac)
    point  = affine x || y, each the canonical integer, 32 bytes LE; the identity is (0, 0) on BN254 G1 and (0, 1) on
             Curve25519 — arkworks' UNCOMPRESSED affine encoding without its flag bits.  The reference sends the compressed
             form (curve.rs:103-115); decompression needs a square root per point, so the device format keeps both coordinates.

Encode = one kernel out of Montgomery form (`arkmpc_fr_from_mont` / `arkmpc_pt_normalize`) + one device->host copy; decode = one
host->device copy + `arkmpc_fr_validate` (reject non-canonical residues, as arkworks' deserialiser does) + `arkmpc_fr_to_mont`
(/ `arkmpc_pt_from_affine` + `arkmpc_pt_validate`: on the curve and in the prime-order subgroup)."""
from __future__ import annotations

import struct
from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from . import _native as nat

TAG_BYTES, TAG_SCALAR, TAG_SCALAR_BATCH, TAG_SCALAR_SHARE, TAG_POINT, TAG_POINT_BATCH, TAG_POINT_SHARE = range(7)
_HEADER = struct.Struct("<QBBHI")  # result_id, tag, field / curve id, reserved, count


class WireError(ValueError):
    """SerializationError of the reference (network.rs): malformed frame or an invalid field element / point."""


@dataclass
class Frame:
    result_id: int
    tag: int
    ident: int
    count: int
    payload: memoryview


def frame(result_id: int, tag: int, ident: int, count: int, payload: bytes) -> bytes:
    body = _HEADER.pack(result_id, tag, ident, 0, count) + payload
    return struct.pack("<Q", len(body)) + body


def parse(buf: bytes) -> Frame:
    mv = memoryview(buf)
    if len(mv) < 8 + _HEADER.size:
        raise WireError("short frame")
    (length,) = struct.unpack_from("<Q", mv, 0)
    if length != len(mv) - 8:
        raise WireError(f"frame length {length} does not match the {len(mv) - 8} bytes that follow")
    result_id, tag, ident, reserved, count = _HEADER.unpack_from(mv, 8)
    if reserved != 0 or tag > TAG_POINT_SHARE:
        raise WireError("unknown tag or non-zero reserved bits")
    return Frame(result_id, tag, ident, count, mv[8 + _HEADER.size:])


# -- scalars -----------------------------------------------------------------------------------------------------------
def encode_scalar_batch(E, result_id: int, plane: torch.Tensor) -> bytes:
    """NetworkPayload::ScalarBatch of a device plane (Montgomery images) -> frame."""
    n = plane.shape[0]
    plain = E.from_mont(plane) if n else plane
    return frame(result_id, TAG_SCALAR_BATCH, E.field, n, E.download(plain).tobytes())


def decode_scalar_batch(E, buf: bytes) -> Tuple[int, torch.Tensor]:
    f = parse(buf)
    if f.tag != TAG_SCALAR_BATCH or f.ident != E.field:
        raise WireError("not a ScalarBatch of this field")
    if len(f.payload) != 32 * f.count:
        raise WireError("payload size does not match the element count")
    plain = E.upload(np.frombuffer(f.payload, dtype=np.uint64).reshape(f.count, 4))
    if not E.validate(plain):
        raise WireError("non-canonical field element (arkworks: InvalidData)")
    return f.result_id, (E.to_mont(plain) if f.count else plain)


def encode_share_batch(E, result_id: int, planes) -> bytes:
    """A vector of ScalarShares as interleaved (share, mac) canonical integers (n x NetworkPayload::ScalarShare)."""
    n = planes[0].shape[0]
    aos = E.share_zip((E.from_mont(planes[0]), E.from_mont(planes[1]))) if n else torch.empty((0, 8), dtype=torch.int64, device=E.tdev)
    return frame(result_id, TAG_SCALAR_SHARE, E.field, n, E.download(aos).tobytes())


def decode_share_batch(E, buf: bytes):
    f = parse(buf)
    if f.tag != TAG_SCALAR_SHARE or f.ident != E.field:
        raise WireError("not a ScalarShare batch of this field")
    if len(f.payload) != 64 * f.count:
        raise WireError("payload size does not match the element count")
    if f.count == 0:
        return f.result_id, (E.empty(0), E.empty(0))
    aos = torch.from_numpy(np.frombuffer(f.payload, dtype=np.int64).reshape(f.count, 8).copy()).to(E.tdev)
    s, m = E.share_unzip(aos)
    if not (E.validate(s) and E.validate(m)):
        raise WireError("non-canonical field element (arkworks: InvalidData)")
    return f.result_id, (E.to_mont(s), E.to_mont(m))


# -- points ------------------------------------------------------------------------------------------------------------
def _coords_from_mont(E, xy: torch.Tensor) -> torch.Tensor:
    """(n, 8) affine Montgomery images of BASE-field elements -> canonical integers.  The base field is not one of the engine's
    scalar fields, so the conversion runs on the host for the few places that need it (wire encode / decode of points)."""
    from . import fields as fl

    q = fl.BASE_MODULUS[E.curve]
    rinv = pow(1 << 256, -1, q)
    a = E.download(xy).reshape(-1, 4)
    out = np.empty_like(a)
    for i in range(a.shape[0]):
        v = sum(int(a[i, j]) << (64 * j) for j in range(4)) * rinv % q
        out[i] = [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    return out.reshape(-1, 8)


def encode_point_batch(E, result_id: int, pts: torch.Tensor) -> bytes:
    """NetworkPayload::PointBatch of device points (projective images) -> frame with canonical affine coordinates."""
    n = pts.shape[0]
    E._curve()
    payload = _coords_from_mont(E, E.pt_normalize(pts)).tobytes() if n else b""
    return frame(result_id, TAG_POINT_BATCH, E.curve, n, payload)


def decode_point_batch(E, buf: bytes) -> Tuple[int, torch.Tensor]:
    from . import fields as fl

    f = parse(buf)
    E._curve()
    if f.tag != TAG_POINT_BATCH or f.ident != E.curve:
        raise WireError("not a PointBatch of this curve")
    if len(f.payload) != 64 * f.count:
        raise WireError("payload size does not match the point count")
    if f.count == 0:
        return f.result_id, E.empty_points(0)
    q = fl.BASE_MODULUS[E.curve]
    a = np.frombuffer(f.payload, dtype=np.uint64).reshape(-1, 4)
    mont = np.empty_like(a)
    for i in range(a.shape[0]):
        v = sum(int(a[i, j]) << (64 * j) for j in range(4))
        if v >= q:
            raise WireError("non-canonical coordinate (arkworks: InvalidData)")
        v = v * (1 << 256) % q
        mont[i] = [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    pts = E.pt_from_affine(E.upload_points(mont.reshape(-1, 8)))
    if not E.pt_validate(pts):
        raise WireError("point off the curve or outside the prime-order subgroup (arkworks: InvalidData)")
    return f.result_id, pts


def json_size_estimate(n_scalars: int) -> int:
    """Bytes serde_json needs for a ScalarBatch of n uniformly random scalars: every byte becomes a decimal number plus a comma
    (average 3.57 characters), plus brackets — what the reference's QUIC transport sends today (quic.rs:303)."""
    return int(n_scalars * (32 * 3.57 + 3)) + 40


__all__ = ["WireError", "Frame", "frame", "parse", "encode_scalar_batch", "decode_scalar_batch", "encode_share_batch", "decode_share_batch",
           "encode_point_batch", "decode_point_batch", "json_size_estimate", "nat"]
