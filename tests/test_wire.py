"""Binary wire format of the batch payloads (ark_mpc_b200/wire.py; reference framing network/quic.rs:292-310, payload
network.rs:45-60, scalar encoding scalar.rs:187-202).  CPU part: framing; GPU part: device encode / decode round trips, the bytes
against the definition (arkworks `serialize_uncompressed` = canonical integer, 32 bytes little-endian) and rejection of
non-canonical or off-curve input."""
import struct

import numpy as np
import pytest

from ark_mpc_b200 import fields as fl
from ark_mpc_b200 import wire


def test_frame_layout_and_parse_errors():
    body = wire.frame(0x1122334455667788, wire.TAG_SCALAR_BATCH, 1, 2, b"\x01" * 64)
    (length,) = struct.unpack_from("<Q", body, 0)
    assert length == len(body) - 8 == 16 + 64
    f = wire.parse(body)
    assert (f.result_id, f.tag, f.ident, f.count, bytes(f.payload)) == (0x1122334455667788, 2, 1, 2, b"\x01" * 64)
    with pytest.raises(wire.WireError):
        wire.parse(body[:-1])            # truncated
    with pytest.raises(wire.WireError):
        wire.parse(body[:10])            # shorter than a header
    bad = bytearray(body)
    bad[8 + 8] = 9                       # unknown tag
    with pytest.raises(wire.WireError):
        wire.parse(bytes(bad))
    assert wire.json_size_estimate(1 << 20) > 3.5 * 32 * (1 << 20)  # what serde_json would send for the same batch


@pytest.mark.gpu
@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
def test_scalar_batch_round_trip_and_bytes(field):
    from ark_mpc_b200.engine import Engine

    E = Engine(0, field)
    p = fl.MODULUS[field]
    vals = [0, 1, p - 1, 2**128, 123456789] + [pow(7, k, p) for k in range(1, 60)]
    plane = E.upload(fl.mont_limbs_batch(field, vals))
    buf = wire.encode_scalar_batch(E, 42, plane)
    f = wire.parse(buf)
    assert f.count == len(vals) and f.tag == wire.TAG_SCALAR_BATCH
    # the payload is the canonical integer, little-endian: what arkworks' serialize_uncompressed writes for Fp256
    assert bytes(f.payload) == b"".join(v.to_bytes(32, "little") for v in vals)
    rid, back = wire.decode_scalar_batch(E, buf)
    assert rid == 42 and np.array_equal(E.download(back), E.download(plane))
    # a residue >= p is rejected like arkworks' deserialiser rejects it
    bad = bytearray(buf)
    bad[8 + 16:8 + 16 + 32] = p.to_bytes(32, "little")
    with pytest.raises(wire.WireError):
        wire.decode_scalar_batch(E, bytes(bad))
    with pytest.raises(wire.WireError):
        wire.decode_scalar_batch(E, buf[:-32] + b"")  # length prefix no longer matches
    rid, empty = wire.decode_scalar_batch(E, wire.encode_scalar_batch(E, 7, E.empty(0)))
    assert rid == 7 and empty.shape[0] == 0
    # shares
    mac = E.upload(fl.mont_limbs_batch(field, list(reversed(vals))))
    sbuf = wire.encode_share_batch(E, 43, (plane, mac))
    rid, (s, m) = wire.decode_share_batch(E, sbuf)
    assert rid == 43 and np.array_equal(E.download(s), E.download(plane)) and np.array_equal(E.download(m), E.download(mac))
    E.close()


@pytest.mark.gpu
@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
def test_point_batch_round_trip(field):
    from ark_mpc_b200.engine import Engine
    from oracle import pyoracle as po

    E = Engine(0, field)
    n = 50
    scal = E.random(3, 0, n)
    scal[0] = 0  # the identity is part of the batch
    pts = E.pt_mul_generator_public(scal)
    buf = wire.encode_point_batch(E, 9, pts)
    f = wire.parse(buf)
    assert f.count == n and f.tag == wire.TAG_POINT_BATCH and len(f.payload) == 64 * n
    # payload = canonical affine coordinates of s_i * G, checked against the big-int oracle
    Cv = po.BN254_G1 if field == "bn254_fr" else po.CURVE25519_EDWARDS
    ks = fl.from_mont_batch(field, E.download(scal))
    for i in (0, 1, n - 1):
        P = Cv.mul(Cv.generator, ks[i])
        want = (0, 0) if P is None else P
        x = int.from_bytes(bytes(f.payload[64 * i:64 * i + 32]), "little")
        y = int.from_bytes(bytes(f.payload[64 * i + 32:64 * i + 64]), "little")
        assert (x, y) == want
    rid, back = wire.decode_point_batch(E, buf)
    assert rid == 9 and np.array_equal(E.download(E.pt_normalize(back)), E.download(E.pt_normalize(pts)))
    # an off-curve point is rejected
    bad = bytearray(buf)
    bad[8 + 16 + 64] ^= 1
    with pytest.raises(wire.WireError):
        wire.decode_point_batch(E, bytes(bad))
    E.close()
