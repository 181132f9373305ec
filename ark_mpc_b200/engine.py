"""Device engine: one `arkmpc_ctx` per (process, device, party) plus tensor-shaped wrappers of the C ABI.

PyTorch is plumbing only here: it owns device memory (int64 tensors used as raw 4xu64 limb storage)
and the CUDA stream.  Every arithmetic call goes through libarkmpc_b200 (include/arkmpc_b200.h).

Vectors: a batch of n scalars is a contiguous int64 CUDA tensor of shape (n, 4) — the Montgomery
image the reference keeps in memory (`Scalar<C>`, online-phase/src/algebra/scalar/scalar.rs:46).
A batch of n `ScalarShare`s is a pair of such planes (share, mac).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat

Planes = Tuple[torch.Tensor, torch.Tensor]


def _limbs_of_int(v: int) -> np.ndarray:
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class Engine:
    """Owns one native context bound to `device`.  Every op is enqueued on the stream that is torch's CURRENT stream for the
    device at the time of the call (the native context is re-pointed when it changed), so torch-side allocations, events and
    `record_stream` made around a call always refer to the stream the kernels run on."""

    def __init__(self, device: int = 0, field: str = "bn254_fr"):
        if not torch.cuda.is_available():
            raise nat.ArkMpcError(nat.ERR_NO_DEVICE, "Engine", "no CUDA device: the gate engine has no CPU fallback")
        self.lib = nat.load()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        self.field_name = field
        self.field = nat.FIELD_IDS[field]
        h = C.c_void_p()
        nat.check(self.lib.arkmpc_ctx_create(self.device, C.byref(h)), "arkmpc_ctx_create")
        self.ctx = h
        self.bind_current_stream()

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "ctx", None):
            self.lib.arkmpc_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def bind_current_stream(self) -> None:
        s = torch.cuda.current_stream(self.tdev).cuda_stream
        if s != getattr(self, "_bound_stream", -1):
            self.lib.arkmpc_ctx_set_stream(self.ctx, C.c_void_p(s))
            self._bound_stream = s

    def hint_independent(self) -> None:
        """The next Beaver kernel launched through this engine does not depend on the kernel launched just before it
        (arkmpc_ctx_hint_independent): its ramp-up overlaps the predecessor's drain."""
        self.lib.arkmpc_ctx_hint_independent(self.ctx)

    def sync(self) -> None:
        self.bind_current_stream()
        nat.check(self.lib.arkmpc_ctx_sync(self.ctx), "arkmpc_ctx_sync", self.ctx)

    @property
    def launches(self) -> int:
        return int(self.lib.arkmpc_ctx_launch_count(self.ctx))

    @property
    def sm_count(self) -> int:
        return int(self.lib.arkmpc_ctx_sm_count(self.ctx))

    # -- storage --------------------------------------------------------------------------------
    def empty(self, n: int) -> torch.Tensor:
        return torch.empty((n, 4), dtype=torch.int64, device=self.tdev)

    def upload(self, limbs: np.ndarray) -> torch.Tensor:
        """(n,4) uint64 host limbs -> device plane."""
        a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
        return torch.from_numpy(a.view(np.int64)).to(self.tdev)

    @staticmethod
    def download(t: torch.Tensor) -> np.ndarray:
        return t.detach().cpu().numpy().view(np.uint64)

    def key_limbs(self, key) -> np.ndarray:
        if isinstance(key, (int, np.integer)):
            raise TypeError("MAC key must be given as its 4-limb Montgomery image (np.uint64[4])")
        k = np.ascontiguousarray(key, dtype=np.uint64).reshape(4)
        return k

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        if t is None:
            return None
        assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous(), "expected a contiguous int64 CUDA plane"
        return C.c_void_p(t.data_ptr())

    def _call(self, name: str, *args) -> None:
        self.bind_current_stream()
        nat.check(getattr(self.lib, name)(self.ctx, *args), name, self.ctx)

    # -- Beaver multiplication ------------------------------------------------------------------
    def beaver_mask(self, x_share, y_share, a_share, b_share, out: Optional[Planes] = None) -> Planes:
        n = x_share.shape[0]
        d, e = out if out is not None else (self.empty(n), self.empty(n))
        self._call("arkmpc_fr_beaver_mask", self.field, n, self._p(x_share), self._p(y_share), self._p(a_share),
                   self._p(b_share), self._p(d), self._p(e))
        return d, e

    def beaver_recombine(self, party: int, key: np.ndarray, d_mine, e_mine, d_peer, e_peer, a: Planes, b: Planes,
                         c: Planes, out: Optional[Planes] = None, open_out: Optional[Planes] = None,
                         want_open: bool = False):
        n = d_mine.shape[0]
        out_s, out_m = out if out is not None else (self.empty(n), self.empty(n))
        if open_out is None and want_open:
            open_out = (self.empty(n), self.empty(n))
        d_o, e_o = open_out if open_out is not None else (None, None)
        k = self.key_limbs(key)
        self._call("arkmpc_fr_beaver_recombine", self.field, int(party), k.ctypes.data_as(C.c_void_p), n,
                   self._p(d_mine), self._p(e_mine), self._p(d_peer), self._p(e_peer),
                   self._p(a[0]), self._p(a[1]), self._p(b[0]), self._p(b[1]), self._p(c[0]), self._p(c[1]),
                   self._p(out_s), self._p(out_m), self._p(d_o), self._p(e_o))
        return (out_s, out_m), (d_o, e_o)

    def beaver_recombine_sum(self, party: int, key, d_mine, e_mine, d_peer, e_peer, a: Planes, b: Planes, c: Planes) -> Planes:
        """sum_i [x_i * y_i] as one ScalarShare (two 1-row planes): phase 2 fused with the tree-sum that follows it."""
        n = d_mine.shape[0]
        out_s, out_m = self.empty(1), self.empty(1)
        k = self.key_limbs(key)
        self._call("arkmpc_fr_beaver_recombine_sum", self.field, int(party), k.ctypes.data_as(C.c_void_p), n,
                   self._p(d_mine), self._p(e_mine), self._p(d_peer), self._p(e_peer),
                   self._p(a[0]), self._p(a[1]), self._p(b[0]), self._p(b[1]), self._p(c[0]), self._p(c[1]),
                   self._p(out_s), self._p(out_m))
        return out_s, out_m

    # -- public-scalar gates --------------------------------------------------------------------
    def _binary(self, name, a, b, out=None):
        n = a.shape[0]
        out = out if out is not None else self.empty(n)
        self._call(name, self.field, n, self._p(a), self._p(b), self._p(out))
        return out

    def add(self, a, b, out=None): return self._binary("arkmpc_fr_add", a, b, out)
    def sub(self, a, b, out=None): return self._binary("arkmpc_fr_sub", a, b, out)
    def mul(self, a, b, out=None): return self._binary("arkmpc_fr_mul", a, b, out)

    def neg(self, a, out=None):
        out = out if out is not None else self.empty(a.shape[0])
        self._call("arkmpc_fr_neg", self.field, a.shape[0], self._p(a), self._p(out))
        return out

    def scale(self, a, s: np.ndarray, out=None):
        out = out if out is not None else self.empty(a.shape[0])
        k = self.key_limbs(s)
        self._call("arkmpc_fr_scale", self.field, a.shape[0], self._p(a), k.ctypes.data_as(C.c_void_p), self._p(out))
        return out

    def to_mont(self, plain):
        out = self.empty(plain.shape[0])
        self._call("arkmpc_fr_to_mont", self.field, plain.shape[0], self._p(plain), self._p(out))
        return out

    def from_mont(self, mont):
        out = self.empty(mont.shape[0])
        self._call("arkmpc_fr_from_mont", self.field, mont.shape[0], self._p(mont), self._p(out))
        return out

    def batch_inverse(self, a, out=None):
        out = out if out is not None else self.empty(a.shape[0])
        self._call("arkmpc_fr_batch_inverse", self.field, a.shape[0], self._p(a), self._p(out))
        return out

    def fft(self, a, inverse: bool = False):
        """ark-poly Radix2EvaluationDomain fft / ifft of one plane whose length is a power of two."""
        n = a.shape[0]
        assert n > 0 and n & (n - 1) == 0, "the caller pads to the domain size (a power of two)"
        out = self.empty(n)
        self._call("arkmpc_fr_fft", self.field, n.bit_length() - 1, int(inverse), self._p(a), self._p(out))
        return out

    def share_fft(self, a: Planes, inverse: bool = False) -> Planes:
        n = a[0].shape[0]
        assert n > 0 and n & (n - 1) == 0, "the caller pads to the domain size (a power of two)"
        o = (self.empty(n), self.empty(n))
        self._call("arkmpc_fr_share_fft", self.field, n.bit_length() - 1, int(inverse), self._p(a[0]), self._p(a[1]), self._p(o[0]), self._p(o[1]))
        return o

    def random(self, seed: int, first: int, n: int):
        out = self.empty(n)
        self._call("arkmpc_fr_random", self.field, C.c_uint64(seed & (2**64 - 1)), C.c_uint64(first), n, self._p(out))
        return out

    def to_bytes_be(self, a) -> torch.Tensor:
        out = torch.empty((a.shape[0], 32), dtype=torch.uint8, device=self.tdev)
        self.bind_current_stream()
        nat.check(self.lib.arkmpc_fr_to_bytes_be(self.ctx, self.field, a.shape[0], self._p(a), C.c_void_p(out.data_ptr())),
                  "arkmpc_fr_to_bytes_be", self.ctx)
        return out

    # -- share gates ----------------------------------------------------------------------------
    def _share_binary(self, name, a: Planes, b: Planes) -> Planes:
        n = a[0].shape[0]
        o = (self.empty(n), self.empty(n))
        self._call(name, self.field, n, self._p(a[0]), self._p(a[1]), self._p(b[0]), self._p(b[1]), self._p(o[0]), self._p(o[1]))
        return o

    def share_add(self, a, b): return self._share_binary("arkmpc_fr_share_add", a, b)
    def share_sub(self, a, b): return self._share_binary("arkmpc_fr_share_sub", a, b)

    def share_neg(self, a: Planes) -> Planes:
        n = a[0].shape[0]
        o = (self.empty(n), self.empty(n))
        self._call("arkmpc_fr_share_neg", self.field, n, self._p(a[0]), self._p(a[1]), self._p(o[0]), self._p(o[1]))
        return o

    def share_add_public(self, party: int, key, a: Planes, v, sub: bool = False) -> Planes:
        n = a[0].shape[0]
        o = (self.empty(n), self.empty(n))
        k = self.key_limbs(key)
        name = "arkmpc_fr_share_sub_public" if sub else "arkmpc_fr_share_add_public"
        self._call(name, self.field, int(party), k.ctypes.data_as(C.c_void_p), n, self._p(a[0]), self._p(a[1]), self._p(v),
                   self._p(o[0]), self._p(o[1]))
        return o

    def share_mul_public(self, a: Planes, v) -> Planes:
        n = a[0].shape[0]
        o = (self.empty(n), self.empty(n))
        self._call("arkmpc_fr_share_mul_public", self.field, n, self._p(a[0]), self._p(a[1]), self._p(v), self._p(o[0]), self._p(o[1]))
        return o

    def mac_check(self, key, opened, mac):
        out = self.empty(opened.shape[0])
        k = self.key_limbs(key)
        self._call("arkmpc_fr_mac_check", self.field, k.ctypes.data_as(C.c_void_p), opened.shape[0], self._p(opened), self._p(mac), self._p(out))
        return out

    def sum_is_zero(self, mine, peer) -> bool:
        flag = C.c_int(0)
        self._call("arkmpc_fr_sum_is_zero", self.field, mine.shape[0], self._p(mine), self._p(peer), C.byref(flag))
        return bool(flag.value)

    def share_sum(self, a: Planes) -> Planes:
        o = (self.empty(1), self.empty(1))
        self._call("arkmpc_fr_share_sum", self.field, a[0].shape[0], self._p(a[0]), self._p(a[1]), self._p(o[0]), self._p(o[1]))
        return o

    def sum(self, a):
        o = self.empty(1)
        self._call("arkmpc_fr_sum", self.field, a.shape[0], self._p(a), self._p(o))
        return o

    # -- point gates (points: int64 tensors (n, words) in the reference's AoS projective image;
    #    PointShares: (n, 2*words) = {share, mac}) ------------------------------------------------
    def bind_curve(self, curve: str) -> None:
        """Select the curve group for the pt_* calls; its scalar field must be this engine's field."""
        cid = nat.CURVE_IDS[curve]
        if cid != self.field:
            raise ValueError(f"curve {curve} does not match the engine's scalar field {self.field_name}")
        self.curve = cid
        self.point_words = int(self.lib.arkmpc_point_bytes(cid)) // 8

    def _curve(self) -> int:
        if getattr(self, "curve", None) is None:
            self.bind_curve({0: "bn254_g1", 1: "curve25519_edwards"}[self.field])
        return self.curve

    def empty_points(self, n: int, share: bool = False) -> torch.Tensor:
        self._curve()
        return torch.empty((n, self.point_words * (2 if share else 1)), dtype=torch.int64, device=self.tdev)

    def upload_points(self, limbs: np.ndarray) -> torch.Tensor:
        a = np.ascontiguousarray(limbs, dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).to(self.tdev)

    def _pt_binary(self, name, a, b):
        out = torch.empty_like(a)
        n = a.numel() * 8 // int(self.lib.arkmpc_point_bytes(self._curve()))
        self._call(name, self.curve, n, self._p(a), self._p(b), self._p(out))
        return out

    def pt_add(self, a, b): return self._pt_binary("arkmpc_pt_add", a, b)      # also PointShare + PointShare (2n points)
    def pt_sub(self, a, b): return self._pt_binary("arkmpc_pt_sub", a, b)

    def pt_neg(self, a):
        out = torch.empty_like(a)
        n = a.numel() * 8 // int(self.lib.arkmpc_point_bytes(self._curve()))
        self._call("arkmpc_pt_neg", self.curve, n, self._p(a), self._p(out))
        return out

    def pt_share_add_public(self, party: int, key, a_ps, pub, sub: bool = False):
        out = torch.empty_like(a_ps)
        k = self.key_limbs(key)
        name = "arkmpc_pt_share_sub_public" if sub else "arkmpc_pt_share_add_public"
        self._call(name, self._curve(), int(party), k.ctypes.data_as(C.c_void_p), pub.shape[0], self._p(a_ps), self._p(pub), self._p(out))
        return out

    def pt_mul(self, scalars, pts):
        out = torch.empty_like(pts)
        self._call("arkmpc_pt_mul", self._curve(), pts.shape[0], self._p(scalars), self._p(pts), self._p(out))
        return out

    def pt_share_mul_public(self, scalars, a_ps):
        out = torch.empty_like(a_ps)
        self._call("arkmpc_pt_share_mul_public", self._curve(), a_ps.shape[0], self._p(scalars), self._p(a_ps), self._p(out))
        return out

    def pt_mul_authenticated(self, s: Planes, pts):
        out = self.empty_points(pts.shape[0], share=True)
        self._call("arkmpc_pt_mul_authenticated", self._curve(), pts.shape[0], self._p(s[0]), self._p(s[1]), self._p(pts), self._p(out))
        return out

    def pt_mul_generator(self, s: Planes):
        out = self.empty_points(s[0].shape[0], share=True)
        self._call("arkmpc_pt_mul_generator", self._curve(), s[0].shape[0], self._p(s[0]), self._p(s[1]), self._p(out))
        return out

    def pt_mul_generator_public(self, scalars):
        out = self.empty_points(scalars.shape[0])
        self._call("arkmpc_pt_mul_generator_public", self._curve(), scalars.shape[0], self._p(scalars), self._p(out))
        return out

    def pt_beaver_mask(self, x_share, P_ps, a_share, b_share, out=None):
        n = x_share.shape[0]
        d, E = out if out is not None else (self.empty(n), self.empty_points(n))
        self._call("arkmpc_pt_beaver_mask", self._curve(), n, self._p(x_share), self._p(P_ps), self._p(a_share), self._p(b_share),
                   self._p(d), self._p(E))
        return d, E

    def pt_beaver_recombine(self, party: int, key, d_mine, d_peer, E_mine, E_peer, a: Planes, b: Planes, c: Planes, out=None,
                            want_open: bool = False):
        n = d_mine.shape[0]
        out = out if out is not None else self.empty_points(n, share=True)
        d_o, E_o = (self.empty(n), self.empty_points(n)) if want_open else (None, None)
        k = self.key_limbs(key)
        self._call("arkmpc_pt_beaver_recombine", self._curve(), int(party), k.ctypes.data_as(C.c_void_p), n, self._p(d_mine), self._p(d_peer),
                   self._p(E_mine), self._p(E_peer), self._p(a[0]), self._p(a[1]), self._p(b[0]), self._p(b[1]), self._p(c[0]), self._p(c[1]),
                   self._p(out), self._p(d_o), self._p(E_o))
        return out, (d_o, E_o)

    def pt_mac_check(self, key, opened, a_ps):
        out = torch.empty_like(opened)
        k = self.key_limbs(key)
        self._call("arkmpc_pt_mac_check", self._curve(), k.ctypes.data_as(C.c_void_p), opened.shape[0], self._p(opened), self._p(a_ps), self._p(out))
        return out

    def pt_sum_is_identity(self, mine, peer) -> bool:
        flag = C.c_int(0)
        self._call("arkmpc_pt_sum_is_identity", self._curve(), mine.shape[0], self._p(mine), self._p(peer), C.byref(flag))
        return bool(flag.value)

    def pt_validate(self, pts) -> bool:
        """True iff every point is on the curve and in the prime-order subgroup (what the recombination needs of E_peer)."""
        flag = C.c_int(0)
        n = pts.numel() * 8 // int(self.lib.arkmpc_point_bytes(self._curve()))
        self._call("arkmpc_pt_validate", self.curve, n, self._p(pts), C.byref(flag))
        return bool(flag.value)

    def pt_sum(self, pts) -> torch.Tensor:
        out = self.empty_points(1)
        n = pts.numel() * 8 // int(self.lib.arkmpc_point_bytes(self._curve()))
        self._call("arkmpc_pt_sum", self.curve, n, self._p(pts), self._p(out))
        return out

    def pt_share_sum(self, a_ps) -> torch.Tensor:
        out = self.empty_points(1, share=True)
        self._call("arkmpc_pt_share_sum", self._curve(), a_ps.shape[0], self._p(a_ps), self._p(out))
        return out

    def pt_msm(self, scalars, pts) -> torch.Tensor:
        """sum_i scalars[i] * pts[i] (public MSM, bucket method)."""
        out = self.empty_points(1)
        self._call("arkmpc_pt_msm", self._curve(), pts.shape[0], self._p(scalars), self._p(pts), self._p(out))
        return out

    def pt_msm_authenticated(self, s: Planes, pts) -> torch.Tensor:
        out = self.empty_points(1, share=True)
        self._call("arkmpc_pt_msm_authenticated", self._curve(), pts.shape[0], self._p(s[0]), self._p(s[1]), self._p(pts), self._p(out))
        return out

    def pt_normalize(self, pts) -> torch.Tensor:
        """(n, words) projective -> (n, 8) canonical affine (x, y) Montgomery limbs."""
        n = pts.numel() * 8 // int(self.lib.arkmpc_point_bytes(self._curve()))
        out = torch.empty((n, 8), dtype=torch.int64, device=self.tdev)
        self._call("arkmpc_pt_normalize", self.curve, n, self._p(pts), self._p(out))
        return out

    def pt_from_affine(self, xy) -> torch.Tensor:
        """(n, 8) canonical affine (x, y) Montgomery limbs -> (n, words) projective image with Z = 1."""
        n = xy.shape[0]
        out = self.empty_points(n)
        self._call("arkmpc_pt_from_affine", self.curve, n, self._p(xy), self._p(out))
        return out

    # -- layout ---------------------------------------------------------------------------------
    def share_unzip(self, aos: torch.Tensor) -> Planes:
        """(n, 8) AoS ScalarShare image -> (share, mac) planes."""
        n = aos.shape[0]
        o = (self.empty(n), self.empty(n))
        self._call("arkmpc_share_unzip", n, self._p(aos), self._p(o[0]), self._p(o[1]))
        return o

    def share_zip(self, a: Planes) -> torch.Tensor:
        n = a[0].shape[0]
        out = torch.empty((n, 8), dtype=torch.int64, device=self.tdev)
        self._call("arkmpc_share_zip", n, self._p(a[0]), self._p(a[1]), self._p(out))
        return out

    # -- host-buffer end-to-end path ------------------------------------------------------------
    def batch_mul_begin_host(self, party: int, key, x: np.ndarray, y: np.ndarray, a: np.ndarray, b: np.ndarray,
                             c: np.ndarray, de_mine: np.ndarray):
        """x..c: (n,8) uint64 AoS host arrays; de_mine: (2n,4) uint64 host output.  Returns a session handle."""
        n = x.shape[0]
        k = self.key_limbs(key)
        sess = C.c_void_p()
        hp = lambda arr: C.c_void_p(arr.ctypes.data)
        self._call("arkmpc_fr_batch_mul_begin_host", self.field, int(party), k.ctypes.data_as(C.c_void_p), n,
                   hp(x), hp(y), hp(a), hp(b), hp(c), hp(de_mine), C.byref(sess))
        return sess

    def batch_mul_begin_host_shares(self, party: int, key, x_share: np.ndarray, y_share: np.ndarray, a: np.ndarray, b: np.ndarray,
                                    c: np.ndarray, de_mine: np.ndarray):
        """As batch_mul_begin_host with x, y given as (n,4) planes of their share halves (the operands' MACs are not inputs)."""
        n = x_share.shape[0]
        k = self.key_limbs(key)
        sess = C.c_void_p()
        hp = lambda arr: C.c_void_p(arr.ctypes.data)
        self._call("arkmpc_fr_batch_mul_begin_host_shares", self.field, int(party), k.ctypes.data_as(C.c_void_p), n,
                   hp(x_share), hp(y_share), hp(a), hp(b), hp(c), hp(de_mine), C.byref(sess))
        return sess

    def batch_mul_finish_host(self, sess, de_peer: np.ndarray, out: np.ndarray, de_open: Optional[np.ndarray] = None) -> None:
        hp = lambda arr: C.c_void_p(arr.ctypes.data) if arr is not None else None
        nat.check(self.lib.arkmpc_fr_batch_mul_finish_host(sess, hp(de_peer), hp(out), hp(de_open)),
                  "arkmpc_fr_batch_mul_finish_host", self.ctx)

    def host_path_bytes(self, n: int, with_open: bool = False) -> Tuple[int, int]:
        """(host->device, device->host) bytes one party's begin + finish move for n gates."""
        up, down = C.c_uint64(0), C.c_uint64(0)
        self._call("arkmpc_fr_batch_mul_host_bytes", n, int(with_open), C.byref(up), C.byref(down))
        return int(up.value), int(down.value)

    def validate(self, a) -> bool:
        """True iff every element of the plane is a canonical residue (< p)."""
        flag = C.c_int(0)
        self._call("arkmpc_fr_validate", self.field, a.shape[0], self._p(a), C.byref(flag))
        return bool(flag.value)

    def pinned_empty(self, shape, dtype=np.uint64) -> np.ndarray:
        """Pinned host array (torch-owned pinned storage viewed as numpy)."""
        t = torch.empty(tuple(shape), dtype=torch.int64).pin_memory()
        arr = t.numpy().view(dtype)
        arr_base = arr  # keep the tensor alive through the array's base chain
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(t)
        return arr_base
