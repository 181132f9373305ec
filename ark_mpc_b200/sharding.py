"""Multi-GPU: a batch of independent gates sharded by contiguous index range, one process per GPU.

Every gate of the hot path is element-wise (/root/reference/online-phase/src/algebra/scalar/authenticated_scalar.rs:479-484,
:517-523, :679-684, :906-911), so K1/K2 need no inter-GPU traffic; the two-party exchange is shard-aligned.  The one
collective north_star names is an all-gather of the opened d || e of `batch_open` when every device needs the whole vector:

  * `all_gather_rows`            torch.distributed all-gather (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests)
  * `OpenGather.recombine_gather` the fused form: the recombine kernel stores this rank's opened rows into every rank's
                                 gathered planes through CUDA IPC peer mappings (arkmpc_fr_beaver_recombine_gather)

Field elements cannot be summed by a collective (no modular `ncclSum`): cross-GPU sums all-gather the per-rank partial
ScalarShares (64 B each) and add them mod p locally (`all_reduce_share_sum`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _native as nat


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous index range [k*n/G, (k+1)*n/G) of rank k."""
    return n * rank // world, n * (rank + 1) // world


def shard_sizes(n: int, world: int) -> List[int]:
    return [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]


def all_gather_rows(local: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Concatenate every rank's rows along dim 0 (equal shard sizes).  Works on CUDA (NCCL) and CPU (gloo) tensors."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        chunks = list(out.chunk(world, dim=0))
        dist.all_gather(chunks, local.contiguous(), group=group)
    return out


def all_gather_rows_ragged(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """As above for shard sizes that differ by one (n not divisible by world): pad to the largest shard, gather, trim."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local.shape[0] == sizes[rank], "local shard does not match shard_bounds"
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = all_gather_rows(pad, group)
    return torch.cat([full[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


def all_reduce_share_sum(engine, partial, group=None):
    """Sum of per-rank partial ScalarShares: all-gather of the (1,4) share and mac planes, then a local modular sum."""
    s = all_gather_rows(partial[0], group)
    m = all_gather_rows(partial[1], group)
    return engine.share_sum((s, m))


class _DevArray:
    """Exposes a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i8", "data": (ptr, False), "version": 2}


class OpenGather:
    """Per-rank gathered planes for the opened d and e of a sharded batch_mul.  `planes()` are (world*n, 4) int64 CUDA tensors
    over memory owned by this object.  Two transports behind the same calls:

      * "multicast"  the planes of all ranks are bound to one NVSwitch multicast window (arkmpc_mc_*): the recombine kernel stores
                     each opened element once and the switch replicates it into every rank's copy;
      * "ipc"        CUDA IPC peer mappings of every other rank's planes: the kernel stores each element world times.

    transport="auto" is "ipc": an all-gather is bound by what every rank must RECEIVE ((world-1)/world of the gathered planes),
    which multicast does not reduce — it only cuts the sender's egress, and adds the loop-back of the rank's own rows — and
    measured on 2 x B200 the per-peer stores are 3x faster (0.106 vs 0.323 ms for 2^20 rows per rank, profiles/r02c_*).
    transport="multicast" (or ARKMPC_GATHER=multicast) selects the window explicitly; "multicast_or_ipc" falls back to IPC when
    the device, the driver or the container cannot build it."""

    def __init__(self, engine, n_local: int, group=None, transport: str = "auto"):
        import os

        self.E, self.n, self.group = engine, n_local, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("OpenGather supports up to 8 ranks (one NVSwitch domain)")
        transport = os.environ.get("ARKMPC_GATHER", transport)
        if transport not in ("auto", "ipc", "multicast", "multicast_or_ipc"):
            raise ValueError(f"unknown gather transport {transport!r}")
        self._own, self._peers, self._mc = [], [], None
        self.transport, self.fallback_reason = None, None
        if transport in ("multicast", "multicast_or_ipc") and self.world > 1:
            try:
                self._init_multicast()
                self.transport = "multicast"
            except _McUnavailable as e:
                if transport == "multicast":
                    raise
                self.fallback_reason = str(e)
        if self.transport is None:
            self._init_ipc()
            self.transport = "ipc"
        dist.barrier(group=group)

    # -- multicast window ------------------------------------------------------------------------------------------------
    def _all_ok(self, ok: bool, why: str = "") -> None:
        """Collective agreement: every rank proceeds, or every rank tears down and falls back."""
        flags: List[object] = [None] * self.world
        dist.all_gather_object(flags, (bool(ok), why), group=self.group)
        bad = [w for o, w in flags if not o]
        if bad:
            self._close_mc()
            raise _McUnavailable(bad[0] or "a rank could not build the multicast window")

    def _init_multicast(self) -> None:
        lib, ctx = self.E.lib, self.E.ctx
        plane = self.world * self.n * 32
        sup = C.c_int(0)
        lib.arkmpc_mc_supported(ctx, C.byref(sup))
        self._all_ok(bool(sup.value), "device / driver without multicast support")
        h = C.c_void_p()
        owner = [None]
        ok, why = True, ""
        if self.rank == 0:
            rc = lib.arkmpc_mc_open(ctx, 2 * plane, self.world, 0, 0, -1, C.byref(h))
            if rc == nat.OK:
                pid, fd = C.c_int(0), C.c_int(0)
                lib.arkmpc_mc_export(h, C.byref(pid), C.byref(fd))
                owner = [(pid.value, fd.value)]
                self._mc = h
            else:
                ok, why = False, (lib.arkmpc_last_error(ctx) or b"").decode()
        dist.broadcast_object_list(owner, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        if self.rank != 0 and owner[0] is not None:
            rc = lib.arkmpc_mc_open(ctx, 2 * plane, self.world, self.rank, owner[0][0], owner[0][1], C.byref(h))
            if rc == nat.OK:
                self._mc = h
            else:
                ok, why = False, (lib.arkmpc_last_error(ctx) or b"").decode()
        elif self.rank != 0:
            ok, why = False, "the owner rank could not create the multicast object"
        self._all_ok(ok, why)  # doubles as the barrier "every device has been added"
        loc, mc = C.c_void_p(), C.c_void_p()
        rc = lib.arkmpc_mc_bind(self._mc, C.byref(loc), C.byref(mc))
        self._all_ok(rc == nat.OK, (lib.arkmpc_last_error(ctx) or b"").decode() if rc != nat.OK else "")
        self._mc_ptrs = (mc.value, mc.value + plane)
        shape = (self.world * self.n, 4)
        self.d_all = torch.as_tensor(_DevArray(loc.value, shape), device=self.E.tdev)
        self.e_all = torch.as_tensor(_DevArray(loc.value + plane, shape), device=self.E.tdev)

    def _close_mc(self) -> None:
        if self._mc is not None:
            self.E.lib.arkmpc_mc_close(self._mc)
            self._mc = None

    # -- CUDA IPC peer mappings ------------------------------------------------------------------------------------------
    def _init_ipc(self) -> None:
        lib, ctx = self.E.lib, self.E.ctx
        n_local = self.n
        handles = []
        for _ in range(2):
            p = C.c_void_p()
            nat.check(lib.arkmpc_malloc(ctx, self.world * n_local * 32, C.byref(p)), "arkmpc_malloc", ctx)
            self._own.append(p)
            h = (C.c_uint8 * 64)()
            nat.check(lib.arkmpc_ipc_export(ctx, p, h), "arkmpc_ipc_export", ctx)
            handles.append(bytes(h))
        everyone: List[object] = [None] * self.world
        dist.all_gather_object(everyone, handles, group=self.group)
        self._ptrs = [[None] * self.world, [None] * self.world]  # [d|e][rank]
        for r in range(self.world):
            for which in range(2):
                if r == self.rank:
                    self._ptrs[which][r] = self._own[which].value
                else:
                    q = C.c_void_p()
                    buf = (C.c_uint8 * 64).from_buffer_copy(everyone[r][which])
                    nat.check(lib.arkmpc_ipc_import(ctx, buf, C.byref(q)), "arkmpc_ipc_import", ctx)
                    self._peers.append(q)
                    self._ptrs[which][r] = q.value
        self._arr = [(C.c_void_p * self.world)(*self._ptrs[which]) for which in range(2)]
        shape = (self.world * n_local, 4)
        self.d_all = torch.as_tensor(_DevArray(self._own[0].value, shape), device=self.E.tdev)
        self.e_all = torch.as_tensor(_DevArray(self._own[1].value, shape), device=self.E.tdev)

    def planes(self):
        return self.d_all, self.e_all

    def my_rows(self):
        lo, hi = self.rank * self.n, (self.rank + 1) * self.n
        return self.d_all[lo:hi], self.e_all[lo:hi]

    def recombine_gather(self, party: int, key, d_mine, e_mine, d_peer, e_peer, a, b, c, out):
        """Fused K2 + all-gather.  The caller synchronises (stream sync + barrier) before reading rows written by peers."""
        E = self.E
        k = E.key_limbs(key)
        common = (E.field, int(party), k.ctypes.data_as(C.c_void_p), self.n, E._p(d_mine), E._p(e_mine), E._p(d_peer), E._p(e_peer),
                  E._p(a[0]), E._p(a[1]), E._p(b[0]), E._p(b[1]), E._p(c[0]), E._p(c[1]), E._p(out[0]), E._p(out[1]), self.world, self.rank)
        if self.transport == "multicast":
            E._call("arkmpc_fr_beaver_recombine_gather_mc", *common, C.c_void_p(self._mc_ptrs[0]), C.c_void_p(self._mc_ptrs[1]))
        else:
            E._call("arkmpc_fr_beaver_recombine_gather", *common, self._arr[0], self._arr[1])

    def recombine_then_nccl(self, party: int, key, d_mine, e_mine, d_peer, e_peer, a, b, c, out):
        """Baseline: K2 writes this rank's opened rows, then two NCCL all-gathers (in place on the gathered planes)."""
        E = self.E
        dr, er = self.my_rows()
        E.beaver_recombine(party, key, d_mine, e_mine, d_peer, e_peer, a, b, c, out=out, open_out=(dr, er))
        dist.all_gather_into_tensor(self.d_all, dr, group=self.group)
        dist.all_gather_into_tensor(self.e_all, er, group=self.group)

    def close(self):
        lib, ctx = self.E.lib, self.E.ctx
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.d_all = self.e_all = None
        for q in self._peers:
            lib.arkmpc_ipc_release(ctx, q)
        self._close_mc()
        dist.barrier(group=self.group)
        for p in self._own:
            lib.arkmpc_free(ctx, p)
        self._peers, self._own = [], []


class _McUnavailable(RuntimeError):
    pass


class NativeAllGather:
    """`arkmpc_allgather_open` — the plain NCCL all-gather behind the C ABI (what a Rust host binds): a communicator owned by the
    native context, its id distributed here through torch.distributed (any host channel does)."""

    def __init__(self, engine, group=None):
        self.E, self.group = engine, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        ident = [None]
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            nat.check(engine.lib.arkmpc_nccl_unique_id(buf), "arkmpc_nccl_unique_id", engine.ctx)
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        buf = (C.c_uint8 * 128).from_buffer_copy(ident[0])
        nat.check(engine.lib.arkmpc_nccl_init(engine.ctx, self.world, self.rank, buf), "arkmpc_nccl_init", engine.ctx)

    def allgather_open(self, d_local: torch.Tensor, e_local: Optional[torch.Tensor], d_all: torch.Tensor, e_all: Optional[torch.Tensor]) -> None:
        E = self.E
        E._call("arkmpc_allgather_open", d_local.shape[0], E._p(d_local), E._p(e_local), E._p(d_all), E._p(e_all))

    def close(self) -> None:
        self.E.lib.arkmpc_nccl_destroy(self.E.ctx)
