"""World-size-2 `gloo` tests (CPU) of the multi-GPU host logic: contiguous index-range sharding and the all-gather of
opened rows (ark_mpc_b200/sharding.py).  The per-gate arithmetic needs no collective, so what is covered
here is exactly the part of the N > 1 path that is not a kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ark_mpc_b200 import sharding as sh


def test_shard_bounds_partition():
    for n in (0, 1, 7, 1024, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            lo = 0
            for r in range(world):
                a, b = sh.shard_bounds(n, r, world)
                assert a == lo and b >= a
                lo = b
            assert lo == n
            sizes = sh.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 4, dtype=torch.int64).reshape(n_total, 4) * 0x9E3779B97F4A7C15 % (1 << 62)
        lo, hi = sh.shard_bounds(n_total, rank, world)
        got = sh.all_gather_rows_ragged(full[lo:hi].clone(), n_total)
        ok = torch.equal(got, full)
        if n_total % world == 0:
            ok = ok and torch.equal(sh.all_gather_rows(full[lo:hi].clone()), full)
        # partial sums: every rank contributes one row; the gathered plane is what the local modular sum consumes
        part = full[lo:lo + 1].clone() if hi > lo else torch.zeros((1, 4), dtype=torch.int64)
        rows = sh.all_gather_rows(part)
        ok = ok and rows.shape == (world, 4) and torch.equal(rows[rank], part[0])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [64, 1001])
def test_all_gather_rows_world2(n_total):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
