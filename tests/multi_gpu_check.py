#!/usr/bin/env python
"""Multi-GPU check, launched with torchrun (one rank per GPU): a batch_mul sharded by index range; the opened d || e
gathered on every rank by (a) K2 + NCCL all-gather (torch.distributed), (b) K2 + arkmpc_allgather_open (NCCL behind the C
ABI), (c) the fused recombine+gather kernel over CUDA IPC peer mappings and (d) the fused kernel with NVSwitch multicast
stores (when the box supports it); all must equal the rows every rank can recompute from the shared seeds, over the WHOLE
gathered planes (peer-written rows included).  Prints one OK line per rank."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from ark_mpc_b200.engine import Engine
from ark_mpc_b200 import sharding as sh


def shard_data(E, n, rank_seed):
    """Both parties' operands of one shard, generated on the device from (seed, rank)."""
    key0 = E.download(E.random(900, 0, 1))[0].copy()
    key1 = E.download(E.random(901, 0, 1))[0].copy()
    key = E.download(E.add(E.upload(key0.reshape(1, 4)), E.upload(key1.reshape(1, 4))))[0].copy()

    def shared(s, val=None):
        v = E.random(rank_seed + s, 0, n) if val is None else val
        s0, m0 = E.random(rank_seed + s + 1, 0, n), E.random(rank_seed + s + 2, 0, n)
        return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

    xv, x0, x1 = shared(10)
    yv, y0, y1 = shared(20)
    av, a0, a1 = shared(30)
    bv, b0, b1 = shared(40)
    _, c0, c1 = shared(50, E.mul(av, bv))
    P = [dict(key=key0, x=x0, y=y0, a=a0, b=b0, c=c0), dict(key=key1, x=x1, y=y1, a=a1, b=b1, c=c1)]
    return P, E.sub(xv, av), E.sub(yv, bv), E.mul(xv, yv)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    E = Engine(local, "bn254_fr")
    P, d_want, e_want, xy = shard_data(E, n, 7919 * rank)
    # what every rank expects to see after the gather: rank r's opened rows, recomputed from r's seeds
    want_d = torch.cat([shard_data(E, n, 7919 * r)[1] for r in range(world)], dim=0)
    want_e = torch.cat([shard_data(E, n, 7919 * r)[2] for r in range(world)], dim=0)
    de = [E.beaver_mask(P[p]["x"][0], P[p]["y"][0], P[p]["a"][0], P[p]["b"][0]) for p in (0, 1)]
    used = []
    for transport in ("ipc", "multicast_or_ipc"):
        G = sh.OpenGather(E, n, transport=transport)
        if transport != "ipc" and G.transport == "ipc":
            print(f"rank {rank}: multicast unavailable ({G.fallback_reason}); ipc already covered", flush=True)
            G.close()
            continue
        used.append(G.transport)
        for mode in ("nccl", "fused"):
            G.d_all.zero_()
            G.e_all.zero_()
            torch.cuda.synchronize()
            dist.barrier()
            out = (E.empty(n), E.empty(n))
            args = (0, P[0]["key"], de[0][0], de[0][1], de[1][0], de[1][1], P[0]["a"], P[0]["b"], P[0]["c"], out)
            (G.recombine_then_nccl if mode == "nccl" else G.recombine_gather)(*args)
            torch.cuda.synchronize()
            dist.barrier()
            assert torch.equal(G.d_all, want_d) and torch.equal(G.e_all, want_e), f"rank {rank}: {G.transport}/{mode} gather differs"
            out1 = E.beaver_recombine(1, P[1]["key"], de[1][0], de[1][1], de[0][0], de[0][1], P[1]["a"], P[1]["b"], P[1]["c"])[0]
            assert torch.equal(E.add(out[0], out1[0]), xy), f"rank {rank}: product shares do not open to x*y"
        G.close()
    # the plain collective behind the C ABI: arkmpc_nccl_init + arkmpc_allgather_open
    NG = sh.NativeAllGather(E)
    d_all, e_all = E.empty(world * n), E.empty(world * n)
    d_all.zero_()
    e_all.zero_()
    out = (E.empty(n), E.empty(n))
    _, (d_open, e_open) = E.beaver_recombine(0, P[0]["key"], de[0][0], de[0][1], de[1][0], de[1][1], P[0]["a"], P[0]["b"], P[0]["c"], out=out, want_open=True)
    NG.allgather_open(d_open, e_open, d_all, e_all)
    torch.cuda.synchronize()
    dist.barrier()
    assert torch.equal(d_all, want_d) and torch.equal(e_all, want_e), f"rank {rank}: arkmpc_allgather_open differs"
    NG.close()
    # cross-GPU sum of the inner product: partial ScalarShares gathered, added mod p locally
    part = E.share_sum(out)
    tot = sh.all_reduce_share_sum(E, part)
    parts = sh.all_gather_rows(part[0])
    assert torch.equal(tot[0], E.sum(parts))
    print(f"rank {rank}/{world}: multi-GPU open gather OK (nccl == arkmpc_allgather_open == fused[{', '.join(used)}] == expected, n={n}/rank)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
