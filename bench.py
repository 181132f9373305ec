#!/usr/bin/env python
"""bench.py — authenticated Beaver multiplications/sec (BASELINE.json metric) on N B200s.

A step = one `AuthenticatedScalarResult::batch_mul` of 2^log2_batch gates for BOTH parties
(/root/reference/online-phase/src/algebra/scalar/authenticated_scalar.rs:848-879; both parties on the
measured device with the d/e exchange by pointer, as the reference's bench does with its in-memory
MockNetwork, benches/batch_ops.rs:20-40): mask(p0), mask(p1), fused recombine(p0), fused recombine(p1).

  value     two-party multiplications / s, operands resident in HBM (planar layout), CUDA events
  e2e       same metric through the host-buffer C ABI (arkmpc_fr_batch_mul_{begin,finish}_host):
            pinned host AoS inputs, H2D/D2H inside the timed region
  roofline  the dominant kernel (fused recombine, 384 algorithmic B/gate) vs measured HBM peak
  cpu_baseline / --impl reference   the CPU restatement of the reference path (oracle/ark_oracle.c,
            kind "port": the Rust reference cannot be built here) on all host cores
  configs   one object per BASELINE.json config, each with value / roofline / cpu_baseline:
            [0] n = 1024 share + batch_mul + open_authenticated over Curve25519 Fr through the fabric mirror
                (the reference bench's shape, benches/batch_ops.rs:20-40), [1] = the headline above,
            [2] 2^20 AuthenticatedPoint scalar-muls (Curve25519), [3] inner product of length 2^22,
            [4] 2^24 Beaver muls sharded over the N GPUs with the batch_open all-gather INSIDE the step (N > 1)

N > 1: one process per GPU (torchrun), the batch is sharded by index range; `value` has NO data-path
collective (every gate is element-wise; weak scaling, 2^log2_batch gates per GPU) and
`value_with_open_gather` is the same step with party 0's opened d || e all-gathered onto every rank
from inside the recombine kernel (NVSwitch multicast stores, or per-peer stores when multicast is
unavailable), verified over the WHOLE gathered planes against rows recomputed from every rank's seed.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "authenticated_beaver_mults_per_sec"
UNIT = "mults/s"
BYTES_RECOMBINE = 384  # K2 reads 4x32 + 3x64, writes 64
BYTES_MASK = 192       # K1 reads 4x32, writes 2x32
IMAD_WIDE_PER_GATE = 466  # SASS count of beaver_recombine_kernel<Bn254Fr> (tools/sass_hist.sh)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(kernel_substr="beaver_recombine_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py writes the 'traffic ... per launch' lines)."""
    import glob
    import re

    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*recombine_full.txt"))):
        vals, cur = [], None
        for line in open(path):
            if line.startswith("## "):
                cur = line
            m = re.search(r"traffic \(dram read\+write\) bytes per launch\s+(\d+)", line)
            if m and cur and kernel_substr in cur and "pt_" + kernel_substr not in cur:
                vals.append(int(m.group(1)))
        if vals:
            best = (sum(vals) / len(vals), os.path.relpath(path, ROOT))
    return best


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons; rows are time-stamped on arrival so that only the
    samples that fall inside the timed region are summarised."""

    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw"]
    REASONS = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.err = index, [], None, None

    def _spawn(self, prefix):
        q = ",".join(self.FIELDS + [f"{prefix}.{r}" for r in self.REASONS])
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def start(self):
        try:
            self.proc = self._spawn("clocks_event_reasons")
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:  # nvidia-smi missing
            self.proc, self.err = None, repr(e)

    def _read(self):
        for prefix in ("clocks_event_reasons", "clocks_throttle_reasons"):
            got = False
            for line in self.proc.stdout:
                cells = [c.strip() for c in line.split(",")]
                try:
                    float(cells[0])
                except Exception:
                    continue  # an error line (unknown field on an older nvidia-smi)
                got = True
                self.rows.append((time.perf_counter(), cells))
            if got or prefix == "clocks_throttle_reasons":
                return
            try:
                self.proc = self._spawn("clocks_throttle_reasons")
            except Exception:
                return

    def wait_first(self, timeout=5.0):
        t0 = time.perf_counter()
        while not self.rows and self.proc and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvidia-smi unavailable {self.err or ''}".strip()]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [(t, r) for t, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for t, r in rows if t_begin is None or (t_begin <= t <= t_end + 0.06)]
        window = "timed region"
        if len(inside) < 2:
            inside, window = [r for _, r in rows], "whole bench (timed region shorter than the sampling period)"
        sm = sorted(float(r[0]) for r in inside)
        reasons = [nm for k, nm in enumerate(self.REASONS) if any(r[3 + k].lower().startswith("active") for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(inside[0][1]), "reasons": reasons,
                "samples": len(inside), "window": window, "power_w_max": max(float(r[2]) for r in inside)}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (C restatement, all host threads), same metric/config."""
    if rank != 0:
        return
    import numpy as np

    from oracle import coracle as co
    from tests.util import TwoPartyData, aos

    fid = {"bn254_fr": 0, "curve25519_fr": 1}[args.field]
    n_full = 1 << args.log2_batch
    cores = os.cpu_count() or 1
    D = TwoPartyData(fid, n_full, seed=0xA11CE)
    g = lambda t: (aos(*t[0]), aos(*t[1]))
    full = (g(D.x), g(D.y), g(D.a), g(D.b), g(D.c))
    # calibrate on 2^16 gates, then bound the per-step sample so that warmup+steps stay within ~budget seconds
    cut = lambda m: tuple((t[0][:m], t[1][:m]) for t in full)
    m0 = min(n_full, 1 << 16)
    co.two_party_batch_mul(fid, cores, D.keys, *cut(m0), want_open=False)
    t0 = time.perf_counter()
    co.two_party_batch_mul(fid, cores, D.keys, *cut(m0), want_open=False)
    per_gate = (time.perf_counter() - t0) / m0
    budget = 90.0
    n = n_full
    while n > (1 << 14) and per_gate * n * (args.steps + args.warmup) > budget:
        n >>= 1
    ins = (D.keys,) + cut(n)
    for _ in range(args.warmup):
        co.two_party_batch_mul(fid, cores, *ins, want_open=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.two_party_batch_mul(fid, cores, *ins, want_open=False)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "dtype_note": "4 x u64 limbs, Montgomery CIOS (the reference's arkworks representation)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} of the 2^{args.log2_batch} gates (two-party batch_mul) per step, {args.steps} steps, "
                                   "unfused reference gate sequence (oracle/ark_oracle.c), static index partition"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Rust reference unbuildable here (no cargo/rustc, arkworks not vendored): C restatement of its gate "
                "sequence, arithmetic only (omits the reference executor's per-element bookkeeping, so it flatters the reference)",
    }
    emit_json(line)


def workload_config(args, world):
    return {"workload": f"2^{args.log2_batch} authenticated scalar Beaver muls over {args.field} per GPU, both parties, mock net "
                        f"(BASELINE.json configs[1])",
            "field": args.field, "log2_batch_per_gpu": args.log2_batch, "parties": 2,
            "sharding": f"index-range x{world}, no data-path collective",
            "l2_hygiene": "inputs larger than L2: ~0.9 GB touched per step vs 126 MB L2",
            "launch_order": "mask(p0) mask(p1) recombine(p0) recombine(p1); " + (
                "every launch ordered after its predecessor" if (args.no_hint or args.hint_level == 0) else
                "the second launch of each phase carries the independence hint (overlaps the first's drain)" if args.hint_level == 1 else
                "alternate steps write alternate d/e planes; every launch but recombine(p0) carries the independence hint")}


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU e2e: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, so that the pinned
    staging buffers (first touch) and the H2D/D2H traffic stay on the local socket.  Best effort; returns a dict
    describing what was found (node -1 = the platform reports no NUMA affinity for the device, e.g. a single-node VM)."""
    info = {"node": None, "cpus_bound": None, "host_nodes": None}
    try:
        import torch

        info["host_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        info["node"] = int(open(f"{base}/numa_node").read().strip())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus and len(cpus) < (os.cpu_count() or 0):
            os.sched_setaffinity(0, cpus)
            info["cpus_bound"] = len(cpus)
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)[:120]
    return info


class BeaverData:
    """Synthetic two-party inputs of one shard, generated on the device: plaintext x, y, a, b uniform in
    [0, p) from (seed, index), c = a*b, every value and its MAC split additively; one MAC key for the whole job."""

    def __init__(self, E, n, seed, key_seed):
        self.E, self.n, self.seed = E, n, seed
        self.key0, self.key1 = (E.download(E.random(key_seed + p, 0, 1))[0].copy() for p in (0, 1))
        self.key = E.download(E.add(E.upload(self.key0.reshape(1, 4)), E.upload(self.key1.reshape(1, 4))))[0].copy()
        self.keys = (self.key0, self.key1)
        key = self.key

        def shared(s, val=None):
            v = E.random(s, 0, n) if val is None else val
            s0, m0 = E.random(s + 1, 0, n), E.random(s + 2, 0, n)
            return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

        self.xv, x0, x1 = shared(seed + 10)
        self.yv, y0, y1 = shared(seed + 20)
        self.av, a0, a1 = shared(seed + 30)
        self.bv, b0, b1 = shared(seed + 40)
        _, c0, c1 = shared(seed + 50, E.mul(self.av, self.bv))
        self.X, self.Y, self.A, self.B, self.C = (x0, x1), (y0, y1), (a0, a1), (b0, b1), (c0, c1)
        self.P = [dict(key=self.key0, x=x0, y=y0, a=a0, b=b0, c=c0), dict(key=self.key1, x=x1, y=y1, a=a1, b=b1, c=c1)]

    @staticmethod
    def opened_rows(E, n, seed):
        """The opened d = x - a and e = y - b of the shard generated from `seed` (any rank can recompute any shard)."""
        return E.sub(E.random(seed + 10, 0, n), E.random(seed + 30, 0, n)), E.sub(E.random(seed + 20, 0, n), E.random(seed + 40, 0, n))


def shard_seed(base, rank):
    return base + 7919 * rank


def time_steps(step, steps, warmup, stream, world, dist, torch):
    """`warmup` untimed + `steps` timed calls of step() on `stream`, barrier + synchronize on both sides, CUDA events,
    max over ranks.  Returns ms per step."""
    for _ in range(warmup):
        step()
    stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        step()
    ev1.record(stream)
    stream.synchronize()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def measure_open_gather(E, stream, D, n, rank, world, steps, warmup, base_seed, want_nccl=True):
    """The step with north_star's one collective inside it: K1(p0) K1(p1) K2(p1) then party 0's K2 with the opened d || e
    all-gathered onto every rank from inside the kernel (sharding.OpenGather: NVSwitch multicast stores when available, else
    per-peer stores).  After the timed loop EVERY row of the gathered planes — the rows peers wrote included — is compared
    with rows recomputed from the owning rank's seed."""
    import torch
    import torch.distributed as dist

    from ark_mpc_b200 import sharding as sh

    P = D.P
    de = [(E.empty(n), E.empty(n)) for _ in range(2)]
    out = [(E.empty(n), E.empty(n)) for _ in range(2)]
    G = sh.OpenGather(E, n)
    res = {"transport": G.transport, "rows_per_rank": n, "bytes_gathered_per_rank": 64 * n * world}

    def masks():
        E.beaver_mask(P[0]["x"][0], P[0]["y"][0], P[0]["a"][0], P[0]["b"][0], out=de[0])
        E.hint_independent()
        E.beaver_mask(P[1]["x"][0], P[1]["y"][0], P[1]["a"][0], P[1]["b"][0], out=de[1])

    def k2_p1():
        E.beaver_recombine(1, P[1]["key"], de[1][0], de[1][1], de[0][0], de[0][1], P[1]["a"], P[1]["b"], P[1]["c"], out=out[1])

    args0 = lambda: (0, P[0]["key"], de[0][0], de[0][1], de[1][0], de[1][1], P[0]["a"], P[0]["b"], P[0]["c"], out[0])

    def step_fused():
        masks()
        k2_p1()
        E.hint_independent()  # party 0's recombine + gather does not depend on party 1's recombine
        G.recombine_gather(*args0())

    def step_nccl():
        masks()
        k2_p1()
        G.recombine_then_nccl(*args0())

    def verify(tag):
        stream.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        bad = 0
        for r in range(world):
            d_want, e_want = BeaverData.opened_rows(E, n, shard_seed(base_seed, r))
            bad += int(not torch.equal(G.d_all[r * n:(r + 1) * n], d_want)) + int(not torch.equal(G.e_all[r * n:(r + 1) * n], e_want))
        t = torch.tensor([bad], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        if int(t.item()):
            raise SystemExit(f"open gather ({tag}): gathered planes differ from the rows recomputed from the owners' seeds")
        G.d_all.zero_()
        G.e_all.zero_()
        torch.cuda.synchronize()
        dist.barrier()

    ms_f = time_steps(step_fused, steps, warmup, stream, world, dist, torch)
    verify("fused")
    res["ms_per_step"] = ms_f
    res["verified"] = f"all {2 * world} gathered row blocks (2^{n.bit_length() - 1} rows each, peer-written included) equal the rows recomputed from each owner's seed, on every rank"
    # the kernel alone: fused recombine + gather vs recombine followed by NCCL all-gather
    ms_k = time_steps(lambda: G.recombine_gather(*args0()), max(5, min(steps, 50)), 3, stream, world, dist, torch)
    res["fused_recombine_gather_ms"] = ms_k
    res["recv_gbs_per_rank"] = 64 * n * (world - 1) / (ms_k * 1e-3) / 1e9
    if want_nccl:
        ms_n = time_steps(step_nccl, max(5, min(steps, 50)), 3, stream, world, dist, torch)
        verify("nccl")
        res["ms_per_step_nccl_allgather"] = ms_n
        res["recombine_plus_nccl_allgather_ms"] = time_steps(lambda: G.recombine_then_nccl(*args0()), max(5, min(steps, 50)), 3, stream, world, dist, torch)
    G.close()
    return res


def measure_supplementary(args, rank, world, local_rank, workload, log2_batch, field, steps, warmup, unfused_sum=False, cpu_baseline=True):
    """configs[2] (AuthenticatedPoint scalar-muls) and configs[3] (inner product = batch_mul + Sum + open_authenticated pieces);
    one process per GPU, index-range sharding.  Returns the bench object (rank 0) or None."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from ark_mpc_b200 import sharding as sh
    from ark_mpc_b200.engine import Engine

    points = workload == "point_mul"
    n = 1 << log2_batch
    fid = {"bn254_fr": 0, "curve25519_fr": 1}[field]
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        E = Engine(local_rank, field)
        # one MAC key for the whole sharded batch (the same on every rank); everything else is generated per shard
        D = BeaverData(E, n, shard_seed(0xC0FFEE, rank), 0xC0FFEE + 900)
        keys, key, X, Y, A, B, Cc, xv, yv = D.keys, D.key, D.X, D.Y, D.A, D.B, D.C, D.xv, D.yv
        if points:
            # P = y*G shared in the exponent: PointShares (share*G, mac*G)
            Pt = [E.pt_mul_generator(Y[p]) for p in (0, 1)]
            masks = [(E.empty(n), E.empty_points(n)) for _ in (0, 1)]
            outs = [E.empty_points(n, share=True) for _ in (0, 1)]

            def step():
                for p in (0, 1):
                    E.pt_beaver_mask(X[p][0], Pt[p], A[p][0], B[p][0], out=masks[p])
                for p in (0, 1):
                    E.pt_beaver_recombine(p, keys[p], masks[p][0], masks[1 - p][0], masks[p][1], masks[1 - p][1], A[p], B[p], Cc[p], out=outs[p])

            step()
            # correctness gate: the outputs open to (x*y)*G and the MAC shares to key*(x*y)*G, whole batch, affine form
            xy = E.mul(xv, yv)
            opened = E.pt_normalize(E.pt_add(outs[0], outs[1])).reshape(n, 2, 8)
            ok = torch.equal(opened[:, 0, :], E.pt_normalize(E.pt_mul_generator_public(xy))) and \
                torch.equal(opened[:, 1, :], E.pt_normalize(E.pt_mul_generator_public(E.scale(xy, key))))
            del xy, opened
            pw = E.point_words * 8
            alg_bytes = 2 * ((3 * 32 + 2 * pw + 32 + pw) + (2 * 32 + 2 * pw + 6 * 32 + 2 * pw))  # both parties, K1 + K2
            metric, unit = "authenticated_point_mults_per_sec", "point mults/s"
            wl = f"2^{log2_batch} AuthenticatedPoint scalar-muls over {E.field_name.replace('_fr', '')} per GPU, both parties, mock net (BASELINE.json configs[2])"
        else:
            de = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
            outs = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
            sums = [None, None]

            def step():
                for p in (0, 1):
                    E.beaver_mask(X[p][0], Y[p][0], A[p][0], B[p][0], out=de[p])
                for p in (0, 1):
                    if unfused_sum:
                        E.beaver_recombine(p, keys[p], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], A[p], B[p], Cc[p], out=outs[p])
                        sums[p] = E.share_sum(outs[p])
                    else:  # second Beaver phase and the tree-sum in one kernel: the products are never written
                        sums[p] = E.beaver_recombine_sum(p, keys[p], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], A[p], B[p], Cc[p])
                for p in (0, 1):
                    if world > 1:  # cross-GPU sum: all-gather of the per-rank partial ScalarShares (64 B), modular add locally
                        sums[p] = sh.all_reduce_share_sum(E, sums[p])
                opened = E.add(sums[0][0], sums[1][0])            # open of the single result
                chk = [E.mac_check(keys[p], opened, sums[p][1]) for p in (0, 1)]  # MAC-check shares (commitment hash is host-side)
                return opened, chk

            opened, chk = step()
            want = E.sum(E.mul(xv, yv))
            if world > 1:
                want = E.sum(sh.all_gather_rows(want))
            ok = torch.equal(opened, want) and E.sum_is_zero(chk[0], chk[1])
            alg_bytes = 2 * (192 + 384 + 64) if unfused_sum else 2 * (192 + 320)
            metric, unit = "inner_product_elements_per_sec", "elements/s"
            wl = f"secret-shared inner product of length-2^{log2_batch} vectors per GPU (batch_mul + tree-sum + open with MAC check), both parties (BASELINE.json configs[3])"
        if not ok:
            raise SystemExit(f"{workload}: correctness gate failed")
        sampler.wait_first()
        t_begin = time.perf_counter()
        l0 = E.launches
        ms = time_steps(step, steps, warmup, stream, world, dist, torch)
        t_end = time.perf_counter()
        launches = (E.launches - l0) * steps // (steps + warmup)
        clocks = sampler.stop(t_begin, t_end)
        line = None
        if rank == 0:
            peak, peak_src = load_peaks()
            achieved = alg_bytes * n / (ms * 1e-3) / 1e9
            line = {"metric": metric, "value": n * world / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "u32", "dtype_note": "8 x u32 limbs; Montgomery for BN254, special-form 2^255-19 for Curve25519", "data": "synthetic",
                    "config": {"workload": wl, "field": field, "log2_batch_per_gpu": log2_batch, "parties": 2,
                               "sharding": f"index-range x{world}, no data-path collective" if points or world == 1 else
                                           f"index-range x{world}; per step four 64-byte all-gathers of the partial ScalarShares (modular sums cannot use ncclSum)",
                               "l2_hygiene": "inputs larger than L2" if n * 64 > (126 << 20) else "step touches more than L2 in total"},
                    "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                 "traffic": None, "peak_source": peak_src,
                                 "note": ("compute-bound (INT32 multiply pipe): ~6.7k base-field multiplications per party-gate; the HBM fraction is "
                                          "reported for completeness") if points else ("HBM-bound streaming: 640 algorithmic B per party-element" if unfused_sum else
                                                                           "HBM-bound streaming: 512 algorithmic B per party-element (fused recombine + sum)")},
                    "clocks": clocks, "gpu_launches": int(launches) * world}
            if cpu_baseline:
                from oracle import coracle as co
                from tests.util import aos

                cores = os.cpu_count() or 1
                m = min(n, 4096 if points else 1 << 18)
                torch.cuda.synchronize()
                cut = lambda pl: aos(E.download(pl[0][:m].contiguous()), E.download(pl[1][:m].contiguous()))
                x0, x1, y0, y1, a0, a1, b0, b1, c0, c1 = X[0], X[1], Y[0], Y[1], A[0], A[1], B[0], B[1], Cc[0], Cc[1]
                if points:
                    cv = fid
                    Ph = tuple(E.download(Pt[p][:m].contiguous()) for p in (0, 1))
                    ins = (keys, (cut(x0), cut(x1)), Ph, (cut(a0), cut(a1)), (cut(b0), cut(b1)), (cut(c0), cut(c1)))
                    co.two_party_point_mul(cv, cores, *ins, want_open=False)
                    t0 = time.perf_counter()
                    o0, o1, _, _ = co.two_party_point_mul(cv, cores, *ins, want_open=False)
                    dt = time.perf_counter() - t0
                    # parity on the benchmark inputs: a prefix AND a strided sample over the whole batch, both parties, affine form
                    got = [E.download(E.pt_normalize(outs[p][:m].contiguous())) for p in (0, 1)]
                    same = all(np.array_equal(co.pt_normalize(cv, o.reshape(2 * m, -1)), g) for o, g in zip((o0, o1), got))
                    stride = max(1, n // 512)
                    pick = lambda pl: (pl[0][::stride].contiguous(), pl[1][::stride].contiguous())
                    cuts = lambda pl: aos(E.download(pick(pl)[0]), E.download(pick(pl)[1]))
                    Ps = tuple(E.download(Pt[p][::stride].contiguous()) for p in (0, 1))
                    s0, s1, _, _ = co.two_party_point_mul(cv, cores, keys, (cuts(x0), cuts(x1)), Ps, (cuts(a0), cuts(a1)), (cuts(b0), cuts(b1)),
                                                          (cuts(c0), cuts(c1)), want_open=False)
                    gs = [E.download(E.pt_normalize(outs[p][::stride].contiguous())) for p in (0, 1)]
                    same = same and all(np.array_equal(co.pt_normalize(cv, o.reshape(2 * o.shape[0], -1)), g) for o, g in zip((s0, s1), gs))
                    line["parity"] = f"both parties' outputs equal the CPU oracle on the first {m} gates and on every {stride}th gate of the batch (affine form)"
                    if not same:
                        raise SystemExit("point_mul: GPU result differs from the CPU oracle on the benchmark inputs")
                else:
                    ins = (keys, (cut(x0), cut(x1)), (cut(y0), cut(y1)), (cut(a0), cut(a1)), (cut(b0), cut(b1)), (cut(c0), cut(c1)))
                    co.two_party_batch_mul(fid, cores, *ins, want_open=False)
                    t0 = time.perf_counter()
                    o0, o1, _, _ = co.two_party_batch_mul(fid, cores, *ins, want_open=False)
                    co.share_sum(fid, o0), co.share_sum(fid, o1)
                    dt = time.perf_counter() - t0
                line["cpu_baseline"] = {"value": m / dt, "unit": unit, "cores": cores, "kind": "port",
                                        "sample": f"the first {m} elements of the same batch, unfused reference gate sequence (oracle/ark_oracle.c), all host threads"}
        E.close()
    return line


def measure_config0(local_rank, iters=20):
    """BASELINE.json configs[0] — the reference's own bench shape (benches/batch_ops.rs:20-40): n = 1024 random scalars over
    Curve25519 Fr, both vectors shared by party 0, AuthenticatedScalarResult::batch_mul, open_authenticated_batch, results on the
    host; PartyIDBeaverSource triples; in-memory mock network; the clock starts inside each party's closure and the slower
    party counts.  GPU arm = ark_mpc_b200.fabric (launch-latency bound at this size); CPU arm = the C restatement of the same
    gate sequence (arithmetic only)."""
    import random

    import numpy as np

    from ark_mpc_b200 import fabric as fb
    from ark_mpc_b200 import fields as fl

    n, field = 1024, "curve25519_fr"
    p = fl.MODULUS[field]
    rng = random.Random(1024)
    a = [rng.randrange(p) for _ in range(n)]
    b = [rng.randrange(p) for _ in range(n)]
    want = [x * y % p for x, y in zip(a, b)]
    a_l, b_l = fl.mont_limbs_batch(field, a), fl.mont_limbs_batch(field, b)

    def party(fabric):
        t0 = time.perf_counter()
        me = fabric.party_id()
        sa = fabric.batch_share_scalar(a_l if me == 0 else n, 0)
        sb = fabric.batch_share_scalar(b_l if me == 0 else n, 0)
        res = fb.AuthenticatedScalarResult.batch_mul(sa, sb)
        opened = fb.AuthenticatedScalarResult.open_authenticated_batch(res)
        vals = opened.result().to_limbs()  # host copy = "await all"
        return time.perf_counter() - t0, vals

    times = []
    for it in range(iters + 3):
        (t0, v0), (t1, v1) = fb.execute_mock_mpc(party, field=field, device=local_rank)
        if it == 0:
            if fl.from_mont_batch(field, v0) != want:
                raise SystemExit("configs[0]: opened products differ from a*b")
            if not np.array_equal(v0, v1):
                raise SystemExit("configs[0]: the parties opened different values")
        if it >= 3:
            times.append(max(t0, t1))
    times.sort()
    med = times[len(times) // 2]
    py = {"value": n / med, "ms_per_iter": 1e3 * med, "host": "ark_mpc_b200/fabric.py (Python mirror, two threads)"}
    # the same flow through the compiled C++ host mirror (host/arkmpc_host.hpp over the C ABI): what a Rust host's latency looks like
    cpp = None
    exe = os.path.join(ROOT, "tools", "host_bench", "bench_config0")
    if os.path.exists(exe):
        try:
            r = subprocess.run([exe, str(n), "30", "1"], capture_output=True, text=True, timeout=300)
            j = json.loads(r.stdout.strip().splitlines()[-1])
            if "error" not in j:
                cpp = {"value": j["mults_per_s"], "ms_per_iter": 1e3 * j["median_s"], "ms_per_iter_min": 1e3 * j["min_s"], "iters": j["iters"],
                       "host": "tools/host_bench/bench_config0.cpp over host/arkmpc_host.hpp (compiled C++ mirror, two threads)"}
                # the same flow at 64 x the batch: where the fixed per-iteration host latency stops dominating
                r = subprocess.run([exe, str(64 * n), "10", "1"], capture_output=True, text=True, timeout=300)
                j = json.loads(r.stdout.strip().splitlines()[-1])
                if "error" not in j:
                    cpp["at_batch_%d" % (64 * n)] = {"value": j["mults_per_s"], "ms_per_iter": 1e3 * j["median_s"]}
        except Exception as ex:  # noqa: BLE001
            cpp = {"error": repr(ex)[:200]}
    best = cpp if cpp and "value" in cpp else py
    med = best["ms_per_iter"] * 1e-3
    out = {"metric": METRIC, "value": n / med, "unit": UNIT, "ms_per_iter": 1e3 * med, "iters": iters, "python_mirror": py, "cpp_host_mirror": cpp,
           "config": {"workload": "1024-element share + AuthenticatedScalar batch_mul + open_authenticated over Curve25519 Fr, 2 parties in one process, "
                                  "in-memory mock net, PartyIDBeaverSource (BASELINE.json configs[0]; benches/batch_ops.rs:20-40)",
                      "field": field, "batch": n, "parties": 2},
           "note": "host wall clock from inside each party's closure to the opened values on the host, slower party, median; `value` is the compiled "
                   "C++ host mirror when its binary is present (python_mirror beside it); at n = 1024 the GPU path is bound by ~25 kernel launches, "
                   "8 message hand-offs, 30 device buffers and two host SHA3 commitments per party (host profile: ARKMPC_HOST_PROFILE=1 "
                   "tools/host_bench/bench_config0; profiles/r02p_config0_profile.txt), not by arithmetic",
           "roofline": None}
    # CPU arm: the same gate sequence restated in C, both parties, arithmetic only (no executor, no commitment hash)
    from oracle import coracle as co
    from tests.util import TwoPartyData, aos

    fid = 1
    D = TwoPartyData(fid, n, seed=1024)
    g = lambda t: (aos(*t[0]), aos(*t[1]))
    ins = (D.keys, g(D.x), g(D.y), g(D.a), g(D.b), g(D.c))
    v = D.x[0][0]  # a public (n,4) plane for the input-sharing add_public

    import hashlib

    blinder = bytes(32)

    def cpu_iter():
        for sh in (ins[1], ins[2]):  # batch_share_scalar: mask shares + add_public of the masked value, both parties
            for pid in (0, 1):
                co.batch_add_public(fid, pid, D.keys[pid], sh[pid], v)
        o0, o1, _, _ = co.two_party_batch_mul(fid, 1, *ins, want_open=False)
        opened = co.scalar_add(fid, np.ascontiguousarray(o0[:, :4]), np.ascontiguousarray(o1[:, :4]))
        checks = [co.mac_check(fid, D.keys[pid], opened, o) for pid, o in ((0, o0), (1, o1))]
        for pid in (0, 1):  # HashCommitment of the own MAC-check vector, then of the peer's to verify it (commitment.rs:63-89)
            for who in (pid, 1 - pid):
                hashlib.sha3_256(np.ascontiguousarray(checks[who]).tobytes() + blinder).digest()

    cpu_iter()
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        cpu_iter()
    dt = (time.perf_counter() - t0) / reps
    out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
                           "sample": f"{reps} iterations of the same 1024-gate flow (input sharing add_public x2, unfused batch_mul, open, MAC-check vector, "
                                     "the two SHA3 commitments per party through hashlib), both parties on ONE thread (the reference runs them on two); "
                                     "omits the reference executor and its per-element result bookkeeping, so it flatters the reference"}
    return out


_JSON_FD = None


def claim_stdout():
    """Keep stdout for the one JSON line: libraries that write to fd 1 (NCCL prints its version banner there) go to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_json(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-batch", type=int, default=20)
    ap.add_argument("--log2-total", type=int, default=24, help="configs[4]: total gates sharded over the N GPUs (N > 1)")
    ap.add_argument("--field", default="bn254_fr", choices=["bn254_fr", "curve25519_fr"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hint", action="store_true", help="A/B: launch every kernel ordered after its predecessor (no independence hints)")
    ap.add_argument("--hint-level", type=int, default=2, choices=[0, 1, 2], help="0 none; 1 the second launch of each phase; 2 also the next step's masks")
    ap.add_argument("--unfused-sum", action="store_true", help="inner_product: separate recombine and share_sum launches (A/B)")
    ap.add_argument("--configs", default="all", help="'all' (default), 'none', or a comma list of 0,2,3,4: which BASELINE.json configs ride along "
                                                     "under the `configs` key of the headline line")
    ap.add_argument("--workload", default="beaver_fr", choices=["beaver_fr", "point_mul", "inner_product", "config0"],
                    help="beaver_fr = BASELINE.json's metric (configs[1]) with the other configs under `configs`; point_mul = configs[2]; "
                         "inner_product = configs[3]; config0 = configs[0] (single-config lines of the same JSON shape)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from ark_mpc_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def finish(line):
        if rank == 0 and line is not None:
            emit_json(line)
        if world > 1:
            dist.destroy_process_group()

    given = lambda flag: flag in sys.argv
    if args.workload in ("point_mul", "inner_product"):
        points = args.workload == "point_mul"
        field = args.field if not points or args.field != "bn254_fr" or given("--field") else "curve25519_fr"
        log2 = args.log2_batch if given("--log2-batch") else (20 if points else 22)
        steps = args.steps if given("--steps") else (10 if points else 200)
        warmup = max(3, args.warmup if given("--warmup") else 3)
        return finish(measure_supplementary(args, rank, world, local_rank, args.workload, log2, field, steps, warmup, args.unfused_sum,
                                            cpu_baseline=not args.no_cpu_baseline and world == 1))
    if args.workload == "config0":
        return finish(measure_config0(local_rank) if rank == 0 else None)

    which = {"all": {0, 2, 3, 4}, "none": set()}.get(args.configs)
    if which is None:
        which = {int(c) for c in args.configs.split(",") if c.strip()}

    fid = {"bn254_fr": 0, "curve25519_fr": 1}[args.field]
    n = 1 << args.log2_batch
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.Stream()
    open_gather = None
    config4 = None
    with torch.cuda.stream(stream):
        E = Engine(local_rank, args.field)  # binds the current (bench) stream
        base_seed = 0xA11CE
        D = BeaverData(E, n, shard_seed(base_seed, rank), shard_seed(base_seed, rank) + 900)
        P, key, key0, key1, xv, yv = D.P, D.key, D.key0, D.key1, D.xv, D.yv
        x0, x1, y0, y1, a0, a1, b0, b1, c0, c1 = D.X[0], D.X[1], D.Y[0], D.Y[1], D.A[0], D.A[1], D.B[0], D.B[1], D.C[0], D.C[1]
        # two sets of d/e planes, used by alternate steps: like an executor that gives every gate fresh output buffers, so that the
        # masks of step i+1 do not overwrite what the recombines of step i read (no write-after-read edge between steps)
        de_sets = [[(E.empty(n), E.empty(n)) for _ in range(2)] for _ in range(2)]
        de = de_sets[0]
        out = [(E.empty(n), E.empty(n)) for _ in range(2)]

        # Independence hints (arkmpc_ctx_hint_independent).  Level 1: the second launch of each phase (the other party's
        # kernel: different operands, different outputs) overlaps the first's drain.  Level 2: with the alternating d/e sets the masks of
        # the NEXT step are independent of this step's recombines too, so only recombine(p0) — which reads what the masks just wrote —
        # is ordered after its predecessors.
        level = 0 if args.no_hint else args.hint_level
        hint = E.hint_independent

        def mask_both(i=0):
            d = de_sets[i & 1]
            if level >= 2 and i > 0:
                hint()
            E.beaver_mask(P[0]["x"][0], P[0]["y"][0], P[0]["a"][0], P[0]["b"][0], out=d[0])
            if level >= 1:
                hint()
            E.beaver_mask(P[1]["x"][0], P[1]["y"][0], P[1]["a"][0], P[1]["b"][0], out=d[1])

        def recombine_both(i=0):
            d = de_sets[i & 1]
            E.beaver_recombine(0, P[0]["key"], d[0][0], d[0][1], d[1][0], d[1][1], P[0]["a"], P[0]["b"], P[0]["c"], out=out[0])
            if level >= 1:
                hint()
            E.beaver_recombine(1, P[1]["key"], d[1][0], d[1][1], d[0][0], d[0][1], P[1]["a"], P[1]["b"], P[1]["c"], out=out[1])

        def step(i=0):
            mask_both(i)
            recombine_both(i)

        # correctness gate before timing: opened product == x*y, MAC shares sum to key*x*y (whole batch, on device)
        step()
        xy = E.mul(xv, yv)
        ok = torch.equal(E.add(out[0][0], out[1][0]), xy) and torch.equal(E.add(out[0][1], out[1][1]), E.scale(xy, key))
        if not ok:
            raise SystemExit("correctness gate failed: opened product != x*y")
        del xy

        for w in range(args.warmup):
            step(w)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.wait_first()
        t_begin = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # K2 is timed inside the timed region on every 8th step: an event record between two launches takes away the
        # programmatic-dependent-launch overlap of that one edge, so sampling keeps the region representative
        sample_every = 8
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range((args.steps + sample_every - 1) // sample_every)]
        launches0 = E.launches
        ev0.record(stream)
        for i in range(args.steps):
            mask_both(i)
            if i % sample_every == 0:
                kev[i // sample_every][0].record(stream)
                recombine_both(i)
                kev[i // sample_every][1].record(stream)
            else:
                recombine_both(i)
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        launches = E.launches - launches0
        ms = ev0.elapsed_time(ev1)
        k2_ms = sum(a.elapsed_time(b) for a, b in kev) / (2 * len(kev))  # per recombine launch (sampled steps)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        clocks = sampler.stop(t_begin, t_end)

        # ---- e2e: host AoS buffers through the C ABI, copies inside the timed region ----
        # One host thread and one native context per party (the reference runs the parties as two tasks, benches/batch_ops.rs:20-40):
        # begin (upload x,y,a,b,c; mask; download d||e) -> exchange by pointer -> finish (upload the peer's d||e; recombine; download).
        e2e = None
        if args.e2e_steps > 0:
            E1 = Engine(local_rank, args.field)
            engines = (E, E1)
            host = []
            for p in (0, 1):
                hp = {}
                for nm in ("x", "y", "a", "b", "c"):
                    buf = engines[p].pinned_empty((n, 8))
                    buf[:] = E.download(E.share_zip(P[p][nm]))
                    hp[nm] = buf
                hp["de"] = engines[p].pinned_empty((2 * n, 4))
                hp["out"] = engines[p].pinned_empty((n, 8))
                host.append(hp)
            gate = threading.Barrier(2)
            errs = []

            share_operands = [False]  # second pass: x, y as planes of their share halves (arkmpc_fr_batch_mul_begin_host_shares)

            def party_steps(p, reps):
                try:
                    torch.cuda.set_device(local_rank)
                    Ep = engines[p]
                    for _ in range(reps):
                        if share_operands[0]:
                            sess = Ep.batch_mul_begin_host_shares(p, P[p]["key"], host[p]["xs"], host[p]["ys"], host[p]["a"], host[p]["b"],
                                                                  host[p]["c"], host[p]["de"])
                        else:
                            sess = Ep.batch_mul_begin_host(p, P[p]["key"], host[p]["x"], host[p]["y"], host[p]["a"], host[p]["b"], host[p]["c"],
                                                           host[p]["de"])
                        gate.wait()  # both parties' d||e are in host memory: the mock network hands over the pointer
                        Ep.batch_mul_finish_host(sess, host[1 - p]["de"], host[p]["out"])
                        gate.wait()
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)
                    gate.abort()

            def e2e_run(reps):
                th = [threading.Thread(target=party_steps, args=(p, reps)) for p in (0, 1)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()
                if errs:
                    raise errs[0]

            e2e_run(1)
            ref = E.download(E.share_zip(out[0]))
            if not np.array_equal(host[0]["out"], ref):
                raise SystemExit("e2e path result differs from the device-resident path")
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_run(args.e2e_steps)
            torch.cuda.synchronize()
            e2e_ms = 1e3 * (time.perf_counter() - t0) / args.e2e_steps
            if world > 1:
                t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_ms = float(t.item())
            h2d, d2h = E.host_path_bytes(n)
            e2e = {"value": n * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": 2 * h2d, "d2h_bytes_per_step": 2 * d2h,
                   "h2d_gbs_aggregate": 2 * h2d * world / (e2e_ms * 1e-3) / 1e9, "d2h_gbs_aggregate": 2 * d2h * world / (e2e_ms * 1e-3) / 1e9,
                   "api": "arkmpc_fr_batch_mul_begin_host / arkmpc_fr_batch_mul_finish_host (pinned host AoS buffers, one host thread per party)",
                   "host_numa": numa}
            # The same step with the operands given as what the gate reads: x, y as planes of their share halves (the operands' MACs
            # are not inputs of a Beaver multiplication); a, b, c, d || e and the result as before.  Reported beside `e2e`, which keeps
            # the reference's Vec<ScalarShare> images for every operand.
            for p in (0, 1):
                for nm in ("x", "y"):
                    buf = engines[p].pinned_empty((n, 4))
                    buf[:] = host[p][nm][:, :4]
                    host[p][nm + "s"] = buf
                host[p]["out"][:] = 0
            share_operands[0] = True
            e2e_run(1)
            if not np.array_equal(host[0]["out"], ref):
                raise SystemExit("e2e path (share-plane operands) result differs from the device-resident path")
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_run(args.e2e_steps)
            torch.cuda.synchronize()
            so_ms = 1e3 * (time.perf_counter() - t0) / args.e2e_steps
            if world > 1:
                t = torch.tensor([so_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                so_ms = float(t.item())
            so_h2d = h2d - 64 * n
            e2e["share_plane_operands"] = {"value": n * world / (so_ms * 1e-3), "unit": UNIT, "ms_per_step": so_ms, "h2d_bytes_per_step": 2 * so_h2d,
                                           "d2h_bytes_per_step": 2 * d2h, "h2d_gbs_aggregate": 2 * so_h2d * world / (so_ms * 1e-3) / 1e9,
                                           "api": "arkmpc_fr_batch_mul_begin_host_shares / arkmpc_fr_batch_mul_finish_host (x, y: n x 32-byte share "
                                                  "planes; a, b, c: AoS images)"}
            E1.close()
            del host

        # ---- N > 1: the step with the batch_open all-gather inside it (north_star's one collective) ----
        if world > 1:
            try:
                gsteps = max(20, min(args.steps, 200))
                open_gather = measure_open_gather(E, stream, D, n, rank, world, gsteps, 5, base_seed)
                open_gather["value_with_open_gather"] = n * world / (open_gather["ms_per_step"] * 1e-3)
            except SystemExit:
                raise
            except Exception as ex:  # e.g. CUDA IPC unavailable in a restricted container: the headline numbers do not depend on it
                open_gather = {"error": repr(ex)[:300]}
        del de, de_sets, out

        # ---- configs[4]: 2^log2_total gates sharded over the N GPUs, gather inside the step ----
        if world > 1 and 4 in which and args.field == "bn254_fr":
            try:
                n4 = (1 << args.log2_total) // world
                if n4 == n:
                    D4 = D
                else:
                    del D, P
                    D4 = BeaverData(E, n4, shard_seed(base_seed, rank), shard_seed(base_seed, rank) + 900)
                s4 = max(10, min(args.steps, 100))
                g4 = measure_open_gather(E, stream, D4, n4, rank, world, s4, 3, base_seed)
                peak, _ = load_peaks()
                config4 = {"metric": METRIC, "value": n4 * world / (g4["ms_per_step"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": s4, "warmup": 3,
                           "ms_per_step": g4["ms_per_step"], "scaling": "strong",
                           "config": {"workload": f"2^{args.log2_total} authenticated scalar Beaver muls over bn254_fr sharded over {world} GPUs "
                                                  f"(2^{n4.bit_length() - 1} per GPU), both parties, mock net, party 0's opened d || e all-gathered onto every "
                                                  "rank inside the step (BASELINE.json configs[4])",
                                      "field": "bn254_fr", "log2_total": args.log2_total, "parties": 2,
                                      "sharding": f"index-range x{world}; one all-gather per step, fused into the recombine kernel's stores"},
                           "open_gather": g4,
                           "roofline": {"bound": "nvlink", "kernel": "beaver_recombine_gather (party 0)", "achieved": g4["recv_gbs_per_rank"], "peak": 900.0,
                                        "unit": "GB/s", "frac": g4["recv_gbs_per_rank"] / 900.0,
                                        "note": "bytes received per rank (64 B x rows of the other ranks) over the fused kernel's duration vs NVLink 5's 900 GB/s per "
                                                "direction (nominal, task statement); the compute phase of the same step is HBM-bound like configs[1]",
                                        "step_hbm_gbs": (2 * (BYTES_MASK + BYTES_RECOMBINE) + 64) * n4 / (g4["ms_per_step"] * 1e-3) / 1e9, "hbm_peak": peak}}
                del D4
            except SystemExit:
                raise
            except Exception as ex:  # noqa: BLE001
                config4 = {"error": repr(ex)[:300]}

        cpu_line = None
        if rank == 0 and not args.no_cpu_baseline and world == 1:
            from oracle import coracle as co
            from tests.util import aos

            cores = os.cpu_count() or 1
            ha = lambda pl: aos(E.download(pl[0]), E.download(pl[1]))
            ins = ((key0, key1), (ha(x0), ha(x1)), (ha(y0), ha(y1)), (ha(a0), ha(a1)), (ha(b0), ha(b1)), (ha(c0), ha(c1)))
            o0, _, _, _ = co.two_party_batch_mul(fid, cores, *ins, want_open=False)  # warm-up + parity check of the timed data
            E.beaver_mask(x0[0], y0[0], a0[0], b0[0], out=(de0 := (E.empty(n), E.empty(n))))
            E.beaver_mask(x1[0], y1[0], a1[0], b1[0], out=(de1 := (E.empty(n), E.empty(n))))
            (gs, gm), _ = E.beaver_recombine(0, key0, de0[0], de0[1], de1[0], de1[1], a0, b0, c0)
            if not np.array_equal(o0, aos(E.download(gs), E.download(gm))):
                raise SystemExit("GPU result differs from the CPU oracle on the benchmark inputs")
            t0 = time.perf_counter()
            for _ in range(args.cpu_steps):
                co.two_party_batch_mul(fid, cores, *ins, want_open=False)
            dt = (time.perf_counter() - t0) / args.cpu_steps
            t0 = time.perf_counter()
            co.two_party_batch_mul(fid, 1, *ins, want_open=False)
            dt1 = time.perf_counter() - t0
            cpu_line = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"the same 2^{args.log2_batch} two-party batch, {args.cpu_steps} reps, all host threads",
                        "single_thread_value": n / dt1}
            del ins, o0
        sm_count = E.sm_count
        E.close()

    # ---- the other BASELINE.json configs, each a bench object of its own under `configs` ----
    cfgs = [None] * 5
    cfgs[1] = {"see": "this line (headline): value, e2e, roofline, cpu_baseline"}
    sup_cpu = not args.no_cpu_baseline and world == 1
    if 0 in which:
        cfgs[0] = (measure_config0(local_rank) if rank == 0 else None) if world == 1 else {"see": "the N = 1 line (a 1024-gate batch does not shard)"}
        if world > 1:
            dist.barrier()
    if 2 in which:
        cfgs[2] = measure_supplementary(args, rank, world, local_rank, "point_mul", 20, "curve25519_fr", 10, 3, cpu_baseline=sup_cpu)
        if cfgs[2] is not None and world == 1:
            bn = measure_supplementary(args, rank, world, local_rank, "point_mul", 20, "bn254_fr", 6, 3, cpu_baseline=False)
            cfgs[2]["bn254_g1"] = {k: bn[k] for k in ("value", "unit", "ms_per_step", "steps")}
    if 3 in which:
        cfgs[3] = measure_supplementary(args, rank, world, local_rank, "inner_product", 22, "bn254_fr", 100, 3, cpu_baseline=sup_cpu)
    cfgs[4] = config4 if config4 is not None else {"see": "needs N > 1 (SCALE lines carry it): 2^24 gates sharded over the GPUs with the open all-gather in the step"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    traffic = load_traffic() if (args.field == "bn254_fr" and args.log2_batch == 20) else None
    ms_per_step = ms / args.steps
    value = n * world / (ms_per_step * 1e-3)
    achieved = BYTES_RECOMBINE * n / (k2_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "dtype_note": "256-bit field elements as 8 x u32 limbs, Montgomery form, IMAD.WIDE carry chains",
        "data": "synthetic",
        "config": workload_config(args, world),
        "party_gates_per_sec": 2 * value,
        "step_hbm_gbs": (2 * (BYTES_MASK + BYTES_RECOMBINE) * n) / (ms_per_step * 1e-3) / 1e9,
        "roofline": {"bound": "hbm", "kernel": "beaver_recombine_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic[0] if traffic else None,
                     "traffic_source": traffic[1] if traffic else None, "peak_source": peak_src, "kernel_us": 1e3 * k2_ms,
                     "algorithmic_bytes_per_launch": BYTES_RECOMBINE * n,
                     "modmul_equiv_per_sec": 6 * n / (k2_ms * 1e-3),
                     # the second bound of this kernel: wide multiply-adds per gate at the measured issue rate of
                     # 31.5 lanes/clk/SM (tools/pipe_bench.cu, profiles/r01a_pipe_bench.txt) at the sampled SM clock
                     "int_pipe_floor_us": (IMAD_WIDE_PER_GATE * n / (31.5 * sm_count * (clocks.get("sm_mhz") or 1965.0) * 1e6)) * 1e6,
                     "hbm_floor_us": BYTES_RECOMBINE * n / (peak * 1e9) * 1e6},
        "clocks": clocks, "gpu_launches": int(launches) * world,
    }
    if e2e:
        line["e2e"] = e2e
    if open_gather:
        if "value_with_open_gather" in open_gather:
            line["value_with_open_gather"] = open_gather.pop("value_with_open_gather")
        line["open_gather"] = open_gather
    if cpu_line:
        line["cpu_baseline"] = cpu_line
    line["configs"] = cfgs
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
