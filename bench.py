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

N > 1: one process per GPU (torchrun), the batch is sharded by index range with NO data-path
collective (every gate is element-wise); weak scaling, 2^log2_batch gates per GPU.
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
            if m and cur and kernel_substr in cur:
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
            "l2_hygiene": "inputs larger than L2: ~0.9 GB touched per step vs 126 MB L2"}


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU e2e: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, so that the pinned
    staging buffers (first touch) and the H2D/D2H traffic stay on the local socket.  Best effort; returns the node or None."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(f"{base}/numa_node").read().strip())
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_supplementary(args, rank, world, local_rank):
    """configs[2] (2^20 AuthenticatedPoint scalar-muls) and configs[3] (inner product = batch_mul + Sum + open_authenticated pieces)
    as bench lines of the same shape; one process per GPU, index-range sharding, no data-path collective."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from ark_mpc_b200 import sharding as sh
    from ark_mpc_b200.engine import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    points = args.workload == "point_mul"
    field = args.field if not points or args.field != "bn254_fr" or "--field" in sys.argv else "curve25519_fr"
    if points and "--log2-batch" not in sys.argv:
        args.log2_batch = 20
    if not points and "--log2-batch" not in sys.argv:
        args.log2_batch = 22
    n = 1 << args.log2_batch
    steps = args.steps if "--steps" in sys.argv else (10 if points else 200)
    warmup = max(3, args.warmup if "--warmup" in sys.argv else 3)
    fid = {"bn254_fr": 0, "curve25519_fr": 1}[field]
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        E = Engine(local_rank, field)
        seed = 0xC0FFEE + 7919 * rank
        # one MAC key for the whole sharded batch (the same on every rank); everything else is generated per shard
        key0, key1 = (E.download(E.random(0xC0FFEE + 900 + p, 0, 1))[0].copy() for p in (0, 1))
        key = E.download(E.add(E.upload(key0.reshape(1, 4)), E.upload(key1.reshape(1, 4))))[0].copy()
        keys = (key0, key1)

        def shared(s, val=None):
            v = E.random(s, 0, n) if val is None else val
            s0, m0 = E.random(s + 1, 0, n), E.random(s + 2, 0, n)
            return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

        xv, x0, x1 = shared(seed + 10)
        yv, y0, y1 = shared(seed + 20)
        av, a0, a1 = shared(seed + 30)
        bv, b0, b1 = shared(seed + 40)
        _, c0, c1 = shared(seed + 50, E.mul(av, bv))
        X, Y, A, B, Cc = (x0, x1), (y0, y1), (a0, a1), (b0, b1), (c0, c1)
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
            pw = E.point_words * 8
            alg_bytes = 2 * ((3 * 32 + 2 * pw + 32 + pw) + (2 * 32 + 2 * pw + 6 * 32 + 2 * pw))  # both parties, K1 + K2
            metric, unit = "authenticated_point_mults_per_sec", "point mults/s"
            wl = f"2^{args.log2_batch} AuthenticatedPoint scalar-muls over {E.field_name.replace('_fr', '')} per GPU, both parties, mock net (BASELINE.json configs[2])"
        else:
            de = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
            outs = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
            sums = [None, None]

            def step():
                for p in (0, 1):
                    E.beaver_mask(X[p][0], Y[p][0], A[p][0], B[p][0], out=de[p])
                for p in (0, 1):
                    if args.unfused_sum:
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
            alg_bytes = 2 * (192 + 384 + 64) if args.unfused_sum else 2 * (192 + 320)
            metric, unit = "inner_product_elements_per_sec", "elements/s"
            wl = f"secret-shared inner product of length-2^{args.log2_batch} vectors per GPU (batch_mul + tree-sum + open with MAC check), both parties (BASELINE.json configs[3])"
        if not ok:
            raise SystemExit("correctness gate failed")
        for _ in range(warmup):
            step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.wait_first()
        t_begin = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = E.launches
        ev0.record(stream)
        for _ in range(steps):
            step()
        ev1.record(stream)
        stream.synchronize()
        t_end = time.perf_counter()
        launches = E.launches - l0
        ms = ev0.elapsed_time(ev1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        clocks = sampler.stop(t_begin, t_end)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = load_peaks()
    achieved = alg_bytes * n / (ms * 1e-3) / 1e9
    line = {"metric": metric, "value": n * world / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "dtype_note": "8 x u32 limbs; Montgomery for BN254, special-form 2^255-19 for Curve25519", "data": "synthetic",
            "config": {"workload": wl, "field": field, "log2_batch_per_gpu": args.log2_batch, "parties": 2,
                       "sharding": f"index-range x{world}, no data-path collective",
                       "l2_hygiene": "inputs larger than L2" if n * 64 > (126 << 20) else "step touches more than L2 in total"},
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "note": ("compute-bound (INT32 multiply pipe): ~6.7k base-field multiplications per party-gate; the HBM fraction is "
                                  "reported for completeness") if points else ("HBM-bound streaming: 640 algorithmic B per party-element" if args.unfused_sum else
                                                                   "HBM-bound streaming: 512 algorithmic B per party-element (fused recombine + sum)")},
            "clocks": clocks, "gpu_launches": int(launches) * world}
    if not args.no_cpu_baseline and world == 1:
        from oracle import coracle as co
        from tests.util import aos

        cores = os.cpu_count() or 1
        m = min(n, 4096 if points else 1 << 18)
        torch.cuda.synchronize()
        cut = lambda pl: aos(E.download(pl[0][:m].contiguous()), E.download(pl[1][:m].contiguous()))
        if points:
            cv = fid
            Ph = tuple(E.download(Pt[p][:m].contiguous()) for p in (0, 1))
            ins = (keys, (cut(x0), cut(x1)), Ph, (cut(a0), cut(a1)), (cut(b0), cut(b1)), (cut(c0), cut(c1)))
            co.two_party_point_mul(cv, cores, *ins, want_open=False)
            t0 = time.perf_counter()
            o0, _, _, _ = co.two_party_point_mul(cv, cores, *ins, want_open=False)
            dt = time.perf_counter() - t0
            with torch.cuda.stream(stream):  # the engine launches on the bench stream: download on the same one
                got = E.download(E.pt_normalize(outs[0][:m].contiguous()))
            same = np.array_equal(co.pt_normalize(cv, o0.reshape(2 * m, -1)), got)
            if not same:
                raise SystemExit("GPU result differs from the CPU oracle on the benchmark inputs")
        else:
            ins = (keys, (cut(x0), cut(x1)), (cut(y0), cut(y1)), (cut(a0), cut(a1)), (cut(b0), cut(b1)), (cut(c0), cut(c1)))
            co.two_party_batch_mul(fid, cores, *ins, want_open=False)
            t0 = time.perf_counter()
            o0, o1, _, _ = co.two_party_batch_mul(fid, cores, *ins, want_open=False)
            co.share_sum(fid, o0), co.share_sum(fid, o1)
            dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": m / dt, "unit": unit, "cores": cores, "kind": "port",
                                "sample": f"the first {m} elements of the same batch, unfused reference gate sequence (oracle/ark_oracle.c), all host threads"}
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--field", default="bn254_fr", choices=["bn254_fr", "curve25519_fr"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused-sum", action="store_true", help="inner_product: separate recombine and share_sum launches (A/B)")
    ap.add_argument("--workload", default="beaver_fr", choices=["beaver_fr", "point_mul", "inner_product"],
                    help="beaver_fr = BASELINE.json's metric (configs[1]); point_mul = configs[2]; inner_product = configs[3] "
                         "(supplementary lines, same JSON shape)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    claim_stdout()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "beaver_fr":
        return run_supplementary(args, rank, world, local_rank)

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

    fid = {"bn254_fr": 0, "curve25519_fr": 1}[args.field]
    n = 1 << args.log2_batch
    sampler = ClockSampler(local_rank)
    sampler.start()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        E = Engine(local_rank, args.field)  # binds the current (bench) stream
        seed = 0xA11CE + 7919 * rank

        def rnd_key(s):
            return E.download(E.random(s, 0, 1))[0].copy()

        key0, key1 = rnd_key(seed + 900), rnd_key(seed + 901)
        key = E.download(E.add(E.upload(key0.reshape(1, 4)), E.upload(key1.reshape(1, 4))))[0].copy()

        def shared(s, val=None):
            v = E.random(s, 0, n) if val is None else val
            s0, m0 = E.random(s + 1, 0, n), E.random(s + 2, 0, n)
            return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

        xv, x0, x1 = shared(seed + 10)
        yv, y0, y1 = shared(seed + 20)
        av, a0, a1 = shared(seed + 30)
        bv, b0, b1 = shared(seed + 40)
        _, c0, c1 = shared(seed + 50, E.mul(av, bv))
        P = [dict(key=key0, x=x0, y=y0, a=a0, b=b0, c=c0), dict(key=key1, x=x1, y=y1, a=a1, b=b1, c=c1)]
        de = [(E.empty(n), E.empty(n)) for _ in range(2)]
        out = [(E.empty(n), E.empty(n)) for _ in range(2)]

        def mask_both():
            for p in (0, 1):
                E.beaver_mask(P[p]["x"][0], P[p]["y"][0], P[p]["a"][0], P[p]["b"][0], out=de[p])

        def recombine_both():
            for p in (0, 1):
                E.beaver_recombine(p, P[p]["key"], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], P[p]["a"], P[p]["b"], P[p]["c"], out=out[p])

        def step():
            mask_both()
            recombine_both()

        # correctness gate before timing: opened product == x*y, MAC shares sum to key*x*y (whole batch, on device)
        step()
        xy = E.mul(xv, yv)
        ok = torch.equal(E.add(out[0][0], out[1][0]), xy) and torch.equal(E.add(out[0][1], out[1][1]), E.scale(xy, key))
        if not ok:
            raise SystemExit("correctness gate failed: opened product != x*y")
        del xy

        for _ in range(args.warmup):
            step()
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
            mask_both()
            if i % sample_every == 0:
                kev[i // sample_every][0].record(stream)
                recombine_both()
                kev[i // sample_every][1].record(stream)
            else:
                recombine_both()
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

            def party_steps(p, reps):
                try:
                    torch.cuda.set_device(local_rank)
                    Ep = engines[p]
                    for _ in range(reps):
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
            e2e = {"value": n * world / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": 2 * (5 * 64 + 64) * n, "d2h_bytes_per_step": 2 * (64 + 64) * n,
                   "api": "arkmpc_fr_batch_mul_begin_host / arkmpc_fr_batch_mul_finish_host (pinned host AoS buffers, one host thread per party)",
                   "host_numa_node": numa}
            E1.close()

        # ---- N > 1: the batch_open all-gather (north_star's one collective), NCCL vs fused into the kernel's stores ----
        open_allgather = None
        if world > 1:
          try:
              from ark_mpc_b200 import sharding as sh

              G = sh.OpenGather(E, n)
              modes = {}
              for mode in ("nccl", "fused"):
                  fn = G.recombine_then_nccl if mode == "nccl" else G.recombine_gather
                  call = lambda: fn(0, P[0]["key"], de[0][0], de[0][1], de[1][0], de[1][1], P[0]["a"], P[0]["b"], P[0]["c"], out[0])
                  for _ in range(3):
                      call()
                  torch.cuda.synchronize()
                  dist.barrier()
                  a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                  reps = max(5, min(args.steps, 50))
                  a0.record(stream)
                  for _ in range(reps):
                      call()
                  a1.record(stream)
                  stream.synchronize()
                  t = torch.tensor([a0.elapsed_time(a1) / reps], device="cuda", dtype=torch.float64)
                  dist.all_reduce(t, op=dist.ReduceOp.MAX)
                  modes[mode] = float(t.item())
                  dist.barrier()
              if not (torch.equal(G.my_rows()[0], E.add(de[0][0], de[1][0]))):
                  raise SystemExit("gathered rows differ from the local opened values")
              gathered = 64 * n * world
              open_allgather = {"what": "party 0's fused recombine + all-gather of the opened d||e (64 B x 2^%d rows per rank) onto every rank" % args.log2_batch,
                                "recombine_plus_nccl_allgather_ms": modes["nccl"], "fused_recombine_gather_ms": modes["fused"],
                                "bytes_gathered_per_rank": gathered,
                                "fused_recv_gbs_per_rank": gathered * (world - 1) / world / (modes["fused"] * 1e-3) / 1e9}
              G.close()
          except Exception as ex:  # e.g. CUDA IPC unavailable in a restricted container: the headline numbers do not depend on it
            open_allgather = {"error": repr(ex)[:300]}

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
                     # the second bound of this kernel: 501 IMAD.WIDE per gate at the measured issue rate of
                     # 31.5 lanes/clk/SM (tools/pipe_bench.cu, profiles/r01a_pipe_bench.txt) at the sampled SM clock
                     "int_pipe_floor_us": (501 * n / (31.5 * E.sm_count * (clocks.get("sm_mhz") or 1965.0) * 1e6)) * 1e6,
                     "hbm_floor_us": BYTES_RECOMBINE * n / (peak * 1e9) * 1e6},
        "clocks": clocks, "gpu_launches": int(launches) * world,
    }
    if e2e:
        line["e2e"] = e2e
    if open_allgather:
        line["open_allgather"] = open_allgather
    if not args.no_cpu_baseline and world == 1:
        from oracle import coracle as co
        from tests.util import aos

        cores = os.cpu_count() or 1
        ha = lambda pl: aos(E.download(pl[0]), E.download(pl[1]))
        ins = ((key0, key1), (ha(x0), ha(x1)), (ha(y0), ha(y1)), (ha(a0), ha(a1)), (ha(b0), ha(b1)), (ha(c0), ha(c1)))
        o0, _, _, _ = co.two_party_batch_mul(fid, cores, *ins, want_open=False)  # warm-up + parity check of the timed data
        if not np.array_equal(o0, ha(out[0])):
            raise SystemExit("GPU result differs from the CPU oracle on the benchmark inputs")
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            co.two_party_batch_mul(fid, cores, *ins, want_open=False)
        dt = (time.perf_counter() - t0) / args.cpu_steps
        t0 = time.perf_counter()
        co.two_party_batch_mul(fid, 1, *ins, want_open=False)
        dt1 = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"the same 2^{args.log2_batch} two-party batch, {args.cpu_steps} reps, all host threads",
                                "single_thread_value": n / dt1}
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
