#!/usr/bin/env python
"""Quick device timing of the point kernels (CUDA events on the launching stream), both curves."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ark_mpc_b200.engine import Engine

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
n = 1 << log2n
for field, curve in (("curve25519_fr", "curve25519_edwards"), ("bn254_fr", "bn254_g1")):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        E = Engine(0, field)
        E.bind_curve(curve)
        rnd = lambda seed: E.random(seed, 0, n)
        key = E.download(E.random(77, 0, 1))[0].copy()
        xs, a_s, a_m, b_s, b_m, c_s, c_m = (rnd(i) for i in range(1, 8))
        P = E.pt_mul_generator((rnd(20), rnd(21)))
        pts = E.pt_mul_generator_public(rnd(22))

        def timed(name, fn, reps=3):
            fn()
            s.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(s)
            for _ in range(reps):
                fn()
            ev1.record(s)
            s.synchronize()
            ms = ev0.elapsed_time(ev1) / reps
            print(f"{curve:20s} {name:28s} n=2^{log2n} {ms:9.3f} ms  {n / ms / 1e3:9.3f} M elems/s", flush=True)
            return ms

        timed("pt_add", lambda: E.pt_add(pts, pts))
        timed("pt_mul_generator_public", lambda: E.pt_mul_generator_public(xs))
        timed("pt_mul (var-base)", lambda: E.pt_mul(xs, pts))
        timed("pt_mul_authenticated (2x)", lambda: E.pt_mul_authenticated((xs, a_s), pts))
        d, Em = E.pt_beaver_mask(xs, P, a_s, b_s)
        t1 = timed("pt_beaver_mask", lambda: E.pt_beaver_mask(xs, P, a_s, b_s, out=(d, Em)))
        out = E.empty_points(n, share=True)
        t2 = timed("pt_beaver_recombine", lambda: E.pt_beaver_recombine(0, key, d, d, Em, Em, (a_s, a_m), (b_s, b_m), (c_s, c_m), out=out))
        print(f"{curve:20s} two-party point Beaver mults/s (both parties on one GPU): {n / (2 * (t1 + t2)) * 1e3:,.0f}", flush=True)
        timed("pt_normalize", lambda: E.pt_normalize(pts))
        timed("pt_sum", lambda: E.pt_sum(pts))
        timed("pt_msm (bucket method)", lambda: E.pt_msm(xs, pts))
        E.close()
