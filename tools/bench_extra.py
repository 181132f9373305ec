#!/usr/bin/env python
"""Device timing of the "next row" kernels: batch inversion, FFT on shares, linear share gates (CUDA events)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ark_mpc_b200.engine import Engine

PEAK = 6549.8
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    E = Engine(0, "bn254_fr")

    def timed(name, fn, nbytes, n, reps=20):
        for _ in range(3):
            fn()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        s.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:44s} {ms * 1e3:10.1f} us  {n / ms / 1e6:9.3f} G elems/s  {nbytes / ms / 1e6:8.1f} GB/s algorithmic ({nbytes / ms / 1e6 / PEAK:.2f} of HBM peak)", flush=True)

    for log2n in (16, 20, 22, 24):
        n = 1 << log2n
        a, b, c, d = (E.random(i, 0, n) for i in range(1, 5))
        o1, o2 = E.empty(n), E.empty(n)
        key = E.download(E.random(9, 0, 1))[0].copy()
        print(f"--- n = 2^{log2n}")
        timed("fr_add", lambda: E.add(a, b, out=o1), 96 * n, n)
        timed("fr_mul", lambda: E.mul(a, b, out=o1), 96 * n, n)
        timed("share_add (2 planes)", lambda: E.share_add((a, b), (c, d)), 192 * n, n)
        timed("share_add_public (MAC update)", lambda: E.share_add_public(0, key, (a, b), c), 160 * n, n)
        timed("share_mul_public", lambda: E.share_mul_public((a, b), c), 160 * n, n)
        timed("mac_check", lambda: E.mac_check(key, a, b), 96 * n, n)
        timed("share_sum", lambda: E.share_sum((a, b)), 64 * n, n)
        timed("batch_inverse", lambda: E.batch_inverse(a, out=o1), 96 * n, n)  # reads a twice
        timed("fft (one plane, out of place)", lambda: E.fft(a), 64 * n, n, reps=10)
        timed("ifft (one plane)", lambda: E.fft(a, inverse=True), 64 * n, n, reps=10)
        del a, b, c, d, o1, o2
