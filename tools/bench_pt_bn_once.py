#!/usr/bin/env python
"""Point Beaver recombine (BN254 G1, or Curve25519 with a second argument `ed25519`) at n = 2^17 by default, a few launches and nothing else (for `ncu -k regex:pt_beaver_recombine`), and its
event-timed duration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ark_mpc_b200.engine import Engine

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 17)
ed = len(sys.argv) > 2 and sys.argv[2] == "ed25519"
E = Engine(0, "curve25519_fr" if ed else "bn254_fr")
E.bind_curve("curve25519_edwards" if ed else "bn254_g1")
rnd = lambda seed: E.random(seed, 0, n)
key = E.download(E.random(77, 0, 1))[0].copy()
xs, a_s, a_m, b_s, b_m, c_s, c_m = (rnd(i) for i in range(1, 8))
P = E.pt_mul_generator((rnd(20), rnd(21)))
d, Em = E.pt_beaver_mask(xs, P, a_s, b_s)
out = E.empty_points(n, share=True)
run = lambda: E.pt_beaver_recombine(0, key, d, d, Em, Em, (a_s, a_m), (b_s, b_m), (c_s, c_m), out=out)
run()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(3):
    run()
ev1.record()
torch.cuda.synchronize()
blk = "512" if ed else os.environ.get("ARKMPC_PT_BN_BLOCK", "auto")
print(f"{'curve25519_edwards' if ed else 'bn254_g1'} pt_beaver_recombine n=2^{n.bit_length() - 1}: {ev0.elapsed_time(ev1) / 3:.3f} ms (block {blk})")
