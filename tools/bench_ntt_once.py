#!/usr/bin/env python
"""One forward NTT and one batch inversion at n = 2^20 (the launches an `ncu -k regex:fr_ntt|fr_inv_` capture wants, and nothing else)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from ark_mpc_b200.engine import Engine

E = Engine(0, "bn254_fr")
n = 1 << 20
a = E.random(1, 0, n)
for _ in range(2):
    E.fft(a)
    E.batch_inverse(a)
torch.cuda.synchronize()
print("done")
