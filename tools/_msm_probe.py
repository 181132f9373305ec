import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ark_mpc_b200.engine import Engine
log2n = int(sys.argv[1]); field = sys.argv[2]
E = Engine(0, field)
n = 1 << log2n
xs = E.random(1, 0, n); pts = E.pt_mul_generator_public(E.random(22, 0, n))
for _ in range(2):
    E.pt_msm(xs, pts)
torch.cuda.synchronize()
