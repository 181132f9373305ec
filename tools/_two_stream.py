#!/usr/bin/env python
"""Scratch: both parties on one GPU with one stream per party (as execute_mock_mpc runs them) vs one shared stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ark_mpc_b200.engine import Engine

n = 1 << 20
S = [torch.cuda.Stream(), torch.cuda.Stream()]
Es = []
for p in (0, 1):
    with torch.cuda.stream(S[p]):
        Es.append(Engine(0, "bn254_fr"))
E = Es[0]
with torch.cuda.stream(S[0]):
    key = [E.download(E.random(900 + p, 0, 1))[0].copy() for p in (0, 1)]
    P = [dict(x=E.random(10 + p, 0, n), y=E.random(20 + p, 0, n), a=(E.random(30 + p, 0, n), E.random(31 + p, 0, n)), b=(E.random(40 + p, 0, n), E.random(41 + p, 0, n)),
              c=(E.random(50 + p, 0, n), E.random(51 + p, 0, n))) for p in (0, 1)]
    de = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
    out = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
torch.cuda.synchronize()

def step_one_stream():
    with torch.cuda.stream(S[0]):
        for p in (0, 1):
            Es[0].beaver_mask(P[p]["x"], P[p]["y"], P[p]["a"][0], P[p]["b"][0], out=de[p])
        for p in (0, 1):
            Es[0].beaver_recombine(p, key[p], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], P[p]["a"], P[p]["b"], P[p]["c"], out=out[p])

ev_mask = [torch.cuda.Event(), torch.cuda.Event()]
ev_rec = [torch.cuda.Event(), torch.cuda.Event()]
def step_two_streams():
    for p in (0, 1):
        with torch.cuda.stream(S[p]):
            S[p].wait_event(ev_rec[1 - p])       # the peer's recombine of the previous step has read my d/e
            Es[p].beaver_mask(P[p]["x"], P[p]["y"], P[p]["a"][0], P[p]["b"][0], out=de[p])
            ev_mask[p].record(S[p])
    for p in (0, 1):
        with torch.cuda.stream(S[p]):
            S[p].wait_event(ev_mask[1 - p])      # the "network": the peer's d/e are ready
            Es[p].beaver_recombine(p, key[p], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], P[p]["a"], P[p]["b"], P[p]["c"], out=out[p])
            ev_rec[p].record(S[p])

for name, fn in (("one stream", step_one_stream), ("two streams", step_two_streams), ("one stream", step_one_stream), ("two streams", step_two_streams)):
    for p in (0, 1):
        ev_rec[p].record(S[p])
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 500
    e0.record(S[0])
    for _ in range(steps):
        fn()
    S[0].wait_stream(S[1])
    e1.record(S[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name:12s} {ms * 1e3:8.2f} us/step  {n / ms / 1e6:7.3f} G mults/s")
