#!/bin/bash
# compute-sanitizer over a small run of every kernel family (memcheck, then racecheck on the shared-memory kernels)
OUT=gpurun_out/sanitize
mkdir -p $OUT
cat > /tmp/san_small.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from ark_mpc_b200.engine import Engine
for field in ("bn254_fr", "curve25519_fr"):
    E = Engine(0, field)
    n = 3001
    r = lambda s: E.random(s, 0, n)
    key = E.download(E.random(99, 0, 1))[0].copy()
    a, b, c, d, e, f = (r(i) for i in range(1, 7))
    E.add(a, b); E.sub(a, b); E.mul(a, b); E.neg(a); E.scale(a, key); E.to_mont(E.from_mont(a)); E.to_bytes_be(a)
    E.share_add((a, b), (c, d)); E.share_sub((a, b), (c, d)); E.share_neg((a, b)); E.share_add_public(0, key, (a, b), c); E.share_add_public(1, key, (a, b), c, sub=True)
    E.share_mul_public((a, b), c); E.mac_check(key, a, b); E.sum_is_zero(a, E.neg(a)); E.share_sum((a, b)); E.sum(a); E.share_zip((a, b)); E.batch_inverse(a)
    de = E.beaver_mask(a, b, c, d)
    E.beaver_recombine(0, key, de[0], de[1], e, f, (a, b), (c, d), (e, f), want_open=True)
    E.beaver_recombine(1, key, de[0], de[1], e, f, (a, b), (c, d), (e, f))
    if field == "bn254_fr":
        E.fft(r(7)[:2048].contiguous()); E.fft(r(8)[:2048].contiguous(), inverse=True); E.share_fft((r(9)[:64].contiguous(), r(10)[:64].contiguous()))
    m = 130
    s1, s2 = (E.random(20 + i, 0, m) for i in range(2))
    P = E.pt_mul_generator((s1, s2)); pts = E.pt_mul_generator_public(s1)
    E.pt_add(pts, pts); E.pt_sub(pts, pts); E.pt_neg(pts); E.pt_mul(s2, pts); E.pt_share_mul_public(s1, P); E.pt_mul_authenticated((s1, s2), pts)
    E.pt_share_add_public(0, key, P, pts); E.pt_mac_check(key, pts, P); E.pt_sum_is_identity(pts, E.pt_neg(pts)); E.pt_normalize(P)
    dm, Em = E.pt_beaver_mask(s1, P, s2, s1)
    E.pt_beaver_recombine(0, key, dm, dm, Em, Em, (s1, s2), (s2, s1), (s1, s1), want_open=True)
    torch.cuda.synchronize()
    E.close()
print("sanitize workload done")
PY
cp /tmp/san_small.py $OUT/san_small.py
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python $OUT/san_small.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/memcheck.log
ARKMPC_RECOMBINE=tma timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python $OUT/san_small.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $OUT/racecheck.log
