"""GPU parity tests of the scalar-field gates: every call goes through the C ABI (libarkmpc_b200.so)
and is compared limb-for-limb with the CPU oracle on the same inputs."""
import os

import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyoracle as po
from tests.util import FIELD_BY_ID, FIELD_NAME, TwoPartyData, aos, mont_scalar, split_aos

pytestmark = pytest.mark.gpu

FIDS = [0, 1]
SIZES = [1, 31, 1000, 70001]


@pytest.fixture(scope="module")
def engines():
    from ark_mpc_b200.engine import Engine

    return {fid: Engine(0, FIELD_NAME[fid]) for fid in FIDS}


def up(E, a):
    return E.upload(a)


def dn(E, t):
    return E.download(t)


@pytest.mark.parametrize("fid", FIDS)
def test_native_library_is_loaded(engines, fid):
    import ark_mpc_b200._native as nat

    assert nat.load().arkmpc_abi_version() == 2
    with open("/proc/self/maps") as f:
        assert "libarkmpc_b200.so" in f.read()
    assert engines[fid].sm_count >= 100


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", SIZES)
def test_scalar_gates(engines, fid, n):
    E = engines[fid]
    F = FIELD_BY_ID[fid]
    a = co.synth(fid, 1, 0, n)
    b = co.synth(fid, 2, 0, n)
    edge = co.to_mont(fid, co.ints_to_limbs([0, 1, F.p - 1]))
    k = min(3, n)
    a[:k] = edge[:k]
    b[:k] = edge[:k][::-1]
    A, B = up(E, a), up(E, b)
    assert np.array_equal(dn(E, E.add(A, B)), co.scalar_add(fid, a, b))
    assert np.array_equal(dn(E, E.sub(A, B)), co.scalar_sub(fid, a, b))
    assert np.array_equal(dn(E, E.mul(A, B)), co.scalar_mul(fid, a, b))
    zero = np.zeros_like(a)
    assert np.array_equal(dn(E, E.neg(A)), co.scalar_sub(fid, zero, a))
    assert np.array_equal(dn(E, E.from_mont(A)), co.from_mont(fid, a))
    assert np.array_equal(dn(E, E.to_mont(E.from_mont(A))), a)
    s = co.synth(fid, 3, 0, 1)[0]
    assert np.array_equal(dn(E, E.scale(A, s)), co.scalar_mul(fid, a, np.tile(s, (n, 1))))


@pytest.mark.parametrize("fid", FIDS)
def test_random_matches_oracle_generator(engines, fid):
    E = engines[fid]
    got = dn(E, E.random(0xA11CE, 12345, 5000))
    assert np.array_equal(got, co.synth(fid, 0xA11CE, 12345, 5000))


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", SIZES)
def test_beaver_mask_and_recombine_bit_exact(engines, fid, n):
    """authenticated_scalar.rs:848-879 per party, CUDA (fused) vs the oracle's unfused reference sequence."""
    E = engines[fid]
    D = TwoPartyData(fid, n, seed=77 + n, edge=True)
    o0, o1, d_open, e_open = D.oracle_batch_mul()
    masks = []
    for p in (0, 1):
        P = D.party(p)
        d, e = E.beaver_mask(up(E, P["x"][0]), up(E, P["y"][0]), up(E, P["a"][0]), up(E, P["b"][0]))
        want_d, want_e = co.beaver_mask(fid, aos(*P["x"]), aos(*P["y"]), aos(*P["a"]), aos(*P["b"]))
        assert np.array_equal(dn(E, d), want_d) and np.array_equal(dn(E, e), want_e)
        masks.append((d, e))
    for p, want in ((0, o0), (1, o1)):
        P = D.party(p)
        pl = lambda t: (up(E, t[0]), up(E, t[1]))
        (os_, om_), (do, eo) = E.beaver_recombine(p, P["key"], masks[p][0], masks[p][1], masks[1 - p][0], masks[1 - p][1],
                                                  pl(P["a"]), pl(P["b"]), pl(P["c"]), want_open=True)
        assert np.array_equal(dn(E, os_), want[:, :4]), f"share mismatch party {p}"
        assert np.array_equal(dn(E, om_), want[:, 4:]), f"mac mismatch party {p}"
        assert np.array_equal(dn(E, do), d_open) and np.array_equal(dn(E, eo), e_open)
        # without the optional opened outputs
        (os2, om2), _ = E.beaver_recombine(p, P["key"], masks[p][0], masks[p][1], masks[1 - p][0], masks[1 - p][1],
                                           pl(P["a"]), pl(P["b"]), pl(P["c"]))
        assert np.array_equal(dn(E, os2), want[:, :4]) and np.array_equal(dn(E, om2), want[:, 4:])
    # protocol identity: opened product == x*y and MAC shares sum to key*x*y
    xy = co.scalar_mul(fid, D.xv, D.yv)
    assert np.array_equal(co.scalar_add(fid, o0[:, :4], o1[:, :4]), xy)
    assert np.array_equal(co.scalar_add(fid, o0[:, 4:], o1[:, 4:]), co.scalar_mul(fid, xy, np.tile(D.key, (n, 1))))


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", SIZES + [300000])
def test_beaver_recombine_sum_bit_exact(engines, fid, n):
    """Phase 2 fused with the Sum that follows it (inner product, circuits.rs:22-50): one ScalarShare per party, limb-exact against the
    oracle's batch_mul followed by its share sum, and against the unfused device path."""
    E = engines[fid]
    D = TwoPartyData(fid, n, seed=501 + n, edge=True)
    o0, o1, _, _ = D.oracle_batch_mul()
    masks = []
    for p in (0, 1):
        P = D.party(p)
        masks.append(E.beaver_mask(up(E, P["x"][0]), up(E, P["y"][0]), up(E, P["a"][0]), up(E, P["b"][0])))
    for p, want in ((0, o0), (1, o1)):
        P = D.party(p)
        pl = lambda t: (up(E, t[0]), up(E, t[1]))
        args = (p, P["key"], masks[p][0], masks[p][1], masks[1 - p][0], masks[1 - p][1], pl(P["a"]), pl(P["b"]), pl(P["c"]))
        s, m = E.beaver_recombine_sum(*args)
        want_sum = co.share_sum(fid, want)
        assert np.array_equal(dn(E, s)[0], want_sum[:4]), f"share sum mismatch party {p}"
        assert np.array_equal(dn(E, m)[0], want_sum[4:]), f"mac sum mismatch party {p}"
        unfused = E.share_sum(E.beaver_recombine(*args)[0])
        assert np.array_equal(dn(E, unfused[0]), dn(E, s)) and np.array_equal(dn(E, unfused[1]), dn(E, m))


def test_beaver_recombine_sum_empty_and_invalid(engines):
    import ark_mpc_b200._native as nat

    E = engines[0]
    z = E.empty(0)
    s, m = E.beaver_recombine_sum(0, co.synth(0, 1, 0, 1)[0], z, z, z, z, (z, z), (z, z), (z, z))
    assert not dn(E, s).any() and not dn(E, m).any()
    a = E.random(3, 0, 8)
    with pytest.raises(nat.ArkMpcError):
        E.beaver_recombine_sum(2, co.synth(0, 1, 0, 1)[0], a, a, a, a, (a, a), (a, a), (a, a))


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", [1, 1000, 4099])
def test_share_gates(engines, fid, n):
    E = engines[fid]
    a = aos(co.synth(fid, 5, 0, n), co.synth(fid, 6, 0, n))
    b = aos(co.synth(fid, 7, 0, n), co.synth(fid, 8, 0, n))
    v = co.synth(fid, 9, 0, n)
    a[0] = 0
    v[0] = 0
    key = co.synth(fid, 10, 0, 1)[0]
    pl = lambda t: tuple(up(E, h) for h in split_aos(t))
    z = lambda planes: aos(dn(E, planes[0]), dn(E, planes[1]))
    A, B, V = pl(a), pl(b), up(E, v)
    assert np.array_equal(z(E.share_add(A, B)), co.batch_add(fid, a, b))
    assert np.array_equal(z(E.share_sub(A, B)), co.batch_sub(fid, a, b))
    assert np.array_equal(z(E.share_neg(A)), co.batch_neg(fid, a))
    assert np.array_equal(z(E.share_mul_public(A, V)), co.batch_mul_public(fid, a, v))
    for party in (0, 1):
        assert np.array_equal(z(E.share_add_public(party, key, A, V)), co.batch_add_public(fid, party, key, a, v))
        assert np.array_equal(z(E.share_add_public(party, key, A, V, sub=True)), co.batch_add_public(fid, party, key, a, v, sub=True))
    assert np.array_equal(dn(E, E.mac_check(key, V, A[1])), co.mac_check(fid, key, v, a))
    got = E.share_sum(A)
    assert np.array_equal(np.concatenate([dn(E, got[0])[0], dn(E, got[1])[0]]), co.share_sum(fid, a))
    assert np.array_equal(dn(E, E.sum(V))[0], co.share_sum(fid, aos(v, v))[:4])
    # zip / unzip round trip against the reference AoS image
    aos_dev = E.share_zip(A)
    assert np.array_equal(dn(E, aos_dev), a)
    back = E.share_unzip(aos_dev)
    assert np.array_equal(z(back), a)


@pytest.mark.parametrize("fid", FIDS)
def test_sum_is_zero_and_bytes(engines, fid):
    E = engines[fid]
    F = FIELD_BY_ID[fid]
    n = 5000
    a = co.synth(fid, 11, 0, n)
    a[7] = 0
    na = co.scalar_sub(fid, np.zeros_like(a), a)
    assert E.sum_is_zero(up(E, a), up(E, na))
    na[4321, 0] ^= 1
    assert not E.sum_is_zero(up(E, a), up(E, na))
    got = dn(E, E.to_bytes_be(up(E, a[:64])).view(__import__("torch").int64)).view(np.uint8).reshape(-1, 32)
    ints = co.limbs_to_ints(co.from_mont(fid, a[:64]))
    assert [bytes(r) for r in got] == [F.to_bytes_be(v) for v in ints]


@pytest.mark.parametrize("fid", FIDS)
def test_batch_mul_full_size_properties(engines, fid):
    """BASELINE batch 2^20: size-independent checks (protocol identity + MAC consistency) on the whole
    batch computed on the GPU, plus limb-exact comparison with the oracle on a strided sample."""
    E = engines[fid]
    n = 1 << 20
    key0, key1 = co.synth(fid, 900, 0, 1)[0], co.synth(fid, 901, 0, 1)[0]
    key = co.scalar_add(fid, key0.reshape(1, 4), key1.reshape(1, 4))[0]

    def shared(seed, val=None):
        v = E.random(seed, 0, n) if val is None else val
        s0 = E.random(seed + 1, 0, n)
        m0 = E.random(seed + 2, 0, n)
        return v, (s0, m0), (E.sub(v, s0), E.sub(E.scale(v, key), m0))

    xv, x0, x1 = shared(10)
    yv, y0, y1 = shared(20)
    av, a0, a1 = shared(30)
    bv, b0, b1 = shared(40)
    _, c0, c1 = shared(50, E.mul(av, bv))
    m0 = E.beaver_mask(x0[0], y0[0], a0[0], b0[0])
    m1 = E.beaver_mask(x1[0], y1[0], a1[0], b1[0])
    (r0, _) = E.beaver_recombine(0, key0, m0[0], m0[1], m1[0], m1[1], a0, b0, c0)
    (r1, (do, eo)) = E.beaver_recombine(1, key1, m1[0], m1[1], m0[0], m0[1], a1, b1, c1, want_open=True)
    xy = E.mul(xv, yv)
    import torch

    assert torch.equal(E.add(r0[0], r1[0]), xy)
    assert torch.equal(E.add(r0[1], r1[1]), E.scale(xy, key))
    assert torch.equal(do, E.sub(xv, av)) and torch.equal(eo, E.sub(yv, bv))
    # strided sample against the oracle's reference sequence
    idx = np.arange(0, n, 4099)
    h = lambda t: dn(E, t)[idx]
    ha = lambda pl: aos(h(pl[0]), h(pl[1]))
    o0, o1, d, e = co.two_party_batch_mul(fid, 2, (key0, key1), (ha(x0), ha(x1)), (ha(y0), ha(y1)), (ha(a0), ha(a1)),
                                          (ha(b0), ha(b1)), (ha(c0), ha(c1)))
    assert np.array_equal(ha(r0), o0) and np.array_equal(ha(r1), o1)
    assert np.array_equal(h(do), d) and np.array_equal(h(eo), e)


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", [1, 1000, 200001])
@pytest.mark.parametrize("share_planes", [False, True])
def test_host_buffer_batch_mul(engines, fid, n, share_planes):
    """The end-to-end C-ABI path over host buffers (begin -> exchange -> finish), chunked/pipelined: operands as the reference's
    AoS ScalarShare images, or x and y as planes of their share halves (arkmpc_fr_batch_mul_begin_host_shares)."""
    E = engines[fid]
    D = TwoPartyData(fid, n, seed=5 + n)
    o0, o1, d_open, e_open = D.oracle_batch_mul()
    sess, de = [], []
    for p in (0, 1):
        P = D.party(p)
        de_mine = np.empty((2 * n, 4), dtype=np.uint64)
        if share_planes:
            xs, ys = np.ascontiguousarray(P["x"][0]), np.ascontiguousarray(P["y"][0])
            sess.append(E.batch_mul_begin_host_shares(p, P["key"], xs, ys, aos(*P["a"]), aos(*P["b"]), aos(*P["c"]), de_mine))
        else:
            sess.append(E.batch_mul_begin_host(p, P["key"], aos(*P["x"]), aos(*P["y"]), aos(*P["a"]), aos(*P["b"]), aos(*P["c"]), de_mine))
        de.append(de_mine)
    for p, want in ((0, o0), (1, o1)):
        out = np.empty((n, 8), dtype=np.uint64)
        de_open = np.empty((2 * n, 4), dtype=np.uint64) if p == 0 else None
        E.batch_mul_finish_host(sess[p], de[1 - p], out, de_open)
        assert np.array_equal(out, want)
        if de_open is not None:
            assert np.array_equal(de_open[:n], d_open) and np.array_equal(de_open[n:], e_open)


def test_invalid_arguments_are_rejected(engines):
    import ctypes as C

    import ark_mpc_b200._native as nat

    E = engines[0]
    a = E.random(1, 0, 8)
    with pytest.raises(nat.ArkMpcError):
        E._call("arkmpc_fr_add", 99, 8, E._p(a), E._p(a), E._p(a))  # unknown field
    with pytest.raises(nat.ArkMpcError):
        E._call("arkmpc_fr_add", 0, 8, None, E._p(a), E._p(a))  # null pointer
    with pytest.raises(nat.ArkMpcError):
        E._call("arkmpc_fr_add", 0, 7, C.c_void_p(a.data_ptr() + 8), E._p(a), E._p(a))  # misaligned plane
    k = np.zeros(4, dtype=np.uint64)
    with pytest.raises(nat.ArkMpcError):
        E.beaver_recombine(2, k, a, a, a, a, (a, a), (a, a), (a, a))  # bad party id
    # n == 0 is a no-op (authenticated_scalar.rs:854-856 returns an empty vector)
    E._call("arkmpc_fr_add", 0, 0, None, None, None)


def test_multi_gpu_open_gather():
    """Needs >= 2 GPUs on the box (skipped otherwise): tests/multi_gpu_check.py under torchrun, one rank per GPU."""
    import subprocess
    import sys

    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n_dev, 8)}", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tests", "multi_gpu_check.py"), "20000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("multi-GPU open gather OK") == min(n_dev, 8)


import torch  # noqa: E402
import ark_mpc_b200._native as nat  # noqa: E402
from ark_mpc_b200.engine import Engine  # noqa: E402


def test_native_allgather_single_rank():
    """arkmpc_nccl_unique_id / arkmpc_nccl_init / arkmpc_allgather_open on a world of one: NCCL is resolved at run
    time behind the C ABI and the gathered planes equal the local rows (the multi-rank form is covered by multi_gpu_check.py)."""
    import ctypes as C

    E = Engine(0, "bn254_fr")
    n = 3001
    d, e = E.random(1, 0, n), E.random(2, 0, n)
    ident = (C.c_uint8 * 128)()
    nat.check(E.lib.arkmpc_nccl_unique_id(ident), "arkmpc_nccl_unique_id", E.ctx)
    nat.check(E.lib.arkmpc_nccl_init(E.ctx, 1, 0, ident), "arkmpc_nccl_init", E.ctx)
    d_all, e_all = torch.zeros_like(d), torch.zeros_like(e)
    E._call("arkmpc_allgather_open", n, E._p(d), E._p(e), E._p(d_all), E._p(e_all))
    E.sync()
    assert torch.equal(d_all, d) and torch.equal(e_all, e)
    with pytest.raises(nat.ArkMpcError):  # a second communicator on the same context is refused
        nat.check(E.lib.arkmpc_nccl_init(E.ctx, 1, 0, ident), "arkmpc_nccl_init", E.ctx)
    nat.check(E.lib.arkmpc_nccl_destroy(E.ctx), "arkmpc_nccl_destroy", E.ctx)
    with pytest.raises(nat.ArkMpcError):  # no communicator any more
        E._call("arkmpc_allgather_open", n, E._p(d), E._p(e), E._p(d_all), E._p(e_all))
    E.close()


@pytest.mark.parametrize("field", ["bn254_fr", "curve25519_fr"])
def test_validate_rejects_non_canonical_peer_values(field):
    """arkmpc_fr_validate: what arkworks' deserialisation enforces on values from the peer (scalar.rs:187-202)."""
    from ark_mpc_b200 import fields as fl

    E = Engine(0, field)
    n = 5000
    a = E.random(9, 0, n)
    assert E.validate(a)
    p = fl.MODULUS[field]
    for bad in (p, p + 1, (1 << 256) - 1):
        b = a.clone()
        b[n - 7] = torch.from_numpy(fl.int_to_limbs(bad).view(np.int64))
        assert not E.validate(b)
    b = a.clone()
    b[0] = torch.from_numpy(fl.int_to_limbs(p - 1).view(np.int64))
    assert E.validate(b)
    assert E.validate(E.empty(0))
    E.close()


def test_engine_follows_torch_current_stream():
    """an Engine built outside a `with torch.cuda.stream(...)` block and used inside one must run on that stream."""
    E = Engine(0, "bn254_fr")
    s = torch.cuda.Stream()
    n = 1 << 16
    with torch.cuda.stream(s):
        a, b = E.random(1, 0, n), E.random(2, 0, n)
        assert int(E.lib.arkmpc_ctx_get_stream(E.ctx) or 0) == s.cuda_stream
        c = E.mul(a, b)
    s.synchronize()
    c2 = E.mul(a, b)  # back on the default stream
    assert int(E.lib.arkmpc_ctx_get_stream(E.ctx) or 0) == torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    assert torch.equal(c, c2)
    E.close()


@pytest.mark.parametrize("fid", FIDS)
@pytest.mark.parametrize("n", [1000, 1 << 18, (1 << 20) + 77])
def test_independence_hint_keeps_results_and_ordering(engines, fid, n):
    """arkmpc_ctx_hint_independent: the four launches of a two-party step with the second launch of each phase hinted, repeated
    back to back with the SAME buffers (so each phase truly depends on the one before it: recombine reads what the masks wrote,
    the next step's masks overwrite what this step's recombines read).  Every step's outputs must equal the oracle's."""
    E = engines[fid]
    D = TwoPartyData(fid, n, seed=4242 + n)
    o0, o1, _, _ = D.oracle_batch_mul()
    P = [D.party(p) for p in (0, 1)]
    dev = [{k: ((up(E, v[0]), up(E, v[1])) if isinstance(v, tuple) else v) for k, v in P[p].items()} for p in (0, 1)]
    de = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
    out = [(E.empty(n), E.empty(n)) for _ in (0, 1)]
    zero = [t for pair in de + out for t in pair]
    for rep in range(6):
        for t in zero:
            t.zero_()  # torch kernels between the steps: plain launches, ordered as usual
        for p in (0, 1):
            if p == 1:
                E.hint_independent()
            E.beaver_mask(dev[p]["x"][0], dev[p]["y"][0], dev[p]["a"][0], dev[p]["b"][0], out=de[p])
        for p in (0, 1):
            if p == 1:
                E.hint_independent()
            E.beaver_recombine(p, P[p]["key"], de[p][0], de[p][1], de[1 - p][0], de[1 - p][1], dev[p]["a"], dev[p]["b"], dev[p]["c"], out=out[p])
        if rep % 2 == 1:  # un-synchronised back-to-back steps on odd reps, checked ones on even reps
            continue
        for p, want in ((0, o0), (1, o1)):
            assert np.array_equal(dn(E, out[p][0]), want[:, :4]) and np.array_equal(dn(E, out[p][1]), want[:, 4:]), f"rep {rep} party {p}"
    for p, want in ((0, o0), (1, o1)):
        assert np.array_equal(dn(E, out[p][0]), want[:, :4]) and np.array_equal(dn(E, out[p][1]), want[:, 4:])


def test_device_memory_cache_reuses_blocks_in_stream_order():
    """arkmpc_malloc / arkmpc_free (include/arkmpc_b200.h, memory): a freed block is handed to ANOTHER context without a host
    synchronisation, and that context's writes still land after the work the first context had in flight on the block."""
    import ctypes as C

    import torch

    import ark_mpc_b200._native as nat

    lib = nat.load()
    ctx = [C.c_void_p(), C.c_void_p()]
    for c in ctx:
        nat.check(lib.arkmpc_ctx_create(0, C.byref(c)), "arkmpc_ctx_create")
    try:
        n = 1 << 22
        nbytes = n * 32
        nat.check(lib.arkmpc_mem_trim(ctx[0]), "trim", ctx[0])
        held = C.c_size_t(1)
        nat.check(lib.arkmpc_mem_cached_bytes(ctx[0], C.byref(held)), "cached", ctx[0])
        assert held.value == 0
        dev = torch.device("cuda:0")
        x, y, keep, want = (torch.empty((n, 4), dtype=torch.int64, device=dev) for _ in range(4))
        torch.cuda.synchronize()
        u64 = lambda t: C.c_void_p(t.data_ptr())
        a = C.c_void_p()
        nat.check(lib.arkmpc_malloc(ctx[0], nbytes - 8, C.byref(a)), "malloc", ctx[0])  # rounds up to a 2 MiB multiple
        pa = a
        nat.check(lib.arkmpc_fr_random(ctx[0], 0, 11, 0, n, u64(want)), "random", ctx[0])
        nat.check(lib.arkmpc_fr_random(ctx[0], 0, 11, 0, n - 1, pa), "random", ctx[0])
        nat.check(lib.arkmpc_fr_random(ctx[0], 0, 5, 0, n, u64(x)), "random", ctx[0])
        for _ in range(24):  # ~10 ms of work queued on context 0 ahead of the read of the block
            nat.check(lib.arkmpc_fr_batch_inverse(ctx[0], 0, n, u64(x), u64(y)), "inverse", ctx[0])
        nat.check(lib.arkmpc_memcpy_d2d(ctx[0], C.c_void_p(keep.data_ptr()), a, (n - 1) * 32), "d2d", ctx[0])
        nat.check(lib.arkmpc_free(ctx[0], a), "free", ctx[0])
        nat.check(lib.arkmpc_mem_cached_bytes(ctx[0], C.byref(held)), "cached", ctx[0])
        assert held.value == nbytes  # (n * 32 - 8) rounded up to the 2 MiB class
        b = C.c_void_p()
        nat.check(lib.arkmpc_malloc(ctx[1], nbytes, C.byref(b)), "malloc", ctx[1])
        assert b.value == a.value  # the cached block, no cudaMalloc
        nat.check(lib.arkmpc_fr_random(ctx[1], 0, 99, 0, n, b), "random", ctx[1])  # overwrites it
        for c in ctx:
            nat.check(lib.arkmpc_ctx_sync(c), "sync", c)
        assert torch.equal(keep[: n - 1], want[: n - 1])  # the copy read the block before context 1's kernel wrote it
        nat.check(lib.arkmpc_free(ctx[1], b), "free", ctx[1])
        nat.check(lib.arkmpc_mem_trim(ctx[1]), "trim", ctx[1])
        nat.check(lib.arkmpc_mem_cached_bytes(ctx[1], C.byref(held)), "cached", ctx[1])
        assert held.value == 0
    finally:
        for c in ctx:
            lib.arkmpc_ctx_destroy(c)


def test_small_host_to_device_copies_are_staged_and_release_the_source():
    """arkmpc_memcpy_h2d: copies of up to 256 KB go through a pinned ring of 8 slots; the source may be overwritten as soon as
    the call returns, and more copies than slots in a row stay correct."""
    import ctypes as C

    import torch

    import ark_mpc_b200._native as nat

    lib = nat.load()
    ctx = C.c_void_p()
    nat.check(lib.arkmpc_ctx_create(0, C.byref(ctx)), "arkmpc_ctx_create")
    try:
        rng = np.random.default_rng(5)
        sizes = [32, 100, 4096, 32 * 1024, 256 * 1024, 256 * 1024 + 32, 1 << 20] * 5
        total = sum(sizes)
        devbuf = torch.zeros(total, dtype=torch.uint8, device="cuda:0")
        torch.cuda.synchronize()
        want = np.empty(total, dtype=np.uint8)
        off = 0
        for s in sizes:
            src = rng.integers(0, 256, s, dtype=np.uint8)
            want[off:off + s] = src
            nat.check(lib.arkmpc_memcpy_h2d(ctx, C.c_void_p(devbuf.data_ptr() + off), src.ctypes.data_as(C.c_void_p), s), "h2d", ctx)
            src[:] = 0xEE  # pageable source: consumed when the call returns
            off += s
        nat.check(lib.arkmpc_ctx_sync(ctx), "sync", ctx)
        assert np.array_equal(devbuf.cpu().numpy(), want)
    finally:
        lib.arkmpc_ctx_destroy(ctx)


@pytest.mark.parametrize("share_planes", [False, True])
def test_host_buffer_batch_mul_many_chunks(monkeypatch, share_planes):
    """The host path's chunk pipeline with a small staging chunk (2^12 gates: a dozen chunks over three slots, ragged last one),
    so that slot reuse and the chunk boundaries are exercised at a size the oracle finishes quickly."""
    from ark_mpc_b200.engine import Engine

    monkeypatch.setenv("ARKMPC_CHUNK_LOG2", "12")
    E = Engine(0, FIELD_NAME[0])  # the chunk size is read when the context is created
    try:
        n = 12 * 4096 + 777
        D = TwoPartyData(0, n, seed=99)
        o0, o1, d_open, e_open = D.oracle_batch_mul()
        sess, de = [], []
        for p in (0, 1):
            P = D.party(p)
            de_mine = np.empty((2 * n, 4), dtype=np.uint64)
            if share_planes:
                sess.append(E.batch_mul_begin_host_shares(p, P["key"], np.ascontiguousarray(P["x"][0]), np.ascontiguousarray(P["y"][0]),
                                                          aos(*P["a"]), aos(*P["b"]), aos(*P["c"]), de_mine))
            else:
                sess.append(E.batch_mul_begin_host(p, P["key"], aos(*P["x"]), aos(*P["y"]), aos(*P["a"]), aos(*P["b"]), aos(*P["c"]), de_mine))
            de.append(de_mine)
        for p, want in ((0, o0), (1, o1)):
            out = np.empty((n, 8), dtype=np.uint64)
            de_open = np.empty((2 * n, 4), dtype=np.uint64) if p == 0 else None
            E.batch_mul_finish_host(sess[p], de[1 - p], out, de_open)
            assert np.array_equal(out, want)
            if de_open is not None:
                assert np.array_equal(de_open[:n], d_open) and np.array_equal(de_open[n:], e_open)
    finally:
        E.close()
