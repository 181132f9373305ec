"""The C++ host mirror (host/arkmpc_host.hpp): compiles against the C ABI alone on CPU; on a GPU box the two-party test
binary (tests/host_cpp/test_host.cpp) runs the reference-style cases through it and checks them against the C oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_cpp", "test_host.cpp")
EXE = os.path.join(ROOT, "tests", "host_cpp", "test_host")


def build():
    from ark_mpc_b200.build import build_native
    from oracle import coracle

    build_native()
    coracle.build()
    deps = [SRC, os.path.join(ROOT, "host", "arkmpc_host.hpp"), os.path.join(ROOT, "include", "arkmpc_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "host"), SRC,
                               "-o", EXE, "-L", os.path.join(ROOT, "ark_mpc_b200", "lib"), "-larkmpc_b200", "-L", os.path.join(ROOT, "oracle"),
                               "-lark_oracle", "-lpthread", "-Wl,-rpath," + os.path.join(ROOT, "ark_mpc_b200", "lib"),
                               "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


TSRC = os.path.join(ROOT, "tests", "host_cpp", "test_threads.cpp")
TEXE = os.path.join(ROOT, "tests", "host_cpp", "test_threads")


def build_threads():
    from ark_mpc_b200.build import build_native
    from oracle import coracle

    build_native()
    coracle.build()
    deps = [TSRC, os.path.join(ROOT, "include", "arkmpc_b200.h")]
    if not os.path.exists(TEXE) or any(os.path.getmtime(d) > os.path.getmtime(TEXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), TSRC, "-o", TEXE,
                               "-L", os.path.join(ROOT, "ark_mpc_b200", "lib"), "-larkmpc_b200", "-L", os.path.join(ROOT, "oracle"),
                               "-lark_oracle", "-lpthread", "-Wl,-rpath," + os.path.join(ROOT, "ark_mpc_b200", "lib"),
                               "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-Wl,-rpath,/usr/local/cuda/lib64"])
    return TEXE


def test_thread_test_builds_and_refuses_to_run_without_a_gpu(has_gpu):
    exe = build_threads()
    if not has_gpu:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_one_context_driven_from_four_threads():
    """per-context thread safety (executor workers call gate closures concurrently)."""
    r = subprocess.run([build_threads(), "4", "6"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "thread-safety test OK" in r.stdout


def test_host_mirror_compiles_against_the_c_abi_only():
    exe = build()
    # no CUDA or torch headers are needed by the host mirror: only include/arkmpc_b200.h
    text = open(os.path.join(ROOT, "host", "arkmpc_host.hpp")).read()
    assert "cuda_runtime" not in text and "torch" not in text.split("#pragma once")[1]
    assert os.path.exists(exe)


@pytest.mark.parametrize("length", [0, 1, 3, 135, 136, 137, 271, 272, 32 * 1024 + 32, 100003])
def test_host_mirror_sha3_matches_hashlib(length):
    """The commitment hash of open_authenticated (commitment.rs:36-41, `sha3` crate) is SHA3-256 of FIPS 202."""
    import hashlib

    seed = 17 + length
    r = subprocess.run([build(), "--sha3", str(length), str(seed)], capture_output=True, text=True, timeout=120)
    msg = bytes(((seed + 131 * i) & 0xFFFFFFFF) >> 3 & 0xFF for i in range(length))
    assert r.returncode == 0 and r.stdout.strip() == hashlib.sha3_256(msg).hexdigest()


def test_host_mirror_refuses_to_run_without_a_gpu(has_gpu):
    if has_gpu:
        pytest.skip("GPU present")
    r = subprocess.run([build()], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_mirror_two_party_cases():
    r = subprocess.run([build()], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "host mirror tests OK" in r.stdout
