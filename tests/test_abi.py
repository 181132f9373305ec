"""The C-ABI library loads and exports every symbol include/arkmpc_b200.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

from ark_mpc_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "arkmpc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(arkmpc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds():
    decl = declared_symbols()
    assert decl, "no symbols parsed from the header"
    assert sorted(nat.EXPORTED_SYMBOLS) == decl


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/arkmpc_b200.h but not exported"


def test_abi_version_and_status_strings():
    lib = nat.load()
    assert lib.arkmpc_abi_version() >= 1
    assert lib.arkmpc_status_string(0) == b"ok"
    assert b"invalid" in lib.arkmpc_status_string(-1)


def test_no_device_is_reported_not_faked(has_gpu):
    """Without a GPU the library must say so (never emulate): ctx_create fails with NO_DEVICE."""
    if has_gpu:
        return
    lib = nat.load()
    h = ctypes.c_void_p()
    assert lib.arkmpc_ctx_create(0, ctypes.byref(h)) == nat.ERR_NO_DEVICE
    assert not h.value
    assert nat.device_count() == 0
    import pytest
    from ark_mpc_b200.engine import Engine

    with pytest.raises(nat.ArkMpcError):
        Engine(0, "bn254_fr")


def test_missing_library_fails_loudly_not_silently(tmp_path, monkeypatch):
    """No CPU fallback and no hang: with the shared library absent, load() raises an ArkMpcError that says how to build it."""
    import pytest

    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "LIB_PATH", str(tmp_path / "libarkmpc_b200.so"))
    with pytest.raises(nat.ArkMpcError) as e:
        nat.load()
    assert "missing" in str(e.value) and "no CPU fallback" in str(e.value)
