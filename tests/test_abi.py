"""The C ABI: libswapnet_b200.so loads and exports every symbol include/swapnet_b200.h declares,
and the ctypes table in swapnet_b200/_lib.py covers exactly that set (no compute: runs without a GPU)."""
import ctypes
import os
import re

from swapnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "swapnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(sn_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from swapnet_b200 import build

        build.build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert decl == set(_lib.SIGNATURES), (decl ^ set(_lib.SIGNATURES))


def test_error_plumbing_without_gpu():
    lib = _lib.load(build_if_missing=True)
    assert lib.sn_version().startswith(b"swapnet_b200")
    # a null descriptor must come back as an error code + message, not a crash
    h = ctypes.c_void_p()
    rc = lib.sn_tap_gemm_plan_create(None, ctypes.byref(h))
    assert rc != 0 and b"null" in lib.sn_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest

    with pytest.raises(_lib.SwapnetB200Error):
        _lib.load()
