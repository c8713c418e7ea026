"""The C-ABI library loads and exports every symbol include/dvst.h declares; host-side error paths (no GPU needed)."""
import ctypes
import os
import re

import dvst_b200  # noqa: F401
from dvst_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dvst.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvst_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dvst.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version_and_error_reporting():
    L = _lib.lib()
    assert L.dvst_abi_version() == 1
    h = ctypes.c_void_p()

    class Cfg(ctypes.Structure):
        _fields_ = [(n, ctypes.c_int) for n in ("depth", "num_frames", "pos_grid", "attention_type", "precision")]
    bad = Cfg(0, 8, 14, 0, 0)
    assert L.dvst_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b"depth" in L.dvst_last_error()
    ok = Cfg(2, 8, 14, 0, 0)
    assert L.dvst_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    # finalize without weights must fail loudly and name what is missing
    assert L.dvst_finalize_weights(h) != 0
    assert b"patch_embed.w" in L.dvst_last_error()
    assert L.dvst_workspace_bytes(h, 2, 8, 224, 224) > 0
    assert L.dvst_workspace_bytes(h, 0, 8, 224, 224) == -1
    L.dvst_destroy(h)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.lib()
        assert False, "expected DvstError"
    except _lib.DvstError as e:
        assert "no CPU fallback" in str(e)
