"""CPU-side checks of the drop-in boundary: librelp_gpu.so loads (no GPU needed) and exports every
symbol include/*.h declares; the ctypes table covers them all; no compute calls are made."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("relp_gpu.h", "relp_gpu_test.h", "relp_host.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(r[gh]_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from relp_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) > 40
    for name in sorted(decl):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(_lib.SYMBOLS), decl ^ set(_lib.SYMBOLS)


def test_product_path_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "relp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # no import, include, load or path reference of anything under oracle/ (comments may mention it)
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s*[<\"].*oracle", r"oracle/", r"fast_oracle",
                            r"relp_oracle", r"libfast_oracle"):
                    assert not re.search(pat, text, flags=re.M), f"{f} references the oracle ({pat})"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from relp_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(RuntimeError):
        _lib.load()
