"""The C-ABI library loads and exports every symbol include/dtcsim.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "dtcsim.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dtc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from dtcsim import capi
    if not os.path.exists(capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/dtcsim.h but not exported"
    assert set(names) == set(capi.EXPORTED), set(names) ^ set(capi.EXPORTED)
    lib.dtc_version.restype = ctypes.c_int
    assert lib.dtc_version() == 100


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or the emulator)."""
    pkg = os.path.join(ROOT, "noise-resilience-in-discrete-time-crystal-realizations-on-quantum-computers_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert "oracle" not in src.replace("oracle/philox.py", "").replace("the oracle", "") \
                    or f in ("__init__.py",), f"{f} mentions the oracle"
                assert "import emu" not in src and "program_interp" not in src


def test_missing_library_fails_loudly(monkeypatch):
    from dtcsim import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libdtcsim.so")
    with pytest.raises(RuntimeError):
        capi.load()
