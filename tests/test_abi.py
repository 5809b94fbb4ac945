"""CPU-only: the C-ABI library loads and exports every symbol include/mgfb.h declares; without
a GPU it fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest

from mgf_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mgfb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgfb_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(L.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in mgfb.h but not exported"
    bound = {n for n, _, _ in L.SYMBOLS}
    assert bound == set(declared), (bound ^ set(declared))


def test_struct_sizes_match_header():
    assert C.sizeof(L.Config) == 4 * 6 + 16
    assert C.sizeof(L.StepStats) == 64
    assert C.sizeof(L.SolveStats) == 32
    assert L.SHAPE_DTYPE.itemsize == 64


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mgf_b200
    with pytest.raises(mgf_b200.MgfbError) as e:
        mgf_b200.Context(device=0)
    assert e.value.code == L.ERR_CUDA


def test_product_never_touches_oracle():
    """The shipped package must not reference oracle/ (parity would be void)."""
    pkg = os.path.join(ROOT, "mgf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "oracle_lib" not in txt and "oracle/" not in txt, f


def test_integration_md_binds_every_symbol():
    """INTEGRATION.md's Rust `extern "C"` block is the binding a maintainer would paste: it must name every entry point of the header."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if f"pub fn {n}(" not in doc]
    assert not missing, missing


def test_rust_binding_matches_integration_md():
    """bindings/rust/mgf-b200/src/sys.rs is the extern block of INTEGRATION.md, verbatim (the crate cannot be compiled here: no rustc)."""
    import re
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.findall(r"```rust\n(.*?)```", doc, re.S)[0]
    rs = open(os.path.join(ROOT, "bindings", "rust", "mgf-b200", "src", "sys.rs")).read()
    assert rs.endswith(block), "regenerate sys.rs from INTEGRATION.md section 1"
    missing = [n for n in _declared_symbols() if f"pub fn {n}(" not in rs]
    assert not missing, missing
