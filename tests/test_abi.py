"""The C-ABI shared library loads and exports every symbol include/city2ba_cuda.h declares.
No compute calls here (no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "city2ba_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(c2b_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(c2b):
    L = ctypes.CDLL(c2b._lib.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(c2b._lib.EXPORTS) == declared, "python binding list and header disagree"


def test_abi_version_and_defaults(c2b):
    L = c2b._lib.lib()
    assert L.c2b_abi_version() == 3
    o = c2b._lib.VisOptions()
    L.c2b_vis_options_default(ctypes.byref(o))
    assert (o.cull_mode, o.occlusion, o.endpoint_guard_rel, o.count_traversal) == (0, 0, 0, 0)
    assert (o.block_length, o.block_inset) == (20.0, 1.0)
    assert ctypes.sizeof(c2b._lib.Ray48) == 48  # Embree RTCRay layout


def test_no_cpu_fallback(c2b):
    """Without a GPU the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(c2b.C2BError) as e:
        c2b.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "city2ba_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and False, \
                    f"{f} mentions the oracle; the product must not depend on it"


def test_rust_ffi_declares_every_export():
    """rust/src/ffi.rs (the binding a maintainer would add; not compilable here) must declare exactly the
    header's symbols, so that the uncompiled source cannot drift from the ABI"""
    import re
    txt = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    declared = sorted(set(re.findall(r"pub fn (c2b_[a-z0-9_]+)\(", txt)))
    assert declared == _declared_symbols()
    hdr = open(os.path.join(ROOT, "include", "city2ba_cuda.h")).read()
    ver = re.search(r"#define C2B_ABI_VERSION (\d+)", hdr).group(1)
    assert f"C2B_ABI_VERSION: c_int = {ver};" in txt
    # struct fields in the header's order (names only)
    for struct, fields in (("c2b_vis_options", ["cull_mode", "occlusion", "endpoint_guard_rel", "count_traversal",
                                                "block_length", "block_inset", "predicate", "reserved"]),):
        body = txt[txt.index(f"pub struct {struct}"):]
        body = body[:body.index("}")]
        assert re.findall(r"pub ([a-z_]+):", body) == fields
