"""The C-ABI shared library builds for sm_100a, loads, and exports exactly the entry points that
include/cfm_b200.h declares (no compute is launched: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cfm_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from cfm_b200 import _ffi
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert sorted(_ffi.SIGNATURES) == declared


def test_library_loads_and_exports_every_symbol(lib_built):
    from cfm_b200 import _ffi
    assert os.path.exists(lib_built)
    lib = _ffi.load_library(lib_built)  # binds every name in SIGNATURES or raises
    assert lib.cfm_abi_version() == 1
    raw = ctypes.CDLL(lib_built)
    for name in _declared_symbols():
        assert hasattr(raw, name), name


def test_library_holds_sm100a_code_only(lib_built):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", lib_built], capture_output=True,
                         text=True).stdout
    arches = set(re.findall(r"sm_(\d+a?)", out))
    assert arches == {"100a"}, arches


def test_rk_state_struct_layout_matches_header():
    from cfm_b200 import _ffi
    assert ctypes.sizeof(_ffi.RkState) == 72
    assert _ffi.RkState.err_acc.offset == 64


def test_missing_library_fails_loudly(tmp_path):
    from cfm_b200 import _ffi
    with pytest.raises(_ffi.CfmLibraryError):
        _ffi.load_library(str(tmp_path / "nope.so"))


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "cfm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+(scipy|ot)\b", src, flags=re.M), f
