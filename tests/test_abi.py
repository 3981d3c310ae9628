"""CPU suite: the C-ABI library loads, exports every symbol include/scvod.h declares, and fails loudly
(instead of silently falling back to a CPU path) when no CUDA device exists."""
import ctypes
import os
import re

import pytest

import conftest

HEADER = os.path.join(conftest.ROOT, "include", "scvod.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scvod_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    lib = pkg.load_library()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/scvod.h but not exported"
    for n in pkg.EXPORTS:
        assert n in names, f"{n} bound in Python but not declared in include/scvod.h"


def test_oracle_is_not_linked_into_the_product(pkg):
    lib_path = pkg.LIB_PATH
    data = open(lib_path, "rb").read()
    assert b"orc_push_scan" not in data and b"libscvod_oracle" not in data


@pytest.mark.skipif(conftest.has_gpu(), reason="checks the no-GPU failure path")
def test_create_fails_loudly_without_gpu(pkg):
    p = pkg.semantickitti_params()
    with pytest.raises(pkg.ScvodError) as e:
        pkg.SSC(p)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_error_codes_on_bad_arguments(pkg):
    lib = pkg.load_library()
    assert lib.scvod_grid_dims(None, None) == -1
    assert b"null" in lib.scvod_last_error()
    ctx = ctypes.c_void_p()
    p = pkg.semantickitti_params()
    assert lib.scvod_create(ctypes.byref(p), 0, -5, 1, ctypes.byref(ctx)) == -1
