"""The C-ABI library loads on a CPU-only box and exports exactly what include/*.h declares. No compute calls."""
import ctypes
import glob
import os
import re

import pytest

from fairygen_b200 import _lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for path in glob.glob(os.path.join(REPO, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
        names |= set(re.findall(r"\b(fgb_[a-z0-9_]+)\s*\(", text))
    return names


@pytest.fixture(scope="module")
def library():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_header_and_binding_agree():
    assert declared_symbols() == set(_lib.SIGNATURES), "ctypes SIGNATURES must list every declared symbol"


def test_every_declared_symbol_is_exported(library):
    for name in sorted(declared_symbols()):
        assert hasattr(library, name), f"{name} declared in include/fairygen_b200.h but not exported"


def test_abi_version(library):
    assert library.fgb_abi_version() == 1


def test_create_fails_loudly_without_gpu(library):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    rc = library.fgb_create(0, ctypes.byref(handle))
    assert rc != 0 and not handle.value
    assert b"no CPU fallback" in library.fgb_last_error()


def test_null_context_is_an_error_not_a_crash(library):
    assert library.fgb_sync_check(None, None) != 0
    assert library.fgb_gemm_bf16(None, None, 0, None, 0, None, None, 0, 1, 8, 8, 0, None, None, 0, None) != 0
    assert b"ctx is NULL" in library.fgb_last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fairygen_b200/ may import, link or execute it."""
    for path in glob.glob(os.path.join(REPO, "fairygen_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
            text = open(path, errors="ignore").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
            assert "wan_dit_oracle" not in text, path


def test_engine_refuses_cpu():
    from fairygen_b200 import TI2V_5B, WanDiTEngine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WanDiTEngine(TI2V_5B, device="cpu")
