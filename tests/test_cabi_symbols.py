"""The C-ABI library loads and exports every symbol include/dd_engine.h declares; without a GPU the product
path fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest
import torch

import diffusiondepth_b200 as dd
from diffusiondepth_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dd_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    lib = dd.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in dd_engine.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_cabi.SIGNATURES) == set(names)
    assert lib.dd_abi_version() == _cabi.ABI_VERSION


def test_no_torch_types_in_abi():
    src = open(os.path.join(ROOT, "include", "dd_engine.h")).read()
    assert "torch" not in src.lower().replace("pytorch's allocator", "").replace("pytorch", "")
    assert "at::" not in src and "c10::" not in src


def test_config_struct_layout_matches_header():
    assert C.sizeof(_cabi.DDConfig) == 10 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_fails_loudly_without_gpu():
    lib = dd.load_library()
    cfg = _cabi.DDConfig(_cabi.ABI_VERSION, 1, 1, 8, 16, 4, 8, 2, 0, 0)
    h = C.c_void_p()
    rc = lib.dd_create(C.byref(cfg), C.byref(h))
    assert rc == 3, "dd_create must report DD_ERR_UNSUPPORTED when there is no device"
    assert b"no CPU path" in lib.dd_last_error()
    with pytest.raises(dd.EngineError):
        dd.DenoiseEngine("swin", 1, (8, 16), (4, 8), 2, torch.device("cpu"))


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_cabi, "_LIB", None)
    monkeypatch.setenv("DD_ENGINE_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(dd.EngineError):
        _cabi.load_library()


def test_bad_arguments_rejected():
    lib = dd.load_library()
    h = C.c_void_p()
    cfg = _cabi.DDConfig(99, 1, 1, 8, 16, 4, 8, 2, 0, 0)
    assert lib.dd_create(C.byref(cfg), C.byref(h)) == 1
    cfg = _cabi.DDConfig(_cabi.ABI_VERSION, 0, 1, 8, 16, 4, 8, 2, 0, 0)  # Res variant needs cond == latent size
    assert lib.dd_create(C.byref(cfg), C.byref(h)) == 1
    assert lib.dd_create(None, C.byref(h)) == 1
