"""CPU: libb200tts.so loads without a GPU and exports exactly the symbols include/b200tts.h declares."""
import os
import re

import pytest

import b200tts  # noqa: F401
from b200tts import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200tts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200tts_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(capi.SIGNATURES)


def test_library_exports_every_symbol():
    lib = capi.load_library()
    for name in _header_symbols():
        assert hasattr(lib, name), name


def test_error_path_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        capi.Engine(0)
    assert capi.load_library().b200tts_launch_count() == 0
