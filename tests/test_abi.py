"""CPU: the C-ABI library builds/loads and exports every symbol include/wlk_b200.h declares;
the product path fails loudly without a GPU (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "wlk_b200.h")).read()
    return sorted(set(re.findall(r"\b(wlk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from whisperlivekit_b200.build import build
    build()
    from whisperlivekit_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), s
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.wlk_abi_version() == 1


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from whisperlivekit_b200 import _lib
    from whisperlivekit_b200.dims import DIMS
    from whisperlivekit_b200.engine import WhisperEngine
    with pytest.raises(_lib.WlkError, match="no CPU fallback"):
        WhisperEngine(DIMS["micro"], None, [(0, 0)])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "whisperlivekit_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
