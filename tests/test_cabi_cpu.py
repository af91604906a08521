"""CPU-side checks of the product boundary: the library builds, loads and exports every symbol of include/mz_b200.h;
without a GPU every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

import oracle_lib

ROOT = oracle_lib.ROOT


def test_library_exports_every_declared_symbol():
    import minizero_b200
    path = minizero_b200.build_library()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "mz_b200.h")).read()
    declared = set(re.findall(r"\b(mz_[a-z_0-9]+)\s*\(", header))
    assert declared == set(minizero_b200.engine.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import minizero_b200
    with pytest.raises(minizero_b200.EngineError, match="no CUDA device|CUDA"):
        minizero_b200.Engine(minizero_b200.GAME_TICTACTOE, 3, 1, 10)


def test_product_does_not_reference_oracle():
    """the product tree must not import, link or execute anything under oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "minizero_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "liboracle" not in text and "oracle_lib" not in text and "mzo_" not in text, os.path.join(base, f)
