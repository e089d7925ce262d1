"""CPU checks of the drop-in boundary: librbffd.so loads, exports every symbol include/rbffd.h declares, the Python
mirror binds exactly those, and there is no CPU fallback."""
import ctypes
import os
import re

import pytest

import rbffd_b200 as rb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rbffd.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbffd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    names = _header_functions()
    assert len(names) >= 25
    L = ctypes.CDLL(rb._lib.LIB_PATH)
    for nm in names:
        assert hasattr(L, nm), f"{nm} declared in include/rbffd.h but not exported by librbffd.so"
    assert sorted(rb.exported_symbols()) == names


def test_option_struct_layout():
    assert ctypes.sizeof(rb.Options) == 4 * (5 + 4 * 12 + 3 + 5)
    assert ctypes.sizeof(rb.AdvDiffParams) == 8 * 4 + 4 * 8
    o = rb.make_options(2, 5, 30, 3, ["Lap", "Dxx", ("Dk", 1, 4)])
    assert [list(o.ops[i]) for i in range(3)] == [[1, 0, 0, 0], [0, 2, 0, 0], [0, 0, 4, 0]]
    with pytest.raises(ValueError):
        rb.make_options(2, 5, 30, 3, ["Dz"])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(rb.RbffdError) as e:
        rb.Context(0)
    assert e.value.code == rb._lib.ERR_CUDA
    import numpy as np
    with pytest.raises(rb.RbffdError):
        rb.generate_operator(np.random.rand(100, 2), None, 3, 12, 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "radialbasisfinitedifferences.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"liboracle", r"\borc_[a-z]", r"oracle[/.]oracle"):
                    assert not re.search(pat, txt), f"{f} reaches into oracle/ ({pat})"
