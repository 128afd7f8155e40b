"""The C-ABI library loads and exports every symbol include/pomp_b200.h declares; without a GPU
every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from pyoptimalmotionplanning_b200 import _abi
from pyoptimalmotionplanning_b200.descriptors import MetricSpec, SpaceSpec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pomp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pomp_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    names = _declared()
    assert len(names) >= 40
    assert set(names) == set(_abi.SIGNATURES), set(names) ^ set(_abi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.pomp_version() >= 100
    assert lib.pomp_launch_count() >= 0


def test_struct_layouts_match_the_header():
    # sizes implied by the header on LP64
    assert C.sizeof(_abi.MetricDesc) == 4 * 3 + 4 * 8 * 2 + 4 + 8 * 8 * 2 + 8 + 4 * 4 + 8 * 2
    assert C.sizeof(_abi.SpaceDesc) == 4 * 6 + 8 * 8 * 2 + 8 * 2
    assert C.sizeof(_abi.DynDesc) == 4 * 4 + 8 + 8 * 8 + 8 * 64 * 2


def test_compute_fails_loudly_without_a_gpu():
    lib = _abi.load_library()
    if lib.pomp_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.pomp_nn_create(C.byref(MetricSpec.euclidean(2).desc()), 0, 0, C.byref(h))
    assert rc == _abi.POMP_ERR_CUDA
    assert b"no CPU fallback" in lib.pomp_last_error()
    d = SpaceSpec([0, 0], [1, 1]).desc()
    rc = lib.pomp_space_create(C.byref(d), 0, C.byref(h))
    assert rc == _abi.POMP_ERR_CUDA
    from pyoptimalmotionplanning_b200 import NearestNeighbors, PompError
    nn = NearestNeighbors(MetricSpec.euclidean(2))
    with pytest.raises(PompError):
        nn.add([0.1, 0.2])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyoptimalmotionplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "pomp_oracle" not in txt, f
