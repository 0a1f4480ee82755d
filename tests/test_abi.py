"""CPU: the C-ABI library loads and exports every symbol include/lcpb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lcpb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lcpb200_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lcp_physics_b200 import _lib, build
    build.build(verbose=False)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.EXPORTS)
    lib.lcpb200_version.restype = ctypes.c_int
    assert lib.lcpb200_version() == 200


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lcp_physics_b200 import LCPFunction
    from lcp_physics_b200.scenes import make_scenes
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LCPFunction()(*make_scenes(1, 4, 4, dtype=torch.float64))


def test_product_path_never_imports_oracle():
    pkg = os.path.join(ROOT, "lcp_physics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
