"""CPU: the C-ABI library loads and exports every symbol include/subg_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "subg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(subg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from surel_plus_b200 import build, _capi
    build.build()
    L = _capi.load()
    decl = _declared()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/subg_b200.h but not exported"
    assert sorted(_capi.SYMBOLS) == decl
    assert L.subg_abi_version() == 4


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path raises instead of computing on the CPU."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from surel_plus_b200 import subg_acc
    with pytest.raises(Exception):
        subg_acc.gset_sampler(np.array([0, 1, 2], np.int32), np.array([1, 0], np.int32), np.arange(2))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "surel_plus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("pinned oracle", ""), f"{f} mentions oracle/"
