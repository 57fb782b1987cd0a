import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("SUBG_QUIET", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (this container only)")


def _gpu_ready() -> str:
    """'' when the gpu-marked tests can run here, else the reason they are skipped."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device"
    except Exception as ex:  # pragma: no cover
        return f"torch unavailable: {ex!r}"
    from surel_plus_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        return f"{_capi.LIB_PATH} not built (python -m surel_plus_b200.build)"
    return ""


def pytest_collection_modifyitems(config, items):
    from oracle import reference as ref
    have_tree = ref.have_reference_tree()
    skip_ref = pytest.mark.skip(reason="reference tree not present (GPU box)")
    why = None
    for item in items:
        if "reference" in item.keywords and not have_tree:
            item.add_marker(skip_ref)
        if "gpu" in item.keywords:
            if why is None:
                why = _gpu_ready()
            if why:  # a plain `pytest tests` on a CPU host skips the GPU suite instead of failing it
                item.add_marker(pytest.mark.skip(reason=why))


@pytest.fixture(scope="session")
def small_graph():
    """600-node heavy-tailed graph with hubs (deg > M) and 5 isolated nodes."""
    from surel_plus_b200.graphs import synthetic_graph
    return synthetic_graph(605, 3000, seed=1, gamma=3.0, isolated=5)


@pytest.fixture(scope="session")
def mid_graph():
    from surel_plus_b200.graphs import synthetic_graph
    return synthetic_graph(20_000, 120_000, seed=7, gamma=2.0, isolated=3)
