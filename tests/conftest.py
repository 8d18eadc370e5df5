import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _has_gpu() -> bool:
    try:
        import ctypes
        from turdb_b200 import _lib
        c = ctypes.c_int32(0)
        return _lib.load().turdb_cuda_device_count(ctypes.byref(c)) == 0 and c.value > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_required():
    if not _has_gpu():
        pytest.skip("no CUDA device")
    return True


@pytest.fixture(scope="session")
def small_graph():
    """10k x 128 latent-Gaussian corpus, graph built by the oracle's reference-intent insert path."""
    import numpy as np
    from oracle import binding as ob
    from turdb_b200 import datasets as ds
    x = ds.gaussian_latent(10_000, 128, seed=1)
    g = ob.OracleGraph.build(x, m=16, ef_construction=100, mode=ob.BUILD_INTENT, seed=42,
                             row_ids=np.arange(10_000, dtype=np.uint64) * 3 + 7)
    return g, g.export()
