import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "reducedbasismethods.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the CUDA library (nvcc cross-compiles
    without a GPU) and the C oracle once, exactly as `__graft_entry__.build()` does, so that the suite is self-contained."""
    lib = os.path.join(PKG, "lib", "librbm_b200.so")
    orc = os.path.join(ROOT, "oracle", "_build", "libpic_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def ctx():
    """One rbm_ctx on cuda:0 for the whole GPU session (fails loudly without the library / a GPU)."""
    import rbm_b200

    c = rbm_b200.Context(0)
    rbm_b200.set_default_context(c)
    yield c
    rbm_b200.set_default_context(None)
    c.close()


def rel(a, b):
    import numpy as np

    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.max(np.abs(b)) if b.size else 0.0
    return float(np.max(np.abs(a - b)) / d) if d > 0 else float(np.max(np.abs(a - b))) if a.size else 0.0
