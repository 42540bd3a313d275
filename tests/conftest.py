import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def cfg1():
    """Config 1 of BASELINE.json: 7-core hexagonal PL, r=1.5 um, pitch 8 um, IP-Dip/air, 1550 nm."""
    import plfem_b200 as P
    g = P.PhotonicLanternGeometry(arrangement="hexagonal_1plus6_7", core_radius_um=1.5, pitch_um=8.0,
                                  n_core=1.535, n_clad=1.0, wavelength_nm=1550)
    mesh, _ = P.MeshGenerator.generate(g)
    return g, mesh


@pytest.fixture(scope="session")
def small_case():
    """3-core lantern on a coarse mesh: seconds for the oracle, exercises every code path."""
    import plfem_b200 as P
    g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55)
    mesh, _ = P.MeshGenerator.generate(g, refinement=0.4)
    return g, mesh


@pytest.fixture(scope="session")
def cfg2():
    """Config 2 of BASELINE.json: 19-core MCF cross-section, C-band (Cauchy IP-Dip index at 1550 nm), n_modes = 40."""
    import plfem_b200 as P
    g = P.MCFGeometry(19, 8.0, 1.5, P.IPDipCauchy.n(1550), 1.0, 1.55)
    mesh, _ = P.MeshGenerator.generate(g)
    return g, mesh
