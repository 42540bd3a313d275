"""plfem_b200 — B200-native vectorial H-field P2 FEM mode solver.

Drop-in for the hot path of KhaoulaAguech/pl-fem-vectoriel's
``TrueVectorialMaxwellSolver`` (`solver_fem.py:113-239`): same Python API and
mode records, all numerical work in hand-written sm_100a CUDA behind the C ABI
declared in ``include/plfem.h``.
"""
from .geometry import MCFGeometry, PhotonicLanternGeometry, mcf_positions  # noqa: F401
from .config import SimulationConfig, PhysicalConstants, IPDipCauchy  # noqa: F401
from .mesh import MeshTri, MeshGenerator  # noqa: F401
from .losses import (LossCalculator, VectorialLossCalculator, EnhancedLossCalculator,  # noqa: F401
                     PhotonicLanternDesignParameters)

__all__ = ["MCFGeometry", "PhotonicLanternGeometry", "mcf_positions", "SimulationConfig",
           "PhysicalConstants", "IPDipCauchy", "MeshTri", "MeshGenerator",
           "LossCalculator", "VectorialLossCalculator", "EnhancedLossCalculator", "PhotonicLanternDesignParameters",
           "TrueVectorialMaxwellSolver", "ScalarHelmholtzSolver", "ModeRecord"]


def __getattr__(name):
    # the solver pulls in the CUDA library; keep geometry/mesh importable without it
    if name in ("TrueVectorialMaxwellSolver", "ModeRecord", "ScalarHelmholtzSolver"):
        from . import solver_fem
        return getattr(solver_fem, name)
    raise AttributeError(name)
