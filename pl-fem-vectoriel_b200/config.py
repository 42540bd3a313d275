"""Run-time knobs of the host boundary.

The reference imports ``SimulationConfig`` and ``PhysicalConstants`` from a
``config`` module that is absent from its checkout (`solver_fem.py:37`,
`mesh.py:41`); only the fields its code reads are re-created here
(`mesh.py:109,186,313-314`).  Their reference defaults are unknown, so the
defaults below are this repo's choice and are recorded with every benchmark:
``mesh_min_points = mesh_target_points = 0`` means the Delaunay mesh of the
point recipe is used as is (no ``refined()`` pass), which is the configuration
SURVEY.md §8(d) sizes.
"""
from dataclasses import dataclass


@dataclass
class SimulationConfig:
    enable_mesh_cache: bool = True
    cache_max_size: int = 150
    mesh_min_points: int = 0
    mesh_target_points: int = 0


class PhysicalConstants:
    N_SILICA = 1.4440
    N_POLYMER_BASE = 1.5200
    N_AIR = 1.0000
    C_UM_PER_S = 2.99792458e14


class IPDipCauchy:
    """IP-Dip dispersion, n(λ) = A + B/λ² + C/λ⁴ with λ in µm (`README.md:275`)."""
    A, B, C = 1.5259, 0.00860, 0.000210

    @classmethod
    def n(cls, wavelength_nm: float) -> float:
        lam = float(wavelength_nm) / 1000.0
        return cls.A + cls.B / lam ** 2 + cls.C / lam ** 4
