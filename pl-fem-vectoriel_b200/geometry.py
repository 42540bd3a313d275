"""Host-side geometry types consumed by the H-field mode solver.

This is the input type of the drop-in boundary, restated from the reference
(`geometry_unified.py:74-188` core layouts, `:195-363` MCFGeometry,
`:637-678` PhotonicLanternGeometry).  Nothing here runs on the GPU: a geometry
is ~20 numbers.  Both constructor surfaces of ``PhotonicLanternGeometry`` are
accepted:

* the code surface  ``PhotonicLanternGeometry(n_cores, arrangement,
  core_positions, core_radii, n_core, n_clad=1.0, ..., wavelength=1.55)``
  (`geometry_unified.py:642-647`), and
* the README surface ``PhotonicLanternGeometry(arrangement="hexagonal_1plus6_7",
  core_radius_um=1.5, pitch_um=8.0, n_core=1.535, n_clad=1.0,
  wavelength_nm=1550)`` (`README.md:141-148`).

Core coordinates must be bit-identical to the reference's (quadrature points
are tested against the core discs), so every layout is produced with the same
floating-point expression order; `tests/test_geometry.py` checks them against
values frozen from the reference module.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Optional, Tuple

import numpy as np

N_AIR = 1.0
PML_STRENGTH = 3.0
PML_ORDER = 2
PML_THICKNESS_UM = 10.0


def _ring(radius: float, degrees) -> np.ndarray:
    a = np.radians(degrees)
    return radius * np.column_stack([np.cos(a), np.sin(a)])


# layout table: n_cores -> (config_type, has_central, n_peripheral)
_LAYOUT_META = {
    1: ("single_1", True, 0),
    2: ("linear_2", False, 2),
    3: ("triangular_3", False, 3),
    4: ("square_2x2_4", False, 4),
    5: ("pentagonal_ring_5", False, 5),
    6: ("hexagonal_ring_6", False, 6),
    7: ("hexagonal_1plus6_7", True, 6),
    8: ("heptagonal_center_8", True, 7),
    9: ("square_3x3_9", True, 8),
    12: ("hex_double_ring_12", False, 12),
    13: ("hex_1plus6plus6_13", True, 12),
    19: ("hex_1plus6plus12_19", True, 18),
}
SUPPORTED_N = sorted(_LAYOUT_META)

#: arrangement name -> (n_cores, variant); the README passes arrangement names.
ARRANGEMENTS: Dict[str, Tuple[int, Optional[str]]] = {
    meta[0]: (n, None) for n, meta in _LAYOUT_META.items()
}
ARRANGEMENTS["pentagon_center_6"] = (6, "pentagon_center")


def mcf_positions(n_cores: int, pitch: float, variant: Optional[str] = None):
    """Core centres of the 12 published multicore layouts.

    Returns ``(positions (N,2), config_type, has_central_core, n_peripheral,
    R_ring)`` like `geometry_unified.py:74-188`.
    """
    p = float(pitch)
    centre = np.array([[0.0, 0.0]])
    hexa = np.arange(6) * 60
    if n_cores not in _LAYOUT_META:
        raise ValueError(f"n_cores={n_cores} non supporté. Valides : {SUPPORTED_N}")
    name, central, n_per = _LAYOUT_META[n_cores]

    if n_cores == 1:
        pos, r_ring = centre, 0.0
    elif n_cores == 2:
        pos, r_ring = np.array([[-p / 2, 0.0], [p / 2, 0.0]]), p / 2
    elif n_cores == 3:
        pos, r_ring = _ring(p, [90, 210, 330]), p
    elif n_cores == 4:
        h = p / 2
        pos, r_ring = np.array([[-h, -h], [h, -h], [-h, h], [h, h]]), h * np.sqrt(2)
    elif n_cores == 5:
        pos, r_ring = _ring(p, 90 + np.arange(5) * 72), p
    elif n_cores == 6 and variant == "pentagon_center":
        name, central, n_per = "pentagon_center_6", True, 5
        pos, r_ring = np.vstack([centre, _ring(p, 90 + np.arange(5) * 72)]), p
    elif n_cores == 6:
        pos, r_ring = _ring(p, hexa), p
    elif n_cores == 7:
        pos, r_ring = np.vstack([centre, _ring(p, hexa)]), p
    elif n_cores == 8:
        pos, r_ring = np.vstack([centre, _ring(p, np.arange(7) * (360 / 7))]), p
    elif n_cores == 9:
        c = [-p, 0.0, p]
        pos, r_ring = np.array([[x, y] for y in c for x in c]), p * np.sqrt(2)
    elif n_cores in (12, 13):
        rings = [_ring(p, hexa), _ring(p * np.sqrt(3), hexa + 30)]
        if n_cores == 13:
            rings.insert(0, centre)
        pos, r_ring = np.vstack(rings), p * np.sqrt(3)
    else:  # 19 = 1 + 6 + 6 (2p) + 6 (sqrt3 p, rotated 30 deg)
        pos = np.vstack([centre, _ring(p, hexa), _ring(2 * p, hexa),
                         _ring(p * np.sqrt(3), hexa + 30)])
        r_ring = 2 * p
    return pos, name, central, n_per, r_ring


class MCFGeometry:
    """Multicore-fibre cross-section (`geometry_unified.py:195-416`).

    Attributes read by the solver: ``positions (Nc,2)``, ``core_radii (Nc,)``,
    ``n_core``, ``n_clad``, ``k0``, ``epsilon(x, y)``; by the mesh recipe:
    ``domain_radius``, ``pml_thickness``, ``r_core``, ``use_complex_pml``.
    """

    SUPPORTED_N = SUPPORTED_N

    def __init__(self, n_cores, pitch_um, core_radius_um, n_core, n_clad=N_AIR,
                 wavelength_um=1.55, cladding_radius=None,
                 pml_thickness=PML_THICKNESS_UM, pml_strength=PML_STRENGTH,
                 pml_order=PML_ORDER, use_complex_pml=True,
                 taper_length_um=None, variant=None):
        self.n_cores = int(n_cores)
        self.n_core = float(n_core)
        self.n_clad = float(n_clad)
        self.delta_n = self.n_core - self.n_clad
        self.wavelength = float(wavelength_um)
        self.k0 = 2 * np.pi / self.wavelength
        if self.delta_n < 1e-6:
            raise ValueError(f"Δn={self.delta_n:.2e} trop faible")

        (self.positions, self.config_type, self.has_central_core,
         self.n_peripheral, self.R_ring) = mcf_positions(n_cores, pitch_um, variant)
        self.core_radii = np.full(self.n_cores, float(core_radius_um))
        self.core_positions = self.positions
        self.r_core = float(core_radius_um)
        self.V_number = self.k0 * self.r_core * np.sqrt(
            max(self.n_core ** 2 - self.n_clad ** 2, 0.0))

        if n_cores > 1:
            d = [np.linalg.norm(self.positions[i] - self.positions[j])
                 for i in range(n_cores) for j in range(i + 1, n_cores)]
            self.pitch = self.pitch_min = float(np.min(d))
            max_r = float(np.max(np.linalg.norm(self.positions, axis=1)))
        else:
            self.pitch = self.pitch_min = 0.0
            max_r = 0.0
        self.pitch_ratio = self.pitch / (2 * self.r_core) if self.r_core > 0 else 0.0

        self.cladding_radius = (cladding_radius if cladding_radius is not None
                                else max(max_r * 1.8 + self.r_core * 2, 20.0))
        self._domain_radius = max(max_r + self.r_core * 4,
                                  self.cladding_radius + pml_thickness * 1.2)
        self.pml_thickness = float(pml_thickness)
        self.pml_strength = float(pml_strength)
        self.pml_order = int(pml_order)
        self.use_complex_pml = bool(use_complex_pml)
        self.taper_length = taper_length_um

        area_c = n_cores * np.pi * self.r_core ** 2
        area_t = np.pi * (max_r + self.r_core) ** 2 if n_cores > 1 else area_c
        self.packing_efficiency = float(area_c / max(area_t, 1e-9))
        self._hash = self._compute_hash()

    @property
    def domain_radius(self) -> float:
        return self._domain_radius

    @property
    def hash(self) -> str:
        return self._hash

    def _compute_hash(self) -> str:
        h = hashlib.sha256()
        for chunk in (str(self.n_cores).encode(), self.positions.tobytes(),
                      self.core_radii.tobytes(),
                      f"{self.n_core:.6f}{self.n_clad:.6f}{self.wavelength:.6f}".encode(),
                      f"{self.cladding_radius:.4f}{self.pml_thickness:.2f}".encode(),
                      str(self.use_complex_pml).encode()):
            h.update(chunk)
        return h.hexdigest()[:20]

    def epsilon(self, x, y) -> np.ndarray:
        """Complex relative permittivity (`geometry_unified.py:325-347`).

        n_clad² everywhere, n_core² inside any core disc (closed, ``<=``), times
        ``1 + jσ(r)`` inside the PML annulus.  The solver only ever uses the
        real part (`solver_fem.py:132-150`), which the PML factor leaves
        untouched.
        """
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        eps = np.full_like(x, self.n_clad ** 2, dtype=np.complex128)
        for (cx, cy), r in zip(self.positions, self.core_radii):
            eps[(x - cx) ** 2 + (y - cy) ** 2 <= r ** 2] = self.n_core ** 2
        if self.use_complex_pml:
            rd = np.sqrt(x ** 2 + y ** 2)
            start = self._domain_radius - self.pml_thickness
            m = rd > start
            if np.any(m):
                rn = np.clip((rd[m] - start) / self.pml_thickness, 0.0, 1.0)
                eps[m] *= 1.0 + 1j * (self.pml_strength * rn ** self.pml_order)
        return eps

    def validate(self) -> Tuple[bool, str]:
        if self.delta_n < 5e-4:
            return False, f"Δn trop faible ({self.delta_n:.2e})"
        if self.V_number < 0.5:
            return False, f"V-number trop faible ({self.V_number:.2f})"
        if self.V_number > 20.0:
            return False, f"V-number très élevé ({self.V_number:.2f}) → multimode"
        for i in range(self.n_cores):
            for j in range(i + 1, self.n_cores):
                d = np.linalg.norm(self.positions[i] - self.positions[j])
                if d < (self.core_radii[i] + self.core_radii[j]) * 0.85:
                    return False, f"Chevauchement cœurs {i}↔{j}: d={d:.2f}µm"
        return True, "OK"

    def get_info(self) -> Dict:
        return dict(n_cores=self.n_cores, config_type=self.config_type,
                    has_central_core=self.has_central_core,
                    n_peripheral=self.n_peripheral, R_ring_um=float(self.R_ring),
                    pitch_um=float(self.pitch), pitch_ratio=float(self.pitch_ratio),
                    core_radius_um=float(self.r_core), n_core=self.n_core,
                    n_clad=self.n_clad, delta_n=float(self.delta_n),
                    V_number=float(self.V_number), wavelength_um=self.wavelength,
                    cladding_radius_um=float(self.cladding_radius),
                    domain_radius_um=float(self._domain_radius),
                    pml_thickness_um=float(self.pml_thickness),
                    packing_efficiency=float(self.packing_efficiency),
                    taper_length_um=self.taper_length, hash=self.hash)

    def __repr__(self) -> str:
        return (f"MCFGeometry(N={self.n_cores}, {self.config_type}, pitch={self.pitch:.1f}µm, "
                f"r={self.r_core:.2f}µm, V={self.V_number:.2f}, "
                f"n={self.n_core:.4f}/{self.n_clad:.4f})")


class PhotonicLanternGeometry(MCFGeometry):
    """Drop-in for the reference class of the same name, both signatures."""

    def __init__(self, n_cores=None, arrangement=None, core_positions=None,
                 core_radii=None, n_core=None, n_clad=1.0, cladding_radius=None,
                 wavelength=1.55, taper_length=None, pml_thickness=10.0,
                 pml_strength=3.0, pml_order=2, use_complex_pml=True,
                 core_radius_um=None, pitch_um=None, wavelength_nm=None, **kwargs):
        if n_core is None:
            raise TypeError("n_core is required")
        if wavelength_nm is not None:
            wavelength = float(wavelength_nm) / 1000.0

        if core_positions is None:
            # README surface: layout name (or n_cores) + uniform radius + pitch
            if arrangement is not None and arrangement in ARRANGEMENTS:
                n_arr, variant = ARRANGEMENTS[arrangement]
                if n_cores is not None and int(n_cores) != n_arr:
                    raise ValueError(f"arrangement {arrangement!r} has {n_arr} cores, got n_cores={n_cores}")
                n_cores = n_arr
            elif n_cores is not None:
                variant = kwargs.get("variant")
            else:
                raise ValueError(f"unknown arrangement {arrangement!r}; known: {sorted(ARRANGEMENTS)}")
            if core_radius_um is None or pitch_um is None:
                raise TypeError("core_radius_um and pitch_um are required with an arrangement name")
            super().__init__(n_cores, pitch_um, core_radius_um, n_core, n_clad, wavelength,
                             cladding_radius, pml_thickness, pml_strength, pml_order,
                             use_complex_pml, taper_length, variant)
            self.arrangement = self.config_type
            return

        # code surface (`geometry_unified.py:642-678`): explicit positions and radii
        pos = np.atleast_2d(np.asarray(core_positions, dtype=np.float64))
        if len(pos) > 1:
            pitch = float(np.min([np.linalg.norm(pos[i] - pos[j])
                                  for i in range(len(pos)) for j in range(i + 1, len(pos))]))
        else:
            pitch = float(np.max(core_radii)) * 4
        super().__init__(n_cores, pitch, float(np.mean(core_radii)), n_core, n_clad,
                         wavelength, cladding_radius, pml_thickness, pml_strength,
                         pml_order, use_complex_pml, taper_length)
        self.positions = self.core_positions = pos
        self.core_radii = np.asarray(core_radii, dtype=np.float64)
        self.arrangement = str(arrangement)


SAMPLING_WEIGHTS: Dict[int, float] = {2: 0.04, 3: 0.11, 4: 0.13, 5: 0.05, 6: 0.10, 7: 0.30,
                                      8: 0.05, 9: 0.08, 12: 0.07, 13: 0.07, 19: 0.10}


def has_disc_epsilon(geometry) -> bool:
    """True when ``geometry.epsilon`` is the stock disc model above, so the
    device kernel may evaluate it itself from (positions, radii, n_core, n_clad)
    instead of receiving host-evaluated samples."""
    fn = getattr(type(geometry), "epsilon", None)
    return fn is MCFGeometry.epsilon
