"""``TrueVectorialMaxwellSolver`` — drop-in for the reference class (`solver_fem.py:113-239`).

Same constructor, methods, return types and mode-record keys; the arithmetic
the reference delegates to scikit-fem (`asm`), ``scipy.sparse`` (block build,
Dirichlet slicing) and ``scipy.sparse.linalg.eigsh`` (ARPACK + SuperLU) runs in
the CUDA library behind ``include/plfem.h``.  There is no CPU path: without a
B200-class GPU and the built ``libplfem.so`` every call raises.

Host work kept in Python (a few dozen scalars per solve): the LP01 shift
estimate (`:187-193`), the window / divergence / radiation filters and the final
sort (`:206-210`, `:228-239`).

Both surfaces are offered:

* code surface   ``TrueVectorialMaxwellSolver(geometry, use_pml=False)``
  + ``assemble_hfield_system(mesh)`` + ``solve_vectorial_modes(mesh, n_modes_target)``;
* README surface ``TrueVectorialMaxwellSolver(geom, n_modes=10).solve()`` whose
  records also answer ``mode.n_eff``, ``mode.confinement``, ``mode.PDL_dB``,
  ``mode.polarization_state`` (`README.md:151-159`).
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np

from . import _cabi
from .config import SimulationConfig
from .geometry import has_disc_epsilon
from .mesh import MeshGenerator

logger = logging.getLogger("pl_v18.solver_fem")


class ModeRecord(dict):
    """Mode dict of the reference with attribute access and the README alias."""
    _ALIASES = {"polarization_state": "polarization", "neff": "n_eff"}

    def __getattr__(self, name):
        key = self._ALIASES.get(name, name)
        try:
            return self[key]
        except KeyError:
            raise AttributeError(name) from None


class P2BasisView:
    """What the reference's callers read from the scikit-fem ``Basis`` returned by
    ``assemble_hfield_system``: ``N``, ``doflocs``, ``element_dofs``, ``get_dofs().all()``."""

    class _Dofs:
        def __init__(self, ids):
            self._ids = ids

        def all(self):
            return self._ids

    def __init__(self, mesh, element_dofs, doflocs, boundary, interior):
        self.mesh = mesh
        self.element_dofs = element_dofs
        self.doflocs = doflocs
        self.N = doflocs.shape[1]
        self._boundary = boundary
        self.interior_dofs = interior

    def get_dofs(self):
        return self._Dofs(self._boundary)


def sigma_estimate(geometry) -> float:
    """Shift-invert target from the LP01 b-V approximation (`solver_fem.py:187-193`)."""
    n_core, n_clad, k0 = geometry.n_core, geometry.n_clad, geometry.k0
    NA = np.sqrt(max(n_core ** 2 - n_clad ** 2, 1e-6))
    V_geom = k0 * np.mean(geometry.core_radii) * NA
    b_approx = max((1.0 - 2.405 / max(V_geom, 2.41)) ** 2, 0.05)
    n_eff_est = np.sqrt(n_clad ** 2 + b_approx * (n_core ** 2 - n_clad ** 2))
    return (k0 * float(np.clip(n_eff_est, n_clad + 0.05, n_core - 0.005))) ** 2


def _polarization_label(P_x: float, P_y: float):
    """`solver_fem.py:97-105`."""
    ratio = P_x / P_y
    PDL = float(np.clip(10.0 * np.log10(max(P_x, P_y) / min(P_x, P_y)), 0.0, 50.0))
    if ratio > 10.0:
        pol = "TE-like"
    elif ratio > 2.5:
        pol = "HE-like"
    elif ratio > 0.4:
        pol = "Hybrid"
    elif ratio > 0.1:
        pol = "EH-like"
    else:
        pol = "TM-like"
    return pol, PDL


def modes_from_solution(geo, N_solve: int, beta_sq, evecs, met, n_core_dofs: int):
    """Window filter, per-mode records, divergence / radiation filters and sort (`solver_fem.py:200-239`) from
    the eigenvalues, the l2-normalised eigenvectors (may be None: records then carry no DOF arrays) and the
    (k, 8) reductions the device computed.  Returns (modes_guided, modes_raw, frac_core)."""
    frac_core = n_core_dofs / N_solve
    n_core, n_clad, k0 = geo.n_core, geo.n_clad, geo.k0
    modes_raw = []
    for i in range(len(beta_sq)):
        b2 = beta_sq[i]
        if b2 <= 0:
            continue
        beta = np.sqrt(b2)
        ne = beta / k0
        if ne <= n_clad or ne >= n_core * 1.01:
            continue
        div_energy, e_core, e_all, px_c, py_c, px_a, py_a, _ = met[i]
        use_core = n_core_dofs > 0                       # `solver_fem.py:93`
        P_x = float(px_c if use_core else px_a) + 1e-30
        P_y = float(py_c if use_core else py_a) + 1e-30
        pol, PDL_dB = _polarization_label(P_x, P_y)
        conf = float(e_core / e_all)
        modes_raw.append(ModeRecord({
            "n_eff": float(ne), "beta": float(beta),
            # rows of a buffer this solve owns (fresh per call, like the reference's eigsh output)
            "Ex_dofs": evecs[i, :N_solve] if evecs is not None else None,
            "Ey_dofs": evecs[i, N_solve:] if evecs is not None else None,
            "P_x": P_x, "P_y": P_y, "PDL_dB": PDL_dB, "polarization": pol,
            "confinement": conf, "core_overlap": conf,
            "div_ratio": float(div_energy) / max(b2, 1e-12),
            "is_vectorial": True, "method": "H-field_V18.10"}))

    dr = np.array([m["div_ratio"] for m in modes_raw])
    dr_thresh = max(np.median(dr) * 10, dr.min() * 50, 1e-6)      # raises on empty input, like the reference
    modes_phys = [m for m in modes_raw if m["div_ratio"] <= dr_thresh]
    conf_thr = max(5.0 * frac_core, 0.05)
    modes_guided = [m for m in modes_phys if m["confinement"] >= conf_thr] or modes_phys
    modes_guided.sort(key=lambda m: m["n_eff"], reverse=True)
    return modes_guided, modes_raw, frac_core


class TrueVectorialMaxwellSolver:
    def __init__(self, geometry, use_pml: bool = False, n_modes: Optional[int] = None,
                 device: int = 0, refinement: float = 1.0, config: Optional[SimulationConfig] = None, ctx=None):
        _cabi.load()                          # fails loudly when the CUDA library cannot be built/loaded
        self.geometry = geometry
        self.k0 = geometry.k0
        self.use_pml = use_pml                # stored, never read — like the reference (`solver_fem.py:119`)
        self.n_modes = n_modes
        self.device = int(device)
        self._ctx = ctx                       # explicit C-ABI context (one per host thread in a SolverPool)
        self.refinement = refinement
        self.config = config
        self.last_stats: Dict = {}
        self._problems: Dict[int, tuple] = {}  # id(mesh) -> (mesh, Problem): DOF tables/plan reused per mesh
        logger.info(f"Solveur H-field initialisé - k₀={self.k0:.4f} µm⁻¹")

    # ------------------------------------------------------------------ plumbing
    def _problem(self, mesh) -> "_cabi.Problem":
        ent = self._problems.get(id(mesh))
        if ent is None or ent[0] is not mesh:
            if len(self._problems) >= 4:
                self._problems.pop(next(iter(self._problems)))[1].close()
            ent = (mesh, _cabi.Problem(mesh, self._ctx or _cabi.Context.get(self.device)))
            self._problems[id(mesh)] = ent
        return ent[1]

    def _material(self, pb, alpha_p: float = 1.0):
        eps_q = None
        if not has_disc_epsilon(self.geometry):
            # custom epsilon(x, y): sample it on the host at the quadrature points, once
            xy = pb.quad_points()
            eps_q = np.real(self.geometry.epsilon(xy[0], xy[1]))
        return _cabi.material_struct(self.geometry, alpha_p, eps_q)

    # ------------------------------------------------------------------ code surface
    def assemble_hfield_system(self, mesh):
        """`solver_fem.py:122-169` -> ``(A, B, basis, Dxx, Dyy, Dxy, M_inv)`` as SciPy CSR."""
        pb = self._problem(mesh)
        mat, _keep = self._material(pb)
        pb.assemble(mat)
        A, B = pb.export_csr("A"), pb.export_csr("B")
        Dxx, Dyy, Dxy, M_inv = (pb.export_csr(k) for k in ("Dxx", "Dyy", "Dxy", "M_inv"))
        ed, loc, bnd, itr = pb.dofs()
        basis = P2BasisView(mesh, ed, loc, bnd, itr)
        logger.info(f"Assemblage terminé - {basis.N} DOFs P2, matrice {2 * basis.N}×{2 * basis.N}")
        return A, B, basis, Dxx, Dyy, Dxy, M_inv

    def solve_vectorial_modes(self, mesh, n_modes_target: int = 20, v0=None, return_raw: bool = False,
                              **solver_opts) -> List[Dict]:
        """`solver_fem.py:171-239`: modes of [A]{Ht} = β²[B]{Ht} nearest the LP01 shift."""
        pb = self._problem(mesh)
        mat, _keep = self._material(pb)
        geo = self.geometry
        N_solve = pb.n_interior
        sigma = sigma_estimate(geo)
        n_req = min(n_modes_target + 12, 2 * N_solve - 4)
        beta_sq, evecs, met, n_core_dofs, stats = pb.solve_modes(mat, sigma, n_req, tol=_cabi.EIG_TOL, maxiter=12000, v0=v0,
                                                                 **solver_opts)
        self.last_stats = stats.as_dict()
        modes_guided, modes_raw, frac_core = modes_from_solution(geo, N_solve, beta_sq, evecs, met, n_core_dofs)
        if return_raw:
            return modes_guided, dict(beta_sq=beta_sq, evecs=evecs, metrics=met, sigma=sigma,
                                      modes_raw=modes_raw, frac_core=frac_core, stats=self.last_stats)
        return modes_guided

    # ------------------------------------------------------------------ README surface
    def solve(self, mesh=None, n_modes: Optional[int] = None) -> List[ModeRecord]:
        """``TrueVectorialMaxwellSolver(geom, n_modes=10).solve()`` (`README.md:151-152`)."""
        if mesh is None:
            mesh, _ = MeshGenerator.generate(self.geometry, self.refinement, self.config)
        n = n_modes if n_modes is not None else (self.n_modes if self.n_modes is not None else 20)
        return self.solve_vectorial_modes(mesh, n_modes_target=n)

    def close(self):
        for _, pb in self._problems.values():
            pb.close()
        self._problems.clear()


class ScalarHelmholtzSolver:
    """Drop-in for the reference's scalar P2 solver (`solver_fem.py:245-276`): modes of
    ``(K - k0^2 M_eps) v = lambda M v`` nearest ``sigma = -(k0 (n_core - 0.008))^2``, ``n_eff = sqrt(-lambda) / k0``.

    Runs on the same CUDA kernels as the vectorial solver: the scalar pencil occupies the Hx block of the two-unknowns-per-node
    layout (material ``scalar_mode``), the Hy block carries ``(shift M, M)`` whose eigenvalues all equal ``shift``, far from
    ``sigma``, so it never enters the wanted set; no boundary DOF is eliminated (the reference does not call ``get_dofs``).
    Records carry the reference's keys: ``n_eff, beta, field_vector, confinement, core_overlap, PDL_dB = 0,
    polarization = 'scalar', is_vectorial = False``; ``field_vector`` is M-normalised over all N DOFs."""

    def __init__(self, geometry, device: int = 0, ctx=None):
        _cabi.load()
        self.geometry = geometry
        self.k0 = geometry.k0
        self.device = int(device)
        self._ctx = ctx
        self.last_stats: Dict = {}

    def solve(self, mesh, n_modes_target: int = 20) -> List[Dict]:
        geo = self.geometry
        pb = _cabi.Problem(mesh, self._ctx or _cabi.Context.get(self.device))
        try:
            eps_q = None
            if not has_disc_epsilon(geo):
                xy = pb.quad_points()
                eps_q = np.real(geo.epsilon(xy[0], xy[1]))
            mat, _keep = _cabi.material_struct(geo, 1.0, eps_q)
            pb.assemble(mat)
            M = pb.export_csr("M")                                      # plain mass matrix, all N DOFs (for the M-norm)
            _, loc, _, _ = pb.dofs()
            pb.set_dirichlet(False)
            N = pb.N
            sigma = -(self.k0 * (geo.n_core - 0.008)) ** 2
            mat.scalar_mode = 1
            mat.scalar_shift = 1.0e3 * max(1.0, abs(sigma))
            k = min(n_modes_target + 8, N - 4)
            evals, evecs, _met, _nc, stats = pb.solve_modes(mat, sigma, k, tol=_cabi.EIG_TOL, maxiter=6000)
            self.last_stats = stats.as_dict()
        finally:
            pb.close()
        x_dof, y_dof = loc
        in_core = np.zeros(N, dtype=bool)
        for (cx, cy), r in zip(geo.positions, geo.core_radii):
            in_core |= (x_dof - cx) ** 2 + (y_dof - cy) ** 2 <= r ** 2
        modes = []
        for lam, vec in zip(evals, evecs):
            if lam >= 0:
                continue
            ne = np.sqrt(-lam) / self.k0
            if ne <= geo.n_clad or ne >= geo.n_core * 1.005:
                continue
            v = np.array(vec[:N])
            v /= np.sqrt(float(v @ (M @ v))) + 1e-30
            conf = float(np.sum(v[in_core] ** 2) / np.sum(v ** 2))
            modes.append(ModeRecord({"n_eff": float(ne), "beta": float(self.k0 * ne), "field_vector": v, "confinement": conf,
                                     "core_overlap": conf, "PDL_dB": 0.0, "polarization": "scalar", "is_vectorial": False}))
        modes.sort(key=lambda m: m["n_eff"], reverse=True)
        return modes
