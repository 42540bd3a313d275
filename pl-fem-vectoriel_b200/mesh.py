"""Triangular mesh of a lantern cross-section — the step just before the hot path.

``MeshGenerator`` restates the point recipe of the reference
(`mesh.py:223-340`): Cartesian base grid, two polar ring families per core,
one annular family in the PML, clip to 1.01·R, round to 8 decimals, unique,
Qhull Delaunay with options ``QJ Pp``.  ``MeshTri`` is the minimal stand-in
for scikit-fem's class of the same name: it holds ``p (2,V) float64`` and
``t (3,T)`` with every column sorted ascending (scikit-fem's ``sort_t``
normalisation), which is all the solver reads
(`mesh.py:308`, `solver_fem.py:126`).

This is host code: O(V log V) once per geometry, and independent of
wavelength and n_core, so a band sweep re-uses one mesh (class-level cache
keyed like `mesh.py:132-165`).
"""
from __future__ import annotations

import hashlib
import threading
from collections import OrderedDict
from typing import Optional

import numpy as np
from scipy.spatial import Delaunay

from .config import SimulationConfig


class MeshTri:
    def __init__(self, p, t, sort_t: bool = True):
        self.p = np.ascontiguousarray(p, dtype=np.float64)
        t = np.asarray(t, dtype=np.int64)
        self.t = np.ascontiguousarray(np.sort(t, axis=0) if sort_t else t)
        if self.p.shape[0] != 2 or self.t.shape[0] != 3:
            raise ValueError("MeshTri expects p (2,V) and t (3,T)")

    def _edges(self):
        t = self.t
        e = np.hstack([t[[0, 1]], t[[1, 2]], t[[0, 2]]])          # already (min,max) since t is sorted
        facets, inv = np.unique(e, axis=1, return_inverse=True)
        return facets, np.asarray(inv).reshape(3, t.shape[1])

    def refined(self, times: int = 1) -> "MeshTri":
        """Uniform red refinement (each triangle -> 4), [skfem-recall] MeshTri1._uniform."""
        m = self
        for _ in range(int(times)):
            p, t = m.p, m.t
            facets, t2f = m._edges()
            mid = t2f + p.shape[1]
            newp = np.hstack([p, p[:, facets].mean(axis=1)])
            newt = np.hstack([np.vstack([t[0], mid[0], mid[2]]),
                              np.vstack([t[1], mid[0], mid[1]]),
                              np.vstack([t[2], mid[2], mid[1]]),
                              np.vstack([mid[0], mid[1], mid[2]])])
            m = MeshTri(newp, newt)
        return m

    @classmethod
    def init_structured(cls, nx: int, ny: int, half_width: float) -> "MeshTri":
        """(nx × ny) cells on [-half_width, half_width]², every cell split on the same
        diagonal — the synthetic stress mesh of SURVEY.md §8(d) config 5."""
        xs = np.linspace(-half_width, half_width, nx + 1)
        ys = np.linspace(-half_width, half_width, ny + 1)
        X, Y = np.meshgrid(xs, ys)
        p = np.vstack([X.ravel(), Y.ravel()])
        i, j = np.meshgrid(np.arange(nx), np.arange(ny))
        v0 = (j * (nx + 1) + i).ravel()
        t = np.hstack([np.vstack([v0, v0 + 1, v0 + nx + 2]),
                       np.vstack([v0, v0 + nx + 1, v0 + nx + 2])])
        return cls(p, t)


def lantern_point_cloud(geometry, refinement: float = 1.0) -> np.ndarray:
    """The (2, V) vertex cloud of `mesh.py:232-297`, same expression order."""
    R = geometry.domain_radius
    n_base = max(int(25 + 20 * refinement), 16)
    ax = np.linspace(-R, R, n_base, dtype=np.float64)
    X, Y = np.meshgrid(ax, ax)
    chunks = [np.vstack([X.ravel(), Y.ravel()])]

    theta = np.linspace(0, 2 * np.pi, max(int(16 * refinement), 12), endpoint=False)
    positions = np.atleast_2d(np.asarray(
        getattr(geometry, "positions", getattr(geometry, "core_positions", np.zeros((1, 2))))))
    for (cx, cy), r in zip(positions, np.asarray(geometry.core_radii)):
        for radii in (np.linspace(0, r * 0.95, max(int(14 * refinement), 10)),
                      np.linspace(r * 0.90, r * 1.20, max(int(18 * refinement), 14))):
            Rg, Tg = np.meshgrid(radii, theta)
            chunks.append(np.vstack([cx + Rg.ravel() * np.cos(Tg.ravel()),
                                     cy + Rg.ravel() * np.sin(Tg.ravel())]))

    pml_start = R - geometry.pml_thickness * 1.1
    if pml_start > 0:
        th = np.linspace(0, 2 * np.pi, max(int(36 * refinement), 24), endpoint=False)
        Rg, Tg = np.meshgrid(np.linspace(pml_start, R * 0.98, max(int(18 * refinement), 12)), th)
        chunks.append(np.vstack([Rg.ravel() * np.cos(Tg.ravel()), Rg.ravel() * np.sin(Tg.ravel())]))

    pts = np.hstack(chunks)
    pts = pts[:, np.linalg.norm(pts, axis=0) <= R * 1.01]
    pts = np.round(pts.T, decimals=8).T
    return np.unique(pts, axis=1)


def signed_double_area(p: np.ndarray, t: np.ndarray) -> np.ndarray:
    a = p[:, t[1]] - p[:, t[0]]
    b = p[:, t[2]] - p[:, t[0]]
    return a[0] * b[1] - a[1] * b[0]


def drop_flat_triangles(mesh: MeshTri) -> MeshTri:
    """Remove triangles whose three vertices are exactly collinear.

    Deviation from the reference recipe, on purpose: ``Delaunay(..., 'QJ Pp')``
    joggles the input, so collinear points on the convex hull of the clipped
    base grid come back as zero-area triangles (25 of them for the 7-core
    config).  Their affine map is singular — scikit-fem divides by det J = 0 and
    the reference's system fills with NaN — while their area is zero, so
    removing them changes no integral.  Slivers with a tiny non-zero area are
    kept, as the reference keeps them.
    """
    keep = signed_double_area(mesh.p, mesh.t) != 0.0
    return mesh if keep.all() else MeshTri(mesh.p, mesh.t[:, keep])


class MeshGenerator:
    """``MeshGenerator.generate(geometry, refinement, config) -> (mesh, basis)``.

    ``basis`` is ``None`` here: the reference returns a scikit-fem ``Basis``
    that its solver rebuilds anyway (`solver_fem.py:126`); the DOF tables of
    this build live on the device side of the C ABI.
    """
    _cache: "OrderedDict[str, tuple]" = OrderedDict()
    _cache_hits = 0
    _cache_misses = 0
    _cache_lock = threading.Lock()      # the sweep driver meshes on several threads (the reference's cache is not thread-safe)
    MAX_REFINEMENT_ITERATIONS = 5

    @classmethod
    def _key(cls, geometry, refinement: float) -> str:
        h = hashlib.sha256()
        # the mesh does not depend on wavelength or indices: key on shape only
        h.update(np.asarray(geometry.positions, dtype=np.float64).tobytes())
        h.update(np.asarray(geometry.core_radii, dtype=np.float64).tobytes())
        h.update(f"{geometry.domain_radius:.10f}{geometry.pml_thickness:.4f}{refinement:.4f}".encode())
        return h.hexdigest()[:24]

    @classmethod
    def generate(cls, geometry, refinement: float = 1.0,
                 config: Optional[SimulationConfig] = None):
        config = config or SimulationConfig()
        key = cls._key(geometry, refinement) + f"{config.mesh_min_points}:{config.mesh_target_points}"
        with cls._cache_lock:
            if config.enable_mesh_cache and key in cls._cache:
                cls._cache_hits += 1
                cls._cache.move_to_end(key)
                return cls._cache[key]
            cls._cache_misses += 1
        out = cls._generate_mesh(geometry, refinement, config)
        if config.enable_mesh_cache:
            with cls._cache_lock:
                while len(cls._cache) >= max(config.cache_max_size, 1):
                    cls._cache.popitem(last=False)
                cls._cache[key] = out
        return out

    @classmethod
    def _generate_mesh(cls, geometry, refinement: float, config: SimulationConfig):
        mesh = drop_flat_triangles(cls._delaunay_mesh(geometry, refinement))
        it = 0
        while mesh.p.shape[1] < config.mesh_min_points and it < cls.MAX_REFINEMENT_ITERATIONS:
            mesh = mesh.refined()
            it += 1
            if mesh.p.shape[1] > config.mesh_target_points * 2.5:
                break
        return mesh, None

    @staticmethod
    def _delaunay_mesh(geometry, refinement: float) -> MeshTri:
        """The reference's mesh before any clean-up (`mesh.py:232-308`)."""
        pts = lantern_point_cloud(geometry, refinement)
        tri = Delaunay(pts.T, qhull_options="QJ Pp")
        return MeshTri(tri.points.T, tri.simplices.T)

    @classmethod
    def clear_cache(cls):
        cls._cache.clear()
        cls._cache_hits = cls._cache_misses = 0

    @classmethod
    def _estimate_cache_memory_mb(cls) -> float:
        return sum(m.p.nbytes + m.t.nbytes for m, _ in cls._cache.values()) / 2 ** 20

    @classmethod
    def save_cache(cls, filepath):
        """Persist the mesh cache (`mesh.py:385-396`).  Written as a NumPy archive of plain arrays (p and t per key plus the
        hit/miss counters), not a pickle of objects: the file can be read back by any version of this package."""
        arrays = {"keys": np.array(list(cls._cache.keys())), "counters": np.array([cls._cache_hits, cls._cache_misses])}
        for i, (mesh, _) in enumerate(cls._cache.values()):
            arrays[f"p{i}"], arrays[f"t{i}"] = mesh.p, mesh.t
        with open(filepath, "wb") as f:
            np.savez_compressed(f, **arrays)

    @classmethod
    def load_cache(cls, filepath):
        """Load a cache written by `save_cache` (`mesh.py:398-416`); a missing file leaves the cache untouched."""
        import os
        if not os.path.exists(filepath):
            return
        with np.load(filepath, allow_pickle=False) as z:
            cls._cache = OrderedDict((str(k), (MeshTri(z[f"p{i}"], z[f"t{i}"]), None)) for i, k in enumerate(z["keys"]))
            cls._cache_hits, cls._cache_misses = (int(v) for v in z["counters"])

    @classmethod
    def get_cache_stats(cls):
        tot = cls._cache_hits + cls._cache_misses
        return dict(size=len(cls._cache), hits=cls._cache_hits, misses=cls._cache_misses,
                    hit_rate=cls._cache_hits / tot if tot else 0.0)
