"""The oracle against closed-form facts and its own frozen outputs (it has no reference golden
vectors to be pinned to: parity unpinned, see oracle/fem_oracle.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

import plfem_b200 as P
from plfem_b200.mesh import MeshTri, signed_double_area
from oracle import fem_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class UniformGeometry:
    positions = np.zeros((0, 2)); core_radii = np.zeros(0); n_core = 1.5; n_clad = 1.0; k0 = 4.0

    def epsilon(self, x, y):
        return np.full_like(np.asarray(x, dtype=float), 2.0, dtype=complex)


def test_quadrature_and_shape_functions():
    assert abs(O.QUAD_W.sum() - 0.5) < 1e-14
    phi, dx, dy = O.reference_tables()
    assert np.allclose(phi.sum(axis=0), 1.0, atol=1e-15)            # partition of unity
    assert np.allclose(dx.sum(axis=0), 0.0, atol=1e-14) and np.allclose(dy.sum(axis=0), 0.0, atol=1e-14)
    # nodal property at the six reference nodes
    for i in range(6):
        v, _ = O.lbasis(O.REF_DOFLOCS, i)
        assert np.allclose(v, np.eye(6)[i], atol=1e-15)
    # degree-4 exactness: int x^a y^b over the reference triangle = a! b! / (a+b+2)!
    from math import factorial as f
    for a in range(5):
        for b in range(5 - a):
            q = (O.QUAD_W * O.QUAD_X[0] ** a * O.QUAD_X[1] ** b).sum()
            assert abs(q - f(a) * f(b) / f(a + b + 2)) < 1e-14


def test_single_triangle_exact_matrices():
    mesh = MeshTri(np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]), np.array([[0], [1], [2]]))
    basis, m = O.assemble_scalar_matrices(UniformGeometry(), mesh)
    # facets are numbered lexicographically: (0,1) (0,2) (1,2), so local edge (1,2) is global DOF 5
    ed = basis.element_dofs.ravel()
    assert basis.N == 6 and np.array_equal(ed, [0, 1, 2, 3, 5, 4])
    loc = lambda name: m[name].toarray()[np.ix_(ed, ed)]          # back to local numbering
    M = loc("mass")
    exact = np.array([[6, -1, -1, 0, -4, 0], [-1, 6, -1, 0, 0, -4], [-1, -1, 6, -4, 0, 0],
                      [0, 0, -4, 32, 16, 16], [-4, 0, 0, 16, 32, 16], [0, -4, 0, 16, 16, 32]]) / 360.0
    assert np.allclose(M, exact, atol=1e-16)
    assert np.allclose(loc("minv"), exact / 2.0, atol=1e-16)
    K = loc("dxx") + loc("dyy")                          # P2 Laplacian of the unit right triangle
    exactK = np.array([[6, 1, 1, -4, 0, -4], [1, 3, 0, -4, 0, 0], [1, 0, 3, 0, 0, -4],
                       [-4, -4, 0, 16, -8, 0], [0, 0, 0, -8, 16, -8], [-4, 0, -4, 0, -8, 16]]) / 6.0
    assert np.allclose(m["dxx"].toarray(), m["dxx"].toarray().T, atol=1e-15)
    assert np.allclose(K @ np.ones(6), 0, atol=1e-14)
    assert np.allclose(K, exactK, atol=1e-14)
    assert np.allclose(m["kxx"].toarray(), m["dyy"].toarray() / 2.0, atol=1e-15)


def test_global_identities(small_case):
    g, mesh = small_case
    basis, m = O.assemble_scalar_matrices(g, mesh)
    area = 0.5 * np.abs(signed_double_area(mesh.p, mesh.t)).sum()
    one = np.ones(basis.N)
    assert abs(one @ (m["mass"] @ one) - area) < 1e-10 * area
    for k in ("kxx", "kyy", "kxy", "kyx", "dxx", "dyy", "dxy"):
        assert np.abs(m[k] @ one).max() < 1e-9 * np.abs(m[k]).max()
    assert abs(m["kyx"] - m["kxy"].T).max() < 1e-12 * abs(m["kxy"]).max()
    A, B, basis2, *_ = O.assemble_hfield_system(g, mesh)
    assert abs(A - A.T).max() < 1e-12 * abs(A).max() and abs(B - B.T).max() < 1e-14 * abs(B).max()
    assert A.shape == (2 * basis.N, 2 * basis.N) and A.has_sorted_indices
    # a quadratic field is reproduced by its nodal values: x^2 integrates exactly through M
    x = basis.doflocs[0]
    exact = None
    q = (one @ (m["mass"] @ (x * x)))
    xq = basis.x[0]
    assert abs(q - (xq ** 2 * basis.dx).sum()) < 1e-9 * abs(q)
    # faithful_cost evaluates epsilon 180 times instead of once: same matrices
    _, mf = O.assemble_scalar_matrices(g, mesh, faithful_cost=True)
    assert (mf["minv"] != m["minv"]).nnz == 0 and (mf["kxy"] != m["kxy"]).nnz == 0


def test_boundary_dofs_are_the_hull(small_case):
    g, mesh = small_case
    basis = O.P2Basis(mesh)
    b = basis.boundary_dofs()
    r = np.hypot(*basis.doflocs[:, b])
    assert r.min() > 0.9 * g.domain_radius
    assert len(np.setdiff1d(np.arange(basis.N), b)) == basis.N - len(b)


@pytest.mark.parametrize("name", ["small3"])
def test_oracle_regression(name, small_case):
    gold = json.load(open(os.path.join(G, "oracle_cfg.json")))[name]
    g, mesh = small_case
    modes, raw = O.solve_vectorial_modes(g, mesh, 4, return_raw=True)
    s = raw["system"]
    assert (mesh.p.shape[1], mesh.t.shape[1], s["basis"].N, len(s["interior"])) == (gold["V"], gold["T"], gold["N"], gold["N_solve"])
    assert raw["sigma"] == gold["sigma"]
    assert digest(s["A_int"].indptr.astype(np.int64)) == gold["A_int_indptr"]
    assert digest(s["A_int"].indices.astype(np.int64)) == gold["A_int_indices"]
    assert digest(s["B_int"].indices.astype(np.int64)) == gold["B_int_indices"]
    assert np.allclose(raw["beta_sq"], gold["beta_sq"], rtol=1e-9, atol=0)
    assert np.allclose([m["n_eff"] for m in modes], gold["n_eff"], rtol=1e-9, atol=0)
    for m in modes:
        assert set(m) >= {"n_eff", "beta", "Ex_dofs", "Ey_dofs", "P_x", "P_y", "PDL_dB", "polarization",
                          "confinement", "core_overlap", "div_ratio", "is_vectorial", "method"}
        assert abs(np.sum(m["Ex_dofs"] ** 2) + np.sum(m["Ey_dofs"] ** 2) - 1.0) < 1e-12
    assert [m["n_eff"] for m in modes] == sorted((m["n_eff"] for m in modes), reverse=True)


def test_polarization_labels():
    class Gm:
        positions = np.array([[0.0, 0.0]]); core_radii = np.array([1.0])
    x = np.array([0.0, 0.5, 3.0]); y = np.zeros(3)
    for ratio, label in ((100.0, "TE-like"), (5.0, "HE-like"), (1.0, "Hybrid"), (0.2, "EH-like"), (0.01, "TM-like")):
        vx = np.array([np.sqrt(ratio), 0.0, 7.0]); vy = np.array([1.0, 0.0, 9.0])
        pol, pdl, px, py = O.polarization_from_interp(vx, vy, x, y, Gm)
        assert pol == label and abs(px / py - ratio) < 1e-12 * ratio
        assert abs(pdl - min(abs(10 * np.log10(ratio)), 50.0)) < 1e-12
