"""Parity of the CUDA path (through the C ABI and the drop-in class) against the oracle.

Tolerances (BASELINE.json north_star): identical CSR structure, assembled entries within 1e-12
relative, n_eff within 1e-8 relative.  "Relative" for an assembled entry is taken against the
largest magnitude in its row: entries that are pure round-off of cancelling element contributions
(e.g. the vertex / adjacent-edge mass entries, exactly 0 in exact arithmetic) have no meaningful
relative error of their own, and SciPy's own summation order for duplicates is unspecified.
"""
import os

import numpy as np
import pytest

from plfem_b200 import _cabi
from plfem_b200.solver_fem import TrueVectorialMaxwellSolver, ModeRecord
from oracle import fem_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rowwise_rel_err(M, R, row_scale=None):
    """max over entries of |M - R| / (scale of that row); structures must already match.  The scale is
    the largest |R| of the row, or `row_scale` when the matrix can have all-round-off rows (a Dxy row
    on an axis-aligned patch sums to ~1e-17 everywhere: measure it against the Laplacian row)."""
    d = np.abs(M.data - R.data)
    if row_scale is None:
        row_scale = np.asarray(abs(R).max(axis=1).todense()).ravel()
    scale = np.repeat(np.maximum(row_scale, 1e-300), np.diff(R.indptr))
    return float((d / scale).max())


def laplacian_row_scale(Dxx, Dyy):
    return np.asarray((abs(Dxx) + abs(Dyy)).max(axis=1).todense()).ravel()


def same_structure(M, R):
    return M.shape == R.shape and np.array_equal(M.indptr, R.indptr) and np.array_equal(M.indices, R.indices)


def check_A(A, rA):
    """A = sums of scalar matrices, where SciPy's sparse +/- drop results that are exactly 0.0.  Whether
    a cancelling sum of element contributions rounds to exactly 0.0 or to 1e-17 depends on the order in
    which SciPy happens to add duplicates (an unstable std::sort inside csr sum_duplicates), so the two
    structures may differ on such pure round-off entries and nowhere else."""
    assert A.shape == rA.shape
    D = (A - rA).tocsr()
    rowmax = np.maximum(abs(rA).max(axis=1).toarray().ravel(), 1e-300)
    D.eliminate_zeros()
    rel = np.abs(D.data) / np.repeat(rowmax, np.diff(D.indptr))
    assert rel.max() < 1e-12, f"A: row-relative deviation {rel.max():.2e}"
    only = ((A != 0).astype(np.int8) - (rA != 0).astype(np.int8)).tocsr()
    only.eliminate_zeros()
    assert only.nnz <= 2e-3 * rA.nnz                       # a few hundred of ~1e6 entries
    rows = np.repeat(np.arange(A.shape[0]), np.diff(only.indptr))
    vals = np.abs(np.asarray((A + rA)[rows, only.indices]).ravel())
    assert (vals <= 1e-13 * rowmax[rows]).all()            # and every one of them is round-off


@pytest.mark.parametrize("case", ["small_case", "cfg1"])
def test_assembled_system_matches_oracle(case, request):
    g, mesh = request.getfixturevalue(case)
    A, B, basis, Dxx, Dyy, Dxy, M_inv = TrueVectorialMaxwellSolver(g).assemble_hfield_system(mesh)
    rA, rB, rbasis, rDxx, rDyy, rDxy, rMinv = O.assemble_hfield_system(g, mesh)
    assert basis.N == rbasis.N and np.array_equal(basis.doflocs, rbasis.doflocs)
    assert np.array_equal(basis.get_dofs().all(), rbasis.boundary_dofs())
    lap = laplacian_row_scale(rDxx, rDyy)
    for name, M, R, sc in (("B", B, rB, None), ("Dxx", Dxx, rDxx, lap), ("Dyy", Dyy, rDyy, lap), ("Dxy", Dxy, rDxy, lap),
                           ("M_inv", M_inv, rMinv, None)):
        assert same_structure(M, R), f"{name}: CSR structure differs ({M.nnz} vs {R.nnz} nnz)"
        err = rowwise_rel_err(M, R, sc)
        assert err < 1e-12, f"{name}: row-relative deviation {err:.2e}"
        # entry-relative view of the same comparison (north_star: "entries within 1e-12 relative"): the entries that miss
        # 1e-12 of THEIR OWN magnitude must all be cancellation round-off — tiny against their row — and few
        ent = np.abs(M.data - R.data) / np.maximum(np.abs(R.data), 1e-300)
        miss = ent > 1e-12
        scale = np.repeat(np.maximum(sc if sc is not None else np.asarray(abs(R).max(axis=1).todense()).ravel(), 1e-300), np.diff(R.indptr))
        print(f"{case} {name}: {int((M.data == R.data).sum())}/{R.nnz} bit-equal, {int(miss.sum())} entries beyond 1e-12 entry-relative "
              f"(largest of them {np.abs(R.data[miss]).max() / scale[miss].max() if miss.any() else 0:.1e} of its row scale)")
        assert miss.mean() < 0.02, f"{name}: {miss.mean():.3%} of the entries beyond 1e-12 entry-relative"
        assert (np.abs(R.data[miss]) <= 1e-4 * scale[miss]).all(), f"{name}: a non-negligible entry misses 1e-12 entry-relative"
    check_A(A, rA)


def test_interior_matrices_and_all_scalar_blocks(small_case):
    g, mesh = small_case
    pb = _cabi.Problem(mesh)
    mat, keep = _cabi.material_struct(g)
    pb.assemble(mat)
    s = O.interior_system(g, mesh)
    M = pb.export_csr("B_int")
    assert same_structure(M, s["B_int"]) and rowwise_rel_err(M, s["B_int"]) < 1e-12
    check_A(pb.export_csr("A_int"), s["A_int"])
    basis, m = O.assemble_scalar_matrices(g, mesh)
    lap = laplacian_row_scale(m["dxx"], m["dyy"])
    for name, key in (("Kxx", "kxx"), ("Kyy", "kyy"), ("Kxy", "kxy"), ("Kyx", "kyx"), ("M", "mass")):
        M = pb.export_csr(name)
        assert same_structure(M, m[key]), name
        assert rowwise_rel_err(M, m[key], None if name == "M" else lap) < 1e-12, name


def test_custom_epsilon_callable_uses_host_samples(small_case):
    g, mesh = small_case

    class Graded:
        positions, core_radii, n_core, n_clad, k0 = g.positions, g.core_radii, g.n_core, g.n_clad, g.k0

        def epsilon(self, x, y):
            return (1.0 + 1.3 * np.exp(-(np.asarray(x) ** 2 + np.asarray(y) ** 2) / 30.0)).astype(complex)

    A, B, *_ = TrueVectorialMaxwellSolver(Graded()).assemble_hfield_system(mesh)
    rA, rB, *_ = O.assemble_hfield_system(Graded(), mesh)
    check_A(A, rA)
    assert same_structure(B, rB) and rowwise_rel_err(B, rB) < 1e-12


def test_degenerate_mesh_is_an_error(cfg1):
    from plfem_b200.mesh import MeshGenerator
    g, _ = cfg1
    raw = MeshGenerator._delaunay_mesh(g, 1.0)                 # still holds the 25 flat hull triangles
    with pytest.raises(_cabi.PlfemError) as e:
        TrueVectorialMaxwellSolver(g).assemble_hfield_system(raw)
    assert e.value.status == 3


def test_csr_spmv(small_case):
    g, mesh = small_case
    s = O.interior_system(g, mesh)
    x = np.random.default_rng(3).standard_normal(s["A_int"].shape[0])
    ctx = _cabi.Context.get(0)
    for M in (s["A_int"], s["B_int"]):
        y, ms = ctx.spmv_csr(M, x, repeat=3)
        ref = M @ x
        assert np.abs(y - ref).max() <= 1e-13 * np.abs(M).dot(np.abs(x)).max() and ms > 0


def _clusters(n_eff, rgap=1e-6):
    """Indices grouped into runs whose neighbouring n_eff differ by less than rgap (relative): inside such a
    run (a symmetry-degenerate pair, split only by the mesh) any rotation of the eigenvectors is as good as
    any other, ARPACK's own choice depends on its start vector."""
    out, cur = [], [0]
    for i in range(1, len(n_eff)):
        if abs(n_eff[i] - n_eff[i - 1]) <= rgap * abs(n_eff[i]):
            cur.append(i)
        else:
            out.append(cur); cur = [i]
    out.append(cur)
    return out


def _compare_modes(g, mesh, n_modes, rtol_neff=1e-8, **opts):
    solver = TrueVectorialMaxwellSolver(g)
    modes, raw = solver.solve_vectorial_modes(mesh, n_modes, return_raw=True, **opts)
    rmodes, rraw = O.solve_vectorial_modes(g, mesh, n_modes, return_raw=True)
    st = raw["stats"]
    assert st["nconv"] >= len(raw["beta_sq"]) and st["max_residual"] < 1e-9
    # eigenvalues: same k pairs nearest sigma, beta^2 to 2e-8 relative <=> n_eff to 1e-8 relative
    assert len(raw["beta_sq"]) == len(rraw["beta_sq"])
    assert np.abs(raw["beta_sq"] / rraw["beta_sq"] - 1).max() < 2 * rtol_neff
    # compare the unfiltered records (the filters are thresholds on the same numbers)
    mr, rr = raw["modes_raw"], rraw["modes_raw"]
    assert len(mr) == len(rr)
    for m, r in zip(mr, rr):
        assert isinstance(m, ModeRecord) and set(m) == set(r)
        assert abs(m["n_eff"] / r["n_eff"] - 1) < rtol_neff and abs(m["beta"] / r["beta"] - 1) < rtol_neff
        assert m.n_eff == m["n_eff"] and m.polarization_state == m["polarization"]
        assert m["is_vectorial"] is True and m["method"] == "H-field_V18.10"
        assert abs(np.sum(m["Ex_dofs"] ** 2) + np.sum(m["Ey_dofs"] ** 2) - 1) < 1e-12
    vec = lambda m: np.concatenate([m["Ex_dofs"], m["Ey_dofs"]])
    for cl in _clusters([r["n_eff"] for r in rr]):
        V, R = np.array([vec(mr[i]) for i in cl]), np.array([vec(rr[i]) for i in cl])
        # same invariant subspace: all principal angles vanish.  The vectors of a cluster are B-orthogonal, not
        # l2-orthogonal, so both sets are orthonormalised first (cosines = singular values of Qv^T Qr).
        Qv, Qr = np.linalg.qr(V.T)[0], np.linalg.qr(R.T)[0]
        assert np.abs(np.linalg.svd(Qv.T @ Qr, compute_uv=False) - 1).max() < 1e-6
        for key in ("confinement", "P_x", "P_y", "div_ratio"):          # traces over the cluster are rotation-invariant
            a, b = sum(mr[i][key] * (mr[i]["beta"] ** 2 if key == "div_ratio" else 1) for i in cl), \
                sum(rr[i][key] * (rr[i]["beta"] ** 2 if key == "div_ratio" else 1) for i in cl)
            # both sides are Ritz vectors converged to tol = 1e-7 only: small projections carry that error
            assert abs(a - b) <= 5e-6 * abs(b) + (1e-8 if key == "div_ratio" else 1e-9), key
        if len(cl) == 1:
            m, r = mr[cl[0]], rr[cl[0]]
            assert abs(m["PDL_dB"] - r["PDL_dB"]) < 1e-4 and m["polarization"] == r["polarization"]
            assert m["core_overlap"] == m["confinement"]
    if all(len(c) == 1 for c in _clusters([r["n_eff"] for r in rr])):
        assert [m["n_eff"] for m in modes] == pytest.approx([m["n_eff"] for m in rmodes], rel=rtol_neff)
    return st


def test_modes_small_case(small_case):
    g, mesh = small_case
    _compare_modes(g, mesh, 4)


def test_modes_config1(cfg1):
    g, mesh = cfg1
    st = _compare_modes(g, mesh, 10)
    assert st["kernel_launches"] > 0


def test_modes_config2(cfg2):
    """19-core cross-section, n_modes = 40 -> k = 52 eigenpairs on a 93k-unknown system."""
    g, mesh = cfg2
    st = _compare_modes(g, mesh, 40)
    assert st["n_levels"] > 10 and st["nconv"] >= 52


def test_single_vector_and_block_lanczos_agree(small_case):
    g, mesh = small_case
    s = TrueVectorialMaxwellSolver(g)
    a, ra = s.solve_vectorial_modes(mesh, 4, return_raw=True, block=1)      # ARPACK-like single-vector recurrence
    b, rb = s.solve_vectorial_modes(mesh, 4, return_raw=True, block=4)      # default: 4 vectors per operator application
    assert ra["stats"]["n_block_op"] == 0 and rb["stats"]["n_block_op"] > 0
    assert rb["stats"]["n_block_op"] < ra["stats"]["n_op"]                  # fewer sequential operator applications
    assert np.abs(ra["beta_sq"] / rb["beta_sq"] - 1).max() < 1e-12
    assert len(a) == len(b)


def test_block_ldlt_solve_matches_superlu(small_case):
    """`plfem_debug_solve`: the sweeps (TMA-streamed bottom subtrees + per-level kernels) on the block-LDL^T factors agree
    with SuperLU: raw to 1e-7 (symmetrised pivot-block inverses), to 1e-11 after one refinement step."""
    from scipy.sparse.linalg import splu
    from plfem_b200.solver_fem import sigma_estimate
    g, mesh = small_case
    s = O.interior_system(g, mesh)
    sigma = sigma_estimate(g)
    K = (s["A_int"] - sigma * s["B_int"]).tocsc()
    pb = _cabi.Problem(mesh)
    mat, keep = _cabi.material_struct(g)
    pb.solve_modes(mat, sigma, 16)
    b = s["B_int"] @ np.random.default_rng(2).standard_normal(K.shape[0])
    xr = splu(K).solve(b)
    x_raw, x_ref = pb.debug_solve(sigma, b, 0), pb.debug_solve(sigma, b, 1)
    assert np.linalg.norm(x_raw - xr) / np.linalg.norm(xr) < 1e-7
    assert np.linalg.norm(x_ref - xr) / np.linalg.norm(xr) < 1e-11


def test_readme_surface(cfg1):
    g, mesh = cfg1
    modes = TrueVectorialMaxwellSolver(g, n_modes=10).solve()
    assert len(modes) > 0 and all(1.0 < m.n_eff < 1.01 * g.n_core for m in modes)
    assert [m.n_eff for m in modes] == sorted((m.n_eff for m in modes), reverse=True)


def test_start_vector_and_determinism(small_case):
    g, mesh = small_case
    s = TrueVectorialMaxwellSolver(g)
    a = s.solve_vectorial_modes(mesh, 4)
    b = s.solve_vectorial_modes(mesh, 4)
    assert all(np.array_equal(x["Ex_dofs"], y["Ex_dofs"]) and x["n_eff"] == y["n_eff"] for x, y in zip(a, b))
    n2 = 2 * (len(a[0]["Ex_dofs"]))
    c = s.solve_vectorial_modes(mesh, 4, v0=np.random.default_rng(5).uniform(-1, 1, n2))
    assert np.allclose([m["n_eff"] for m in a], [m["n_eff"] for m in c], rtol=1e-9)


# ---- forest of designs: plfem_solve_modes_batch ----------------------------------------------------------------
def _forest_jobs(small_case):
    import plfem_b200 as P
    g, mesh = small_case
    g2 = P.MCFGeometry(3, 6.0, 1.2, P.IPDipCauchy.n(1490), 1.0, 1.49)          # same mesh recipe, another band
    g3 = P.MCFGeometry(4, 5.0, 1.0, 1.53, 1.0, 1.55)                           # another layout, another mesh
    mesh3, _ = P.MeshGenerator.generate(g3, refinement=0.4)
    return [(g, mesh, 4), (g2, mesh, 4), (g3, mesh3, 6)]


def test_forest_matches_single_solves(small_case):
    """Designs solved together as one block-diagonal problem give what each gives alone."""
    from plfem_b200.batch import ForestPool
    jobs = _forest_jobs(small_case)
    with ForestPool(batch=len(jobs), workers=1) as pool:
        forest = pool.solve_forest(jobs, return_raw=True)
        stats = pool.last_stats
    assert all(s["batch_size"] == len(jobs) for s in stats)
    assert len({s["kernel_launches"] for s in stats}) == 1              # the designs share every launch
    for (g, mesh, n), (modes, raw) in zip(jobs, forest):
        alone, araw = TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, n, return_raw=True)
        assert raw["stats"]["nconv"] >= len(raw["beta_sq"]) and raw["stats"]["max_residual"] < 1e-9
        assert np.abs(raw["beta_sq"] / araw["beta_sq"] - 1).max() < 1e-9   # lockstep may stop a step later: same pairs
        assert len(modes) == len(alone)
        ne = np.array([a["n_eff"] for a in alone])
        for i, (m, a) in enumerate(zip(modes, alone)):
            assert abs(m["n_eff"] / a["n_eff"] - 1) < 1e-9
            # polarization label and confinement are properties of the VECTOR, and inside a (near-)degenerate pair the vector is
            # any rotation of the eigenspace: a forest may stop a block step later than the single solve and return another one
            if len(ne) == 1 or np.min(np.abs(np.delete(ne, i) - ne[i])) > 1e-6:
                assert m["polarization"] == a["polarization"]
                assert abs(m["confinement"] - a["confinement"]) < 1e-6


def test_forest_of_identical_designs_is_bit_identical_to_single(small_case):
    from plfem_b200.batch import ForestPool
    g, mesh = small_case
    alone, araw = TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, 4, return_raw=True)
    with ForestPool(batch=3, workers=1) as pool:
        forest = pool.solve_forest([(g, mesh, 4)] * 3, return_raw=True)
    for modes, raw in forest:
        assert np.array_equal(raw["beta_sq"], araw["beta_sq"])
        assert np.array_equal(raw["evecs"], araw["evecs"]) and np.array_equal(raw["metrics"], araw["metrics"])


def test_forest_matches_oracle_config1(cfg1):
    """Two copies of config 1 at different bands in one forest, each against the oracle's eigsh."""
    import plfem_b200 as P
    from plfem_b200.batch import ForestPool
    g, mesh = cfg1
    g2 = P.MCFGeometry(7, 8.0, 1.5, P.IPDipCauchy.n(1600), 1.0, 1.60)
    jobs = [(g, mesh, 10), (g2, mesh, 10)]
    with ForestPool(batch=2, workers=1) as pool:
        forest = pool.solve_forest(jobs, return_raw=True)
    for (gg, mm, n), (modes, raw) in zip(jobs, forest):
        rmodes, rraw = O.solve_vectorial_modes(gg, mm, n, return_raw=True)
        assert np.abs(raw["beta_sq"] / rraw["beta_sq"] - 1).max() < 2e-8      # n_eff to 1e-8 relative
        assert len(modes) == len(rmodes)


def test_pool_solve_many_keeps_job_order(small_case):
    from plfem_b200.batch import ForestPool
    jobs = _forest_jobs(small_case) * 2
    with ForestPool(batch=2, workers=2) as pool:
        out = pool.solve_many(jobs)
    assert len(out) == len(jobs)
    for (g, mesh, n), modes in zip(jobs, out):
        alone = TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, n)
        assert [m["n_eff"] for m in modes] == pytest.approx([m["n_eff"] for m in alone], rel=1e-9)


def _check_record_against_oracle(rec, d, mesh, f):
    """A sweep record against the oracle's mode list of the same design.  The divergence filter of `solver_fem.py:228-231`
    thresholds a per-vector quantity (v^T D v / beta^2) that is NOT defined for the members of a degenerate pair (any rotation
    of the pair is an eigenbasis; ARPACK's own answer changes with its random start vector), so modes that sit within 30 % of
    the threshold, or belong to a degenerate pair straddling it, may legitimately fall on either side: the record must agree
    with the oracle on every mode that is not ambiguous in that sense, and exactly when no mode is."""
    from plfem_b200 import sweep
    modes, raw = O.solve_vectorial_modes(sweep.design_geometry(d), mesh, d["n_modes"], return_raw=True)
    mr = raw["modes_raw"]
    dr = np.array([m["div_ratio"] for m in mr])
    ne_raw = np.array([m["n_eff"] for m in mr])
    thr = max(10 * np.median(dr), 50 * dr.min(), 1e-6)
    amb = np.abs(dr / thr - 1) < 0.3
    for i in range(len(mr)):
        twins = np.abs(ne_raw / ne_raw[i] - 1) < 1e-9
        if twins.sum() > 1 and dr[twins].min() < 1.3 * thr and dr[twins].max() > thr / 1.3:
            amb[i] = True
    guided_ids = {id(m) for m in modes}
    sure = [m for m, a in zip(mr, amb) if id(m) in guided_ids and not a]
    n_amb = int(amb.sum())
    assert len(sure) <= rec[f["n_modes_found"]] <= len(sure) + n_amb, (d["n_cores"], d["wavelength_nm"], len(modes), n_amb)
    if n_amb == 0:
        assert len(modes) == rec[f["n_modes_found"]]
        ne = np.array([m["n_eff"] for m in modes])
        assert abs(ne.max() / rec[f["n_eff_max"]] - 1) < 1e-8 and abs(ne.min() / rec[f["n_eff_min"]] - 1) < 1e-8
        assert abs(ne.mean() / rec[f["n_eff_mean"]] - 1) < 1e-8
        assert abs(np.mean([m["confinement"] for m in modes]) - rec[f["confinement_mean"]]) < 5e-6
    # the leading modes of the record, as long as the oracle's list is unambiguous there
    amb_ne = ne_raw[amb]
    for k, m in enumerate(modes[:sweep.N_PER_MODE]):
        if len(amb_ne) and m["n_eff"] <= amb_ne.max() * (1 + 1e-9):
            break
        assert abs(m["n_eff"] / rec[f[f"n_eff_mode_{k}"]] - 1) < 1e-8


def test_sweep_forest_mode_matches_design_by_design():
    """Config 3 (S/C/L/U bands on the 7-core PL): the sweep driver in forest mode and design by design, and every band
    against the oracle."""
    from plfem_b200 import sweep
    from plfem_b200.mesh import MeshGenerator
    designs = sweep.band_sweep_designs()
    a = sweep.run_sweep(designs, forest=4)
    b = sweep.run_sweep(designs)
    f = {k: i for i, k in enumerate(sweep.RECORD_FIELDS)}
    assert list(a[:, f["success"]]) == [1, 1, 1, 1] and list(b[:, f["success"]]) == [1, 1, 1, 1]
    for key in ("n_modes_found", "n_eff_max", "n_eff_min", "n_eff_mean", "confinement_mean", "n_dofs", "sigma_shift", "wavelength_nm"):
        assert np.allclose(a[:, f[key]], b[:, f[key]], rtol=1e-8, atol=0), key
    assert np.allclose(a[:, f["PDL_mean_dB"]], b[:, f["PDL_mean_dB"]], atol=1e-4)
    mesh, _ = MeshGenerator.generate(sweep.design_geometry(designs[0]), 1.0)
    for i, d in enumerate(designs):                            # ALL four bands against the oracle
        _check_record_against_oracle(a[i], d, mesh, f)


def test_lhs_sample_of_all_layouts_in_forests():
    """Config 4 (sample): LHS designs over the layouts and bands, different meshes and different k in the same forest."""
    from plfem_b200 import sweep
    from plfem_b200.mesh import MeshGenerator
    designs = sweep.lhs_designs(36)[::3]                       # 12 designs: 2 ... 19 cores, all four bands
    rec = sweep.run_sweep(designs, forest=6)
    f = {k: i for i, k in enumerate(sweep.RECORD_FIELDS)}
    assert rec[:, f["success"]].tolist() == [1.0] * len(designs)
    assert (rec[:, f["n_modes_found"]] > 0).all() and (rec[:, f["n_eff_max"]] > 1.0).all()
    for i, d in enumerate(designs):                            # ALL twelve designs against the oracle (2 ... 19 cores)
        mesh, _ = MeshGenerator.generate(sweep.design_geometry(d), 1.0)
        _check_record_against_oracle(rec[i], d, mesh, f)


def test_coarse_structured_mesh_reaches_the_oracle():
    """A coarse structured mesh at this shift — the case that diverged before the pivot-block inverses were symmetrised
    (valid inputs reported SINGULAR where eigsh + SuperLU solve them).  It must now reach the oracle's eigenvalues."""
    import plfem_b200 as P
    from plfem_b200.solver_fem import sigma_estimate
    from scipy.sparse.linalg import eigsh
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
    nx = 60
    xs = np.linspace(-32.0, 32.0, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    idx = np.arange((nx + 1) ** 2).reshape(nx + 1, nx + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    mesh = P.MeshTri(np.vstack([X.ravel(), Y.ravel()]), np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])]))
    pb = _cabi.Problem(mesh)
    mat, keep = _cabi.material_struct(g)
    sigma = sigma_estimate(g)
    vals, _, _, _, st = pb.solve_modes(mat, sigma, 22, want_vectors=False)
    assert st.n_block_op < 200 and st.max_residual < 1e-9
    s = O.interior_system(g, mesh)
    ref = np.sort(eigsh(s["A_int"], k=22, M=s["B_int"], sigma=sigma, which="LM", tol=1e-9)[0])
    assert np.abs(vals / ref - 1).max() < 2e-8


def test_failed_design_does_not_poison_its_forest(small_case):
    """A design whose factorisation is useless (a non-finite core index: every pivot block of its tree is NaN) is reported
    alone; the other designs of the same forest give exactly their single-solve results."""
    import copy
    from plfem_b200.batch import ForestPool
    g, mesh = small_case
    g_bad = copy.copy(g)
    g_bad.n_core = float("nan")
    alone, araw = TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, 4, return_raw=True)
    with ForestPool(batch=3, workers=1) as pool:
        out = pool.solve_forest([(g, mesh, 4), (g_bad, mesh, 4), (g, mesh, 4)], return_raw=True)
    assert isinstance(out[1], _cabi.PlfemError) and out[1].status in (5, 6)
    for modes, raw in (out[0], out[2]):
        assert np.abs(raw["beta_sq"] / araw["beta_sq"] - 1).max() < 1e-9 and len(modes) == len(alone)
        assert raw["stats"]["max_residual"] < 1e-9


@pytest.mark.gpu
def test_structured_mesh_is_solved_without_refinement():
    """120 x 120 structured cells: the mesh family on which the factorisation diverged before the pivot-block inverses were
    symmetrised.  Eigenvalues against the oracle; the probe must find the raw solve accurate."""
    import plfem_b200 as P
    from plfem_b200.solver_fem import sigma_estimate
    from scipy.sparse.linalg import eigsh
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
    nx = 120
    xs = np.linspace(-32.0, 32.0, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    idx = np.arange((nx + 1) ** 2).reshape(nx + 1, nx + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    mesh = P.MeshTri(np.vstack([X.ravel(), Y.ravel()]), np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])]))
    pb = _cabi.Problem(mesh)
    mat, keep = _cabi.material_struct(g)
    sigma = sigma_estimate(g)
    vals, _, _, _, st = pb.solve_modes(mat, sigma, 22, want_vectors=False)
    assert st.refine_steps <= 1 and st.max_residual < 1e-8
    s = O.interior_system(g, mesh)
    ref = np.sort(eigsh(s["A_int"], k=22, M=s["B_int"], sigma=sigma, which="LM", tol=1e-10)[0])
    assert np.abs(np.sort(vals) / ref - 1).max() < 1e-8


@pytest.mark.gpu
def test_scalar_helmholtz_solver_matches_oracle(small_case):
    """`ScalarHelmholtzSolver.solve` (`solver_fem.py:245-276`) on the CUDA kernels (scalar pencil in the Hx block, natural
    boundary condition) against the oracle's restatement with the real eigsh: same modes, n_eff to 1e-8."""
    from plfem_b200.solver_fem import ScalarHelmholtzSolver
    g, mesh = small_case
    ref = O.solve_scalar_modes(g, mesh, 6)
    s = ScalarHelmholtzSolver(g)
    got = s.solve(mesh, 6)
    assert len(got) == len(ref) > 0 and s.last_stats["max_residual"] < 1e-9
    for a, b in zip(got, ref):
        assert abs(a["n_eff"] / b["n_eff"] - 1) < 1e-8
        assert a["polarization"] == "scalar" and a["is_vectorial"] is False and a["PDL_dB"] == 0.0
        assert a["field_vector"].shape == b["field_vector"].shape
    ne = np.array([m["n_eff"] for m in ref])
    lonely = [i for i in range(len(ne)) if np.min(np.abs(np.delete(ne, i) - ne[i])) > 1e-4]     # non-degenerate modes
    assert lonely
    for i in lonely:
        assert abs(got[i]["confinement"] - ref[i]["confinement"]) < 1e-6
        c = abs(float(got[i]["field_vector"] @ ref[i]["field_vector"])) / (np.linalg.norm(got[i]["field_vector"]) * np.linalg.norm(ref[i]["field_vector"]))
        assert c > 1 - 1e-8


@pytest.mark.gpu
def test_config5_two_million_unknowns():
    """Config 5 of BASELINE.json: 500 x 500 structured cells, 1,996,002 unknowns.  The oracle cannot run at this size; what is
    checked is size-independent: every wanted pair converged, the TRUE backward error of every returned eigenpair
    ||A x - lambda B x|| / ((||A|| + |lambda| ||B||) ||x||) is below 1e-9, and the eigenvalues sit inside the guidance window
    next to those of the 120 x 120 mesh of the same cross-section (checked against the oracle above)."""
    import plfem_b200 as P
    from plfem_b200.solver_fem import sigma_estimate
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
    mesh = P.MeshTri.init_structured(500, 500, 32.0)
    pb = _cabi.Problem(mesh)
    assert 2 * pb.n_interior == 1_996_002
    mat, keep = _cabi.material_struct(g)
    sigma = sigma_estimate(g)
    vals, _, _, _, st = pb.solve_modes(mat, sigma, 22, want_vectors=False)
    assert st.nconv == 22 and st.max_residual < 1e-9 and st.refine_steps <= 1
    ne = np.sqrt(vals[vals > 0]) / g.k0
    assert ((ne > g.n_clad) & (ne < 1.01 * g.n_core)).sum() >= 10
    coarse = _cabi.Problem(P.MeshTri.init_structured(120, 120, 32.0))
    cvals, *_ = coarse.solve_modes(mat, sigma, 22, want_vectors=False)
    assert np.abs(np.sort(vals) / np.sort(cvals) - 1).max() < 2e-2          # h-convergence: the same cluster of modes
    pb.close(); coarse.close()


@pytest.mark.gpu
def test_bands_on_one_mesh_share_one_analysis(cfg1):
    """The four bands of config 3 sit on ONE mesh (the recipe uses the geometry only, `mesh.py:232-289`): inside a forest they
    share the ordering and the front plan of the first — same results as analysing each on its own, a fraction of the host time."""
    import plfem_b200 as P
    from plfem_b200.batch import ForestPool
    g0, mesh = cfg1
    jobs = [(P.MCFGeometry(7, 8.0, 1.5, P.IPDipCauchy.n(w), 1.0, w / 1000.0), mesh, 10) for w in (1490, 1550, 1600, 1650)]
    with ForestPool(batch=4, workers=1, share_analysis=True) as pool:
        shared = pool.solve_forest(jobs, return_raw=True)
        st_shared = [r[1]["stats"] for r in shared]
    with ForestPool(batch=4, workers=1, share_analysis=False) as pool:
        own = pool.solve_forest(jobs, return_raw=True)
        st_own = [r[1]["stats"] for r in own]
    for (ma, ra), (mb, rb) in zip(shared, own):
        assert np.array_equal(ra["beta_sq"], rb["beta_sq"]) and len(ma) == len(mb)
    assert all(s["ms_symbolic"] < 0.5 * o["ms_symbolic"] for s, o in zip(st_shared[1:], st_own[1:]))
    assert st_shared[0]["n_fronts"] == st_shared[3]["n_fronts"] == st_own[3]["n_fronts"]


@pytest.mark.gpu
def test_refinement_is_per_design_inside_a_forest(small_case, cfg1):
    """A forest mixing a design whose raw block-LDL^T solve is accurate (the recipe mesh of config 1: probe rho ~ 6e-11, no
    refinement alone) with one that needs a refinement step (a coarse structured mesh: rho > 1e-9): the forest's refinement solve
    must skip the first design's fronts, rows and correction — and both must still give their single-solve eigenvalues and pass
    the backward-error bar."""
    import plfem_b200 as P
    from plfem_b200.batch import ForestPool
    g, mesh = cfg1
    coarse = P.MeshTri.init_structured(60, 60, 32.0)
    jobs = [(g, mesh, 10), (g, coarse, 10), (g, mesh, 10)]
    alone = []
    for gg, mm, n in jobs[:2]:
        modes, raw = TrueVectorialMaxwellSolver(gg).solve_vectorial_modes(mm, n, return_raw=True)
        alone.append(raw)
    assert alone[0]["stats"]["refine_steps"] == 0 and alone[1]["stats"]["refine_steps"] >= 1, \
        (alone[0]["stats"]["probe_rho"], alone[1]["stats"]["probe_rho"])
    with ForestPool(batch=3, workers=1) as pool:
        out = pool.solve_forest(jobs, return_raw=True)
    for (modes, raw), ref in zip(out, (alone[0], alone[1], alone[0])):
        assert raw["stats"]["max_residual"] < 1e-9
        assert np.abs(raw["beta_sq"] / ref["beta_sq"] - 1).max() < 1e-9
    assert out[0][1]["stats"]["refine_steps"] == alone[1]["stats"]["refine_steps"]      # the forest reports the maximum
    assert np.array_equal(out[0][1]["beta_sq"], out[2][1]["beta_sq"])                     # identical designs, identical results


@pytest.mark.gpu
def test_dataflow_and_per_level_sweeps_are_bit_identical(tmp_path):
    """`PLFEM_SWEEP=levels` launches the same sweep tasks level by level instead of as one dataflow launch per direction: the
    eigenvalues must agree bit for bit (the variable is read once per process, hence two subprocesses)."""
    import subprocess
    import sys as _sys
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import plfem_b200 as P; "
            "from plfem_b200.solver_fem import TrueVectorialMaxwellSolver as S; "
            "g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55); mesh, _ = P.MeshGenerator.generate(g, refinement=0.4); "
            "m, raw = S(g).solve_vectorial_modes(mesh, 4, return_raw=True); np.save(sys.argv[1], raw['beta_sq'])" % ROOT)
    outs = []
    for mode in ("", "levels"):
        f = str(tmp_path / f"beta_{mode or 'dataflow'}.npy")
        env = dict(os.environ, PLFEM_SWEEP=mode)
        subprocess.run([_sys.executable, "-c", code, f], env=env, check=True, timeout=300)
        outs.append(np.load(f))
    assert np.array_equal(outs[0], outs[1])


@pytest.mark.gpu
def test_local_first_orthogonalisation_matches_two_full_passes(tmp_path):
    """The block Lanczos step projects first against the last two blocks only (the kept Ritz vectors right after a restart) and
    then against the whole basis; `PLFEM_CGS=full` runs two full passes.  Same number of block steps, eigenvalues equal to
    1e-11 relative, backward errors under the bar in both (the variable is read once per process: two subprocesses)."""
    import json
    import subprocess
    import sys as _sys
    code = ("import sys, json, numpy as np; sys.path.insert(0, %r); import plfem_b200 as P; "
            "from plfem_b200.solver_fem import TrueVectorialMaxwellSolver as S; "
            "g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55); mesh, _ = P.MeshGenerator.generate(g, refinement=0.7); "
            "m, raw = S(g).solve_vectorial_modes(mesh, 10, return_raw=True); np.save(sys.argv[1], raw['beta_sq']); "
            "print(json.dumps({k: raw['stats'][k] for k in ('n_block_op', 'n_restart', 'max_residual')}))" % ROOT)
    outs, stats = [], []
    for mode in ("", "full"):
        f = str(tmp_path / f"beta_cgs_{mode or 'local'}.npy")
        env = dict(os.environ, PLFEM_CGS=mode)
        r = subprocess.run([_sys.executable, "-c", code, f], env=env, check=True, timeout=300, capture_output=True, text=True)
        stats.append(json.loads(r.stdout.strip().splitlines()[-1]))
        outs.append(np.load(f))
    assert stats[0]["n_restart"] >= 1, stats                      # the case exercises the step right after a thick restart
    assert stats[0]["n_block_op"] == stats[1]["n_block_op"], stats
    assert max(s["max_residual"] for s in stats) < 1e-9, stats
    assert np.abs(outs[0] / outs[1] - 1).max() < 1e-11
