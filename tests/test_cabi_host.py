"""C-ABI library without a GPU: it loads, exports every declared symbol, and its host logic
(DOF tables, quadrature points, front plan, restart eigensolver) is right.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from plfem_b200 import _cabi
from oracle import fem_oracle as O
import frontal_reference as FR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "plfem.h")).read()
    declared = sorted(set(re.findall(r"\b(plfem_[a-z_0-9]+)\s*\(", header)))
    assert declared == sorted(_cabi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.plfem_version()


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors have the sizes gcc gives the header's structs (the header must also compile as plain C)."""
    import shutil
    import subprocess
    assert ctypes.sizeof(_cabi.MeshInfo) == 64
    assert ctypes.sizeof(_cabi.Material) == 80
    assert ctypes.sizeof(_cabi.SolveOpts) == 64
    assert ctypes.sizeof(_cabi.SolveStats) == 112
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "plfem.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(plfem_mesh_info), '
                   'sizeof(plfem_material), sizeof(plfem_solve_opts), sizeof(plfem_solve_stats)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(c) for c in (_cabi.MeshInfo, _cabi.Material, _cabi.SolveOpts, _cabi.SolveStats)]


def test_product_has_no_cpu_path(small_case):
    g, mesh = small_case
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_cabi.PlfemError):
        _cabi.Context(0)
    from plfem_b200.solver_fem import TrueVectorialMaxwellSolver
    with pytest.raises(_cabi.PlfemError):
        TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, 4)
    # no module of the product imports the oracle
    pkg = os.path.join(ROOT, "pl-fem-vectoriel_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|scipy\.sparse\.linalg)|fem_oracle", src, re.M), fn


@pytest.mark.parametrize("case", ["small_case", "cfg1"])
def test_dof_tables_match_oracle(case, request):
    g, mesh = request.getfixturevalue(case)
    pb = _cabi.Problem(mesh, host_only=True)
    ed, loc, bnd, itr = pb.dofs()
    B = O.P2Basis(mesh)
    assert pb.N == B.N
    assert np.array_equal(ed, B.element_dofs)
    assert np.array_equal(loc, B.doflocs)                      # bit-identical, incl. "last element wins"
    assert np.array_equal(bnd, B.boundary_dofs())
    assert np.array_equal(itr, np.setdiff1d(np.arange(B.N), B.boundary_dofs()))
    assert np.array_equal(pb.quad_points(), B.x)
    assert pb.info.n_degenerate == 0


def test_degenerate_and_invalid_meshes():
    from plfem_b200.mesh import MeshTri
    flat = MeshTri(np.array([[0.0, 1.0, 2.0, 0.0], [0.0, 0.0, 0.0, 1.0]]), np.array([[0, 0], [1, 1], [2, 3]]))
    pb = _cabi.Problem(flat, host_only=True)
    assert pb.info.n_degenerate == 1

    class Bad:
        p = np.zeros((2, 3)); t = np.array([[0], [1], [7]])
    with pytest.raises(_cabi.PlfemError):
        _cabi.Problem(Bad, host_only=True)


def test_restart_eigensolver():
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 45, 105):
        a = rng.standard_normal((n, n)); a = a + a.T
        w, v = _cabi.symeig(a)
        assert np.allclose(w, np.linalg.eigvalsh(a), atol=1e-12 * max(1, n))
        assert np.abs(a @ v - v * w).max() < 1e-12 * n and np.abs(v.T @ v - np.eye(n)).max() < 1e-13 * n
    # arrowhead + tridiagonal, the shape it sees after a thick restart
    n, p = 45, 22
    a = np.diag(rng.standard_normal(n))
    a[p, :p] = a[:p, p] = 1e-4 * rng.standard_normal(p)
    for j in range(p, n - 1):
        a[j, j + 1] = a[j + 1, j] = rng.standard_normal()
    w, v = _cabi.symeig(a)
    assert np.allclose(w, np.linalg.eigvalsh(a), atol=1e-13) and np.abs(a @ v - v * w).max() < 1e-13


def test_convergence_check_eigensolver_returns_the_last_rows():
    """`symmetric_eigen_tail` (eigenvalues + the last p rows of the eigenvector matrix, what a Lanczos convergence check reads)
    against LAPACK: eigenvalues, and — sign- and rotation-invariant inside a degenerate cluster — the norm of the tail of every
    eigenvalue cluster; the p rows of an orthogonal matrix are orthonormal."""
    rng = np.random.default_rng(3)
    cases = [(1, 1), (2, 2), (5, 4), (44, 4), (70, 4), (45, 4)]
    for n, p in cases:
        a = rng.standard_normal((n, n)); a = a + a.T
        if n == 45:        # a thick-restart projection: diagonal block with degenerate pairs, arrow row, block tridiagonal rest
            q = 22
            a = np.diag(np.repeat(rng.standard_normal((n + 1) // 2), 2)[:n])
            a[q, :q] = a[:q, q] = 1e-4 * rng.standard_normal(q)
            for j in range(q, n - 1):
                a[j, j + 1] = a[j + 1, j] = rng.standard_normal()
        w, t = _cabi.symeig_tail(a, p)
        wr, vr = np.linalg.eigh(a)
        assert np.allclose(w, wr, atol=2e-13 * max(1.0, np.abs(wr).max()))
        assert np.abs(t @ t.T - np.eye(p)).max() < 1e-13
        for j in range(n):
            cl = np.where(np.abs(wr - wr[j]) < 1e-9)[0]
            assert abs(np.linalg.norm(t[:, cl]) - np.linalg.norm(vr[n - p:, cl])) < 1e-12
        wf, vf = _cabi.symeig(a)
        assert np.allclose(wf, wr, atol=2e-13 * max(1.0, np.abs(wr).max())) and np.abs(a @ vf - vf * wf).max() < 1e-12 * max(1.0, np.abs(wr).max())


@pytest.mark.parametrize("leaf,sn", [(24, 64), (8, 16)])
def test_front_plan_solves_the_shifted_system(small_case, leaf, sn):
    """The host-built plan (ordering, update sets, child maps), executed with dense NumPy, must solve
    (A - sigma B) x = b like SuperLU does."""
    from scipy.sparse.linalg import splu
    g, mesh = small_case
    s = O.interior_system(g, mesh)
    K = (s["A_int"] - O.sigma_estimate(g) * s["B_int"]).tocsr()
    pb = _cabi.Problem(mesh, host_only=True)
    pl = pb.plan(leaf, sn)
    n = pl["n"]
    assert n == len(s["interior"]) and sorted(pl["perm"]) == list(range(n))
    # structural invariants
    first, sz, sptr, strct, parent, level = (pl[k] for k in ("first", "s", "sptr", "strct", "parent", "level"))
    assert first[0] == 0 and np.array_equal(first[1:], np.cumsum(sz)[:-1]) and sz.sum() == n and sz.max() <= sn
    for f in range(pl["nfronts"]):
        st = strct[sptr[f]:sptr[f + 1]]
        assert (np.diff(st) > 0).all() and (len(st) == 0 or st[0] >= first[f] + sz[f])
        if parent[f] >= 0:
            assert parent[f] > f and level[parent[f]] > level[f]
        else:
            assert len(st) == 0
    Kp, idx = FR.permuted_operator(K, pl)
    fronts, ch = FR.factor(Kp, pl)
    b = np.random.default_rng(1).standard_normal(2 * n)
    x = FR.solve(fronts, ch, pl, b)
    x = x + FR.solve(fronts, ch, pl, b - Kp @ x)               # the one refinement step the product takes
    xr = splu(Kp.tocsc()).solve(b)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-8
    assert np.linalg.norm(Kp @ x - b) / np.linalg.norm(b) < 1e-9


def test_next_factorisation_spec():
    """Executable specification of the next factorisation on a coarse structured mesh, where the product's
    algebra loses six digits: symmetrised pivot-block inverses recover most of them, delayed pivots in static slots the
    rest — with a handful of handed-up unknowns and no front near its slot capacity."""
    import plfem_b200 as P
    import frontal_reference_delayed as FD
    from scipy.sparse.linalg import splu
    nx = 40
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
    xs = np.linspace(-32.0, 32.0, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    idx = np.arange((nx + 1) ** 2).reshape(nx + 1, nx + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    mesh = P.MeshTri(np.vstack([X.ravel(), Y.ravel()]), np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])]))
    s = O.interior_system(g, mesh)
    K = (s["A_int"] - O.sigma_estimate(g) * s["B_int"]).tocsr()
    pl = _cabi.Problem(mesh, host_only=True).plan()
    Kp, _ = FR.permuted_operator(K, pl)
    rhs = np.random.default_rng(1).standard_normal(Kp.shape[0])
    xr = splu(Kp.tocsc()).solve(rhs)

    def err(x):
        return np.linalg.norm(x - xr) / np.linalg.norm(xr)
    fr, ch = FR.factor(Kp, pl)
    e_product = err(FR.solve(fr, ch, pl, rhs))                       # 3e-7 when this was written
    fr, ch = FD.factor(Kp, pl, tau=1e30, symmetrise=True)
    e_sym = err(FD.solve(fr, ch, pl, rhs))                           # 3e-11
    fr, ch = FD.factor(Kp, pl, tau=30.0, symmetrise=True)
    e_full = err(FD.solve(fr, ch, pl, rhs))                          # 7e-13
    n_delayed = [len(f["dl"]) for f in fr]
    # the same with the product's own inverse algorithm (Gauss-Jordan as in invert_kernel) and its symmetrised write-back
    fr_gj, ch_gj = FR.factor(Kp, pl, inverse=lambda A: FR.gauss_jordan_inverse(A, symmetrise=True))
    e_gj = err(FR.solve(fr_gj, ch_gj, pl, rhs))
    A0 = np.random.default_rng(2).standard_normal((37, 37))
    A0 = A0 + A0.T
    A0[np.diag_indices(37)] *= 1e-3                                  # forces row swaps
    assert np.abs(FR.gauss_jordan_inverse(A0) - np.linalg.inv(A0)).max() < 1e-11 * np.abs(np.linalg.inv(A0)).max()
    assert e_gj < 1e-9 and e_gj < 1e-2 * e_product
    assert e_sym < 1e-9 and e_sym < 1e-2 * e_product
    assert e_full < 1e-11
    assert 0 < sum(n_delayed) < 0.01 * Kp.shape[0] and max(n_delayed) < FD.DC


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_front_plan_on_random_triangulations(seed):
    """Random point clouds (irregular valences, thin triangles at the hull, chords between boundary vertices): the plan built
    from the lifted vertex-graph dissection must be a valid elimination forest of the P2 graph — executed with dense NumPy it
    solves the shifted system."""
    import plfem_b200 as P
    from scipy.spatial import Delaunay
    from scipy.sparse.linalg import splu
    rng = np.random.default_rng(seed)
    n_pts = 350 + 150 * seed
    r = 6.0 * np.sqrt(rng.random(n_pts))
    a = 2 * np.pi * rng.random(n_pts)
    pts = np.column_stack([r * np.cos(a), r * np.sin(a)])
    tri = Delaunay(pts).simplices
    p = pts.T.copy()
    t = tri.T.copy()
    area2 = (p[0, t[1]] - p[0, t[0]]) * (p[1, t[2]] - p[1, t[0]]) - (p[0, t[2]] - p[0, t[0]]) * (p[1, t[1]] - p[1, t[0]])
    t = t[:, np.abs(area2) > 1e-9]
    mesh = P.MeshTri(p, t)
    g = P.MCFGeometry(3, 4.0, 1.2, 1.53, 1.0, 1.55)
    s = O.interior_system(g, mesh)
    K = (s["A_int"] - O.sigma_estimate(g) * s["B_int"]).tocsr()
    pl = _cabi.Problem(mesh, host_only=True).plan(12 + 6 * seed, 16 * (seed + 1))
    n = pl["n"]
    assert n == len(s["interior"]) and sorted(pl["perm"]) == list(range(n))
    first, sz, sptr, strct, parent, level = (pl[k] for k in ("first", "s", "sptr", "strct", "parent", "level"))
    assert sz.sum() == n and sz.max() <= 16 * (seed + 1)
    for f in range(pl["nfronts"]):
        st = strct[sptr[f]:sptr[f + 1]]
        assert (np.diff(st) > 0).all() and (len(st) == 0 or st[0] >= first[f] + sz[f])
        assert (parent[f] > f and level[parent[f]] > level[f]) if parent[f] >= 0 else len(st) == 0
    Kp, _ = FR.permuted_operator(K, pl)
    fronts, ch = FR.factor(Kp, pl)
    rhs = rng.standard_normal(2 * n)
    x = FR.solve(fronts, ch, pl, rhs)
    x = x + FR.solve(fronts, ch, pl, rhs - Kp @ x)
    xr = splu(Kp.tocsc()).solve(rhs)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-7


def test_front_plan_is_the_frozen_one(cfg1, small_case):
    """The numerical results on the GPU (and their bit-for-bit reproducibility) depend on the front plan.  Host-side speed
    work on the symbolic phase must not change it silently: these digests were taken from the plan the GPU parity tests of
    round 1 ran with.  An intended change of the ordering regenerates tests/golden/front_plan.json (same code as below)."""
    import hashlib
    import json

    def digest(mesh, opts):
        pl = _cabi.Problem(mesh, host_only=True).plan(*opts)
        h = hashlib.sha256()
        for k in ("perm", "first", "s", "parent", "level", "sptr", "strct", "cmap_ptr", "cmap", "foff"):
            h.update(np.ascontiguousarray(pl[k]).tobytes())
        return h.hexdigest()
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "front_plan.json")))
    assert digest(cfg1[1], (0, 0)) == gold["cfg1_default"]
    assert digest(cfg1[1], (8, 16)) == gold["cfg1_leaf8_sn16"]
    assert digest(small_case[1], (0, 0)) == gold["small_default"]


def test_host_analysis_digests_of_the_micro_benchmark(cfg1, small_case, tmp_path):
    """The stand-alone harness of the host analysis (scripts/micro/host_analysis.cpp: DOF tables, front plan and the MERGE of a
    forest's plans, which no C-ABI entry exposes without a GPU) must reproduce the digests taken before the host code was
    reworked for speed: the merged plan of a forest is what the device factorises."""
    import json
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_analysis"
    subprocess.run(["g++", "-O2", "-std=c++17", "-mavx2", "-ffp-contract=off", "-I" + os.path.join(root, "pl-fem-vectoriel_b200", "csrc"),
                    os.path.join(root, "scripts", "micro", "host_analysis.cpp"),
                    os.path.join(root, "pl-fem-vectoriel_b200", "csrc", "symbolic.cpp"), "-o", str(exe), "-lpthread"], check=True,
                   capture_output=True)
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "front_plan.json")))
    for name, mesh in (("small", small_case[1]), ("cfg1", cfg1[1])):
        f = tmp_path / (name + ".mesh")
        p_ = np.ascontiguousarray(mesh.p, dtype=np.float64)
        t_ = np.ascontiguousarray(mesh.t, dtype=np.int64)
        with open(f, "wb") as fh:
            np.array([p_.shape[1], t_.shape[1]], dtype=np.int64).tofile(fh)
            p_.tofile(fh)
            t_.tofile(fh)
        for threads in ("1", "4"):
            out = subprocess.run([str(exe), str(f), "2", threads], check=True, capture_output=True, text=True).stdout
            got = dict(re.findall(r"(dof|plan|merge) ([0-9a-f]{16})", out))
            assert got == gold["host_analysis_" + name], (name, threads, out)


def test_solve_iter_keeps_job_order_and_bounds_the_forests_in_flight():
    """`ForestPool.solve_iter` (host logic only: the forest solve is replaced): results come back in job order while at most
    workers + max(4, workers / 2) forests are submitted and not yet consumed, however slowly the first one finishes."""
    import threading
    import time
    from plfem_b200.batch import ForestPool, default_workers
    assert default_workers(16) >= 6 and default_workers(4) >= 6
    pool = ForestPool(batch=3, workers=2)
    lock = threading.Lock()
    state = {"started": 0, "consumed": 0, "max_ahead": 0}

    def fake_forest(jobs):
        with lock:
            state["started"] += 1
            state["max_ahead"] = max(state["max_ahead"], state["started"] - state["consumed"])
        time.sleep(0.2 if jobs[0] == 0 else 0.002)          # the first forest is the slow one
        return [j * 10 for j in jobs]
    pool.solve_forest = fake_forest
    try:
        out = []
        for r in pool.solve_iter(list(range(60))):
            out.append(r)
            if len(out) % 3 == 0:
                with lock:
                    state["consumed"] += 1
        assert out == [10 * j for j in range(60)]
        assert state["started"] == 20 and state["max_ahead"] <= 2 + 4
    finally:
        pool.close()


def test_pinned_pool_carves_blocks_out_of_slabs_and_merges_them_back(monkeypatch):
    """`_cabi.PinnedPool` (host logic; the page-locked allocation is replaced by plain buffers): blocks of many sizes never
    overlap, freed blocks merge back into whole slabs, no new slab is taken while a free piece fits, and beyond the cap the
    pool falls back to ordinary arrays."""
    import ctypes as C
    import gc

    class FakeLib:
        def __init__(self):
            self.bufs = []

        def plfem_host_alloc(self, nbytes, ref):
            b = (C.c_char * nbytes)()
            self.bufs.append(b)
            ref._obj.value = C.addressof(b)
            return 0
    fake = FakeLib()
    monkeypatch.setattr(_cabi, "load", lambda: fake)
    pool = _cabi.PinnedPool(cap_bytes=64 << 20)
    pool.SLAB = 16 << 20
    rng = np.random.default_rng(0)
    live = []
    for it in range(1500):
        if live and rng.random() < 0.5:
            live.pop(rng.integers(len(live)))
        else:
            live.append(pool.empty((int(rng.integers(1, 400000)),)))
            live[-1][:] = it
    spans = sorted((x.ctypes.data, x.ctypes.data + x.nbytes) for x in live)
    assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
    assert all((x == x[0]).all() for x in live)
    n_slabs = len(fake.bufs)
    assert 1 <= n_slabs <= 4
    live.clear()
    gc.collect()
    assert pool.in_use == 0 and len(pool.free) == n_slabs and all(f[1] == 16 << 20 for f in pool.free)
    a = pool.empty((1 << 20,))                       # 8 MiB: fits a free slab, no new allocation
    assert len(fake.bufs) == n_slabs
    big = [pool.empty((3 << 20,)) for _ in range(4)]       # 24 MiB each: own slabs until the cap, then ordinary arrays
    assert pool.reserved <= 64 << 20 and all(b.shape == (3 << 20,) for b in big)
    del a, big
