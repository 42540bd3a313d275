"""Accuracy of the eigenpairs with and without iterative refinement inside the operator (config 1 / 2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate

for cfg in sys.argv[1:] or ["cfg1"]:
    w, g, mesh = bench.make_case(cfg)
    ctx = _cabi.Context.get(0)
    sigma = sigma_estimate(g)
    mat, keep = _cabi.material_struct(g)
    pb = _cabi.Problem(mesh, ctx)
    k = min(w["n_modes"] + 12, 2 * pb.n_interior - 4)
    res = {}
    for refine in (2, 1, -1):
        vals, vecs, met, ncore, st = pb.solve_modes(mat, sigma, k, refine=refine)
        res[refine] = (vals, vecs, met)
        s = st.as_dict()
        print(cfg, "refine", refine, "block_ops", s["n_block_op"], "lanczos ms", round(s["ms_lanczos"], 2), "max_residual", s["max_residual"], flush=True)
    v2, X2, m2 = res[2]
    for refine in (1, -1):
        v, X, m = res[refine]
        print(cfg, "refine", refine, "vs 2: beta^2 rel dev max", np.abs(v / v2 - 1).max(), " |1-|<x,x2>|| max", np.abs(1 - np.abs(np.sum(X * X2, axis=1))).max(),
              " metrics dev max", np.abs(m[:, :7] - m2[:, :7]).max(), flush=True)
        print("   per-mode beta^2 dev:", np.array2string(np.abs(v / v2 - 1), precision=1))
