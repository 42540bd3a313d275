"""Accuracy and cost of relaxing the operator accuracy late in the Lanczos run (PLFEM_RELAX_AT)."""
import os, sys
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
    vals, vecs, met, ncore, st = pb.solve_modes(mat, sigma, k)
    s = st.as_dict()
    print(cfg, "RELAX_AT", os.environ.get("PLFEM_RELAX_AT"), "block_ops", s["n_block_op"], "lanczos ms", round(s["ms_lanczos"], 2), "max_residual", s["max_residual"], flush=True)
    np.savez(f"/tmp/relax_{cfg}_{os.environ.get('PLFEM_RELAX_AT', '0')}.npz", vals=vals, vecs=vecs, met=met)
    ref = f"/tmp/relax_{cfg}_0.npz"
    if os.path.exists(ref) and os.environ.get("PLFEM_RELAX_AT", "0") != "0":
        r = np.load(ref)
        print("   vs accurate: beta^2 rel dev max", np.abs(vals / r["vals"] - 1).max(), " 1-|<x,x0>| max", np.abs(1 - np.abs(np.sum(vecs * r["vecs"], axis=1))).max(),
              " metrics dev max", np.abs(met[:, :7] - r["met"][:, :7]).max())
