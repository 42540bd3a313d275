"""One-off diagnostics on the GPU box: parity numbers and phase timings for a named case."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plfem_b200 as P
from plfem_b200 import _cabi
from plfem_b200.solver_fem import TrueVectorialMaxwellSolver
from oracle import fem_oracle as O

case = sys.argv[1] if len(sys.argv) > 1 else "small"
if case == "small":
    g = P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55); mesh, _ = P.MeshGenerator.generate(g, refinement=0.4); nm = 4
else:
    g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55); mesh, _ = P.MeshGenerator.generate(g); nm = 10
print("mesh", mesh.p.shape, mesh.t.shape, flush=True)
s = TrueVectorialMaxwellSolver(g)
t = time.time(); A, B, basis, Dxx, Dyy, Dxy, Minv = s.assemble_hfield_system(mesh); print("gpu assemble+export s", time.time() - t, flush=True)
rA, rB, rbasis, rDxx, rDyy, rDxy, rMinv = O.assemble_hfield_system(g, mesh)
for name, M, R in (("A", A, rA), ("B", B, rB), ("Dxx", Dxx, rDxx), ("Dyy", Dyy, rDyy), ("Dxy", Dxy, rDxy), ("Minv", Minv, rMinv)):
    same = M.shape == R.shape and np.array_equal(M.indptr, R.indptr) and np.array_equal(M.indices, R.indices)
    if same:
        d = np.abs(M.data - R.data)
        print(name, "structure same nnz", M.nnz, "max abs dev", d.max(), "max entry-rel dev", (d / np.maximum(np.abs(R.data), 1e-300)).max(), "bit-equal frac", (M.data == R.data).mean(), flush=True)
    else:
        print(name, "STRUCTURE DIFFERS", M.nnz, R.nnz, "symdiff", abs((M != 0).astype(int) - (R != 0).astype(int)).nnz, flush=True)
for rep in range(3):
    t = time.time(); modes, raw = s.solve_vectorial_modes(mesh, nm, return_raw=True); dt = time.time() - t
    print("solve wall s", dt, json.dumps(raw["stats"]), flush=True)
t = time.time(); rmodes, rraw = O.solve_vectorial_modes(g, mesh, nm, return_raw=True); print("oracle solve s", time.time() - t)
print("beta_sq rel dev", np.abs(raw["beta_sq"] / rraw["beta_sq"] - 1).max())
print("gpu  ", raw["beta_sq"][:6]); print("ref  ", rraw["beta_sq"][:6])
print("nmodes", len(modes), len(rmodes))
for m, r in list(zip(modes, rmodes))[:5]:
    v, rv = np.concatenate([m["Ex_dofs"], m["Ey_dofs"]]), np.concatenate([r["Ex_dofs"], r["Ey_dofs"]])
    print(m["n_eff"], r["n_eff"], "dot", abs(v @ rv), "conf", m["confinement"], r["confinement"], "div", m["div_ratio"], r["div_ratio"], m["polarization"], r["polarization"])
for opts in (dict(block=1), dict(block=4, ncv=48), dict(block=4, ncv=64), dict(block=4, ncv=96), dict(refine=-1), dict(refine=2), dict(leaf_nodes=16), dict(leaf_nodes=48)):
    modes2, raw2 = s.solve_vectorial_modes(mesh, nm, return_raw=True, **opts)
    print(opts, "dev", np.abs(raw2["beta_sq"] / rraw["beta_sq"] - 1).max(), json.dumps(raw2["stats"]), flush=True)
