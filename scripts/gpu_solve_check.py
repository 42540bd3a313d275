"""GPU factorisation vs SuperLU on config 1: residuals by row class (sliver DOFs vs the rest)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.sparse.linalg import splu
import plfem_b200 as P
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate
from oracle import fem_oracle as O
g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55); mesh, _ = P.MeshGenerator.generate(g)
s = O.interior_system(g, mesh); sigma = sigma_estimate(g)
K = (s["A_int"] - sigma * s["B_int"]).tocsr(); B = s["B_int"]
pb = _cabi.Problem(mesh); mat, keep = _cabi.material_struct(g)
vals, vecs, met, nc, st = pb.solve_modes(mat, sigma, 22)
rowmax = np.asarray(abs(K).max(axis=1).todense()).ravel(); big = rowmax > 1e6
print("rows with huge entries:", big.sum())
rng = np.random.default_rng(0)
b = B @ rng.standard_normal(K.shape[0])
xr = splu(K.tocsc()).solve(b)
for refine in (0, 1, 2, 100, 101):   # >= 100: per-level kernels instead of the persistent operator kernel
    x = pb.debug_solve(sigma, b, refine)
    r = K @ x - b
    print(f"refine {refine}: rel err vs splu {np.linalg.norm(x-xr)/np.linalg.norm(xr):.2e} resid all {np.linalg.norm(r)/np.linalg.norm(b):.2e} "
          f"resid big rows {np.abs(r[big]).max():.2e} resid other rows {np.abs(r[~big]).max():.2e} | err at big rows {np.abs(x-xr)[big].max():.2e} |x| there {np.abs(xr[big]).max():.2e}")
rr = K @ xr - b; print(f"splu: resid big rows {np.abs(rr[big]).max():.2e} other {np.abs(rr[~big]).max():.2e}")
# eigenvector residuals by row class
A = s["A_int"]
for i in (0, 10, 21):
    x = vecs[i]; r = A @ x - vals[i] * (B @ x)
    print(i, "eigres big rows", np.abs(r[big]).max(), "other", np.abs(r[~big]).max(), "|x| big", np.abs(x[big]).max())
