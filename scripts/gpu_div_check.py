"""Is the divergence-energy deviation on config 2 ours or ARPACK's (tol = 1e-7)?  Compare both with a tightly converged eigsh."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import fem_oracle as O
from plfem_b200.solver_fem import TrueVectorialMaxwellSolver, sigma_estimate
from scipy.sparse.linalg import eigsh
w, g, mesh = bench.make_case(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
n_modes = w["n_modes"]
modes, raw = TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, n_modes, return_raw=True)
s = O.interior_system(g, mesh)
it = s["interior"]; N = len(it)
Dxx, Dyy, Dxy = (s[k][it, :][:, it] for k in ("Dxx", "Dyy", "Dxy"))
sigma = sigma_estimate(g); k = n_modes + 12
def div_energy(V):
    out = []
    for v in V.T:
        v = v / np.linalg.norm(v); vx, vy = v[:N], v[N:]
        out.append(vx @ (Dxx @ vx) + 2 * vx @ (Dxy @ vy) + vy @ (Dyy @ vy))
    return np.array(out)
res = {}
for tol in (1e-7, 1e-13):
    lam, V = eigsh(s["A_int"], k=k, M=s["B_int"], sigma=sigma, which="LM", tol=tol, maxiter=12000, v0=np.ones(2 * N))
    o = np.argsort(lam); res[tol] = (lam[o], div_energy(V[:, o]))
ours = (raw["beta_sq"], raw["metrics"][:, 0])
for name, (lam, de) in (("ours", ours), ("eigsh tol 1e-7", res[1e-7])):
    ref_lam, ref_de = res[1e-13]
    print(f"{name:16s} vs eigsh tol 1e-13: beta^2 rel dev {np.abs(lam / ref_lam - 1).max():.2e}; div energy: max |d| / max(|ref|) {np.abs(de - ref_de).max() / np.abs(ref_de).max():.2e}"
          f", sum-over-all rel dev {abs(de.sum() - ref_de.sum()) / abs(ref_de.sum()):.2e}")
