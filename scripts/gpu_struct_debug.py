import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plfem_b200 as P
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate
from oracle import fem_oracle as O
from scipy.sparse.linalg import eigsh, splu

def smesh(nx, L=64.0):
    xs = np.linspace(-L / 2, L / 2, nx + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    p = np.vstack([X.ravel(), Y.ravel()])
    idx = np.arange((nx + 1) * (nx + 1)).reshape(nx + 1, nx + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    return P.MeshTri(p, np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])]))

g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
ctx = _cabi.Context.get(0)
mat, keep = _cabi.material_struct(g)
sigma = sigma_estimate(g)
for nx in [int(a) for a in sys.argv[1:]] or [24, 48, 96, 144]:
    mesh = smesh(nx)
    pb = _cabi.Problem(mesh, ctx)
    pl = pb.plan()
    u = np.diff(pl["sptr"])
    try:
        vals, vecs, met, ncore, st = pb.solve_modes(mat, sigma, 22, want_vectors=False)
        msg = f"ok resid={st.max_residual:.2e} block_ops={st.n_block_op} fac={st.ms_factor:.1f}ms lan={st.ms_lanczos:.1f}ms"
    except Exception as e:
        vals = None
        msg = "FAILED " + str(e)[:100]
    line = f"nx={nx} dim={2 * pb.n_interior} levels={pl['nlevels']} max_s={pl['s'].max()} max_u={u.max()} {msg}"
    if nx <= 60:
        s = O.interior_system(g, mesh)
        ref = np.sort(eigsh(s["A_int"], k=22, M=s["B_int"], sigma=sigma, which="LM", tol=1e-9)[0])
        if vals is not None:
            line += f" dev_vs_eigsh={np.abs(vals / ref - 1).max():.2e}"
        # raw solve check
        K = (s["A_int"] - sigma * s["B_int"]).tocsc()
        b = s["B_int"] @ np.ones(K.shape[0])
        xr = splu(K).solve(b)
        try:
            for refine in (100, 101):
                x = pb.debug_solve(sigma, b, refine)
                line += f" solve_err(refine={refine - 100})={np.linalg.norm(x - xr) / np.linalg.norm(xr):.2e}"
        except Exception as e:
            line += " debug_solve failed " + str(e)[:60]
    print(line, flush=True)
    pb.close()
