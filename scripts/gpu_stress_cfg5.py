"""Config 5 of SURVEY.md 8(d): structured stress mesh (unit cells split on the same diagonal), cfg-1 core discs.
Reports the assembly / SpMV bandwidths and the factorisation rate at ~2M unknowns (real arithmetic)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import plfem_b200 as P
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 500
k = 22
g = P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55)
L = 64.0
xs = np.linspace(-L / 2, L / 2, nx + 1)
X, Y = np.meshgrid(xs, xs, indexing="xy")
p = np.vstack([X.ravel(), Y.ravel()])
idx = np.arange((nx + 1) * (nx + 1)).reshape(nx + 1, nx + 1)
a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
t = np.hstack([np.vstack([a, b, d]), np.vstack([a, d, c])])
mesh = P.MeshTri(p, t)
print(f"mesh V={p.shape[1]} T={t.shape[1]}", flush=True)
ctx = _cabi.Context.get(0)
t0 = time.perf_counter(); pb = _cabi.Problem(mesh, ctx); t1 = time.perf_counter()
print(f"problem: N={pb.N} interior={pb.n_interior} dim={2 * pb.n_interior}  dof tables {t1 - t0:.2f} s", flush=True)
mat, keep = _cabi.material_struct(g)
sigma = sigma_estimate(g)
out = {}
for rep in range(2):
    t0 = time.perf_counter()
    vals, vecs, met, ncore, st = pb.solve_modes(mat, sigma, k, want_vectors=False)
    dt = time.perf_counter() - t0
    s = st.as_dict()
    print(f"solve {rep}: {dt:.2f} s  n_eff[0..3]={np.sqrt(vals[-3:]) / g.k0}  " + json.dumps({kk: (round(v, 3) if isinstance(v, float) else v) for kk, v in s.items()}), flush=True)
prof = pb.profile_kernels(mat, sigma, repeat=3)
peak = 6553.6
for name, (ms, nbytes) in prof.items():
    print(f"{name:22s} {ms:10.3f} ms  {nbytes / 1e6:10.1f} MB  {nbytes / ms / 1e6:8.1f} GB/s  ({nbytes / ms / 1e6 / peak:.3f} of HBM peak)")
print(f"factorisation: {s['factor_flops'] / 1e9:.1f} GFLOP in {s['ms_factor']:.1f} ms = {s['factor_flops'] / s['ms_factor'] / 1e9:.2f} TFLOP/s FP64")
