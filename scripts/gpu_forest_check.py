"""Timing of forests of config-1 designs (resident problems, symbolic analysis redone per solve)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate

w, g, mesh = bench.make_case(sys.argv[1] if len(sys.argv) > 1 else "cfg1")
ctx = _cabi.Context.get(0)
sigma = sigma_estimate(g)
mat, keep = _cabi.material_struct(g)
sizes = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 16]
pbs = [_cabi.Problem(mesh, ctx) for _ in range(max(sizes))]
k = min(w["n_modes"] + 12, 2 * pbs[0].n_interior - 4)
ref = None
for nb in sizes:
    for rep in range(3):
        t0 = time.perf_counter()
        out = _cabi.solve_modes_batch(ctx, pbs[:nb], [mat] * nb, [sigma] * nb, [k] * nb, want_vectors=False,
                                      ncv=int(os.environ.get('NCV', 0)), leaf_nodes=int(os.environ.get('LEAF', 0)), max_sn_nodes=int(os.environ.get('MAXSN', 0)))
        dt = time.perf_counter() - t0
    st = out[0][4].as_dict()
    vals = out[0][0]
    if ref is None:
        ref = vals
    ok = all(o[5] == 0 for o in out)
    dev = max(float(np.abs(o[0] / ref - 1).max()) for o in out)
    print(f"nb={nb:3d} wall={1e3*dt:8.2f} ms  per-solve={1e3*dt/nb:7.2f} ms  {nb/dt:7.1f} solves/s | sym_wall={st['ms_symbolic_wall']:.1f} "
          f"sym_own={st['ms_symbolic']:.1f} asm={st['ms_assemble']:.2f} fac={st['ms_factor']:.2f} lan={st['ms_lanczos']:.2f} "
          f"met={st['ms_metrics']:.2f} entries={st['factor_entries']} levels={st['n_levels']} fronts={st['n_fronts']} launches={st['kernel_launches']} block_ops={st['batch_block_ops']} ok={ok} dev={dev:.1e} resid={st['max_residual']:.1e}", flush=True)
prof, nbp = ctx.profile_last(repeat=10)
peak = 6553.6
for name, (ms, nbytes) in prof.items():
    print(f"  {name:22s} {ms:9.4f} ms {nbytes / 1e6:9.1f} MB {nbytes / ms / 1e6:8.1f} GB/s ({nbytes / ms / 1e6 / peak:.3f} of HBM peak) [forest of {nbp}]", flush=True)
