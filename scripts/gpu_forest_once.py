"""Two forest solves of nb config-1 designs (for ncu: skip the launches of the first)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w, g, mesh = bench.make_case("cfg1")
ctx = _cabi.Context.get(0)
sigma = sigma_estimate(g); mat, keep = _cabi.material_struct(g)
pbs = [_cabi.Problem(mesh, ctx) for _ in range(nb)]
k = 22
for _ in range(reps):
    out = _cabi.solve_modes_batch(ctx, pbs, [mat] * nb, [sigma] * nb, [k] * nb, want_vectors=False)
    print(out[0][4].as_dict())
