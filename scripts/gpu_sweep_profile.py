"""One forest of config-1 designs, then the isolated sweeps of plfem_profile_last once more (L2 flushed):
run under `ncu --metrics gpu__time_duration.sum -k regex:forward|backward` to get the per-launch times of a sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from plfem_b200 import _cabi
from plfem_b200.solver_fem import sigma_estimate

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 12
w, g, mesh = bench.make_case(name)
ctx = _cabi.Context.get(0)
sigma = sigma_estimate(g)
mat, keep = _cabi.material_struct(g)
pbs = [_cabi.Problem(mesh, ctx) for _ in range(nb)]
k = min(w["n_modes"] + 12, 2 * pbs[0].n_interior - 4)
out = _cabi.solve_modes_batch(ctx, pbs, [mat] * nb, [sigma] * nb, [k] * nb, want_vectors=False)
st = out[0][4].as_dict()
print("forest", nb, "block_ops", st["batch_block_ops"], "refine", st["refine_steps"], "launches", st["kernel_launches"], "resid", st["max_residual"], flush=True)
print("PROFILE_BEGIN", flush=True)
rng = os.environ.get("NCU_RANGE")          # under `ncu --profile-from-start off`: capture only the isolated sweeps below
if rng:
    import torch
    torch.cuda.cudart().cudaProfilerStart()
prof, nbp = ctx.profile_last(repeat=int(os.environ.get("REPEAT", 1)))
for n_, (ms, nbytes) in prof.items():
    print(f"  {n_:22s} {ms:9.4f} ms {nbytes / 1e6:9.1f} MB {nbytes / ms / 1e6:8.1f} GB/s ({nbytes / ms / 1e6 / 6553.6:.3f})", flush=True)
if rng:
    torch.cuda.cudart().cudaProfilerStop()
