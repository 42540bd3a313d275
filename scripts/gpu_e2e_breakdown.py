"""Where does the end-to-end time of one forest go on the host side? (single worker, config 1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from plfem_b200 import _cabi
from plfem_b200.batch import ForestPool
from plfem_b200.solver_fem import TrueVectorialMaxwellSolver, modes_from_solution, sigma_estimate
from concurrent.futures import ThreadPoolExecutor
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w, g, mesh = bench.make_case("cfg1")
ctx = _cabi.Context.get(0)
tp = ThreadPoolExecutor(B)
for rep in range(3):
    t = [time.perf_counter()]
    pbs = list(tp.map(lambda _: _cabi.Problem(mesh, ctx), range(B))); t.append(time.perf_counter())
    mats = [_cabi.material_struct(g)[0] for _ in range(B)]; sig = [sigma_estimate(g)] * B; ks = [22] * B; t.append(time.perf_counter())
    res = _cabi.solve_modes_batch(ctx, pbs, mats, sig, ks, want_vectors=True); t.append(time.perf_counter())
    out = [modes_from_solution(g, pb.n_interior, r[0], r[1], r[2], r[3]) for pb, r in zip(pbs, res)]; t.append(time.perf_counter())
    for pb in pbs: pb.close()
    t.append(time.perf_counter())
    st = res[0][4]
    print(f"B={B}: problems {1e3*(t[1]-t[0]):.1f} ms | materials {1e3*(t[2]-t[1]):.1f} | batch call {1e3*(t[3]-t[2]):.1f} (lib total {st.ms_total:.1f}: sym_wall {st.ms_symbolic_wall:.1f} asm {st.ms_assemble:.1f} fac {st.ms_factor:.1f} lan {st.ms_lanczos:.1f} met {st.ms_metrics:.1f}) | records {1e3*(t[4]-t[3]):.1f} | close {1e3*(t[5]-t[4]):.1f} | total {1e3*(t[5]-t[0]):.1f}", flush=True)
res = _cabi.solve_modes_batch(ctx, [_cabi.Problem(mesh, ctx) for _ in range(B)], mats, sig, ks, want_vectors=False)
print("without vectors: lib total", res[0][4].ms_total, "met", res[0][4].ms_metrics)
