"""Per-forest timing of the end-to-end path with several worker threads (config 1)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from plfem_b200 import _cabi
from plfem_b200.batch import ForestPool
B, NW, NF = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
w, g, mesh = bench.make_case("cfg1")
pool = ForestPool(batch=B, workers=NW, want_vectors=(len(sys.argv) < 5 or sys.argv[4] != "novec"))
jobs = [(g, mesh, 10)] * B
log = []
orig = pool.solve_forest
def timed(js, *a, **k):
    t0 = time.perf_counter(); r = orig(js, *a, **k); t1 = time.perf_counter()
    st = pool.last_stats[0]
    log.append((t1 - t0, st["ms_total"], st["ms_symbolic_wall"], st["ms_lanczos"], st["ms_metrics"]))
    return r
pool.solve_forest = timed
pool.on_every_worker(lambda p_, c: p_.solve_forest(jobs)); pool.on_every_worker(lambda p_, c: p_.solve_forest(jobs))
log.clear()
t0 = time.perf_counter()
n = sum(1 for _ in pool.solve_iter(jobs * NF))
dt = time.perf_counter() - t0
a = np.array(log)
print(f"B={B} workers={NW}: {n / dt:.1f} solves/s; per forest: python wall {1e3 * a[:, 0].mean():.1f} ms, lib total {a[:, 1].mean():.1f}, sym_wall {a[:, 2].mean():.1f}, lanczos {a[:, 3].mean():.1f}, metrics {a[:, 4].mean():.1f}")
pool.close()
