"""Config 4 sample: LHS designs over the 12 layouts through the sweep driver in forest mode; a few against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from plfem_b200 import sweep
n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
designs = sweep.lhs_designs(n)
t0 = time.perf_counter()
rec = sweep.run_sweep(designs, forest=12)
dt = time.perf_counter() - t0
f = {k: i for i, k in enumerate(sweep.RECORD_FIELDS)}
ok = rec[:, f["success"]]
print(f"{n} designs in {dt:.2f} s ({n / dt:.1f}/s incl. meshing), success {int(ok.sum())}/{n}")
for i, d in enumerate(designs):
    print(i, d["n_cores"], d.get("variant"), round(d["core_radius_um"], 2), round(d["pitch_um"], 2), d["wavelength_nm"], "ok" if ok[i] else "FAILED",
          "modes", rec[i, f["n_modes_found"]], "nverts", rec[i, f["n_vertices"]], "neff_max", rec[i, f["n_eff_max"]])
# oracle comparison on three of them (the smallest meshes)
from oracle import fem_oracle as O
from plfem_b200.mesh import MeshGenerator
order = np.argsort(rec[:, f["n_vertices"]])[:3]
for i in order:
    d = designs[i]
    g = sweep.design_geometry(d)
    mesh, _ = MeshGenerator.generate(g, 1.0)
    modes = O.solve_vectorial_modes(g, mesh, d["n_modes"])
    ne = max(m["n_eff"] for m in modes)
    print("oracle check design", i, "n_eff_max", ne, "ours", rec[i, f["n_eff_max"]], "rel dev", abs(ne / rec[i, f["n_eff_max"]] - 1), "modes", len(modes), rec[i, f["n_modes_found"]])
