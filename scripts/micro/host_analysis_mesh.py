"""Writes the mesh of a bench workload (cfg1, cfg2, cfg5) or of the tests' small case as the binary file
scripts/micro/host_analysis.cpp reads: int64 V, T; float64 p (2, V); int64 t (3, T)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import plfem_b200 as P  # noqa: E402


def mesh_of(name):
    if name == "small":
        return P.MeshGenerator.generate(P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55), refinement=0.4)[0]
    import bench
    return bench.make_case(name)[2]


if __name__ == "__main__":
    name, out = sys.argv[1], sys.argv[2]
    m = mesh_of(name)
    p = np.ascontiguousarray(m.p, dtype=np.float64)
    t = np.ascontiguousarray(m.t, dtype=np.int64)
    with open(out, "wb") as f:
        np.array([p.shape[1], t.shape[1]], dtype=np.int64).tofile(f)
        p.tofile(f)
        t.tofile(f)
    print(out, p.shape, t.shape)
