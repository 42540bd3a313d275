"""Generate tests/golden/*.json|npz — run in the build container only (reads /root/reference).

What is frozen and where it comes from
  geometry.json   : the REAL reference module geometry_unified.py (imported from /root/reference):
                    core positions of the 12 layouts, domain radii, hashes, V numbers, eps samples.
  mesh_recipe.json: the REAL reference mesh.py::MeshGenerator._generate_mesh, run with stub
                    `skfem`/`geometry`/`config` modules (scikit-fem is not installable here; the stub
                    MeshTri only stores p and the column-sorted t).  Digests of p and t.
  oracle_cfg.json : outputs of this repo's oracle (oracle/fem_oracle.py: NumPy restatement of
                    scikit-fem + the real SciPy eigsh) on config 1 and on the small 3-core case:
                    sigma, eigenvalues, n_eff, CSR structure digests.  These pin the oracle against
                    regressions; they are NOT reference outputs (parity unpinned, see the oracle header).
  losses.json     : the REAL reference module losses.py (imported from /root/reference with a stand-in for the absent
                    `config.PhotonicLanternDesignParameters`: a class that stores its keyword arguments): its own
                    self-check (XT, PDL of 7 synthetic modes, `losses.py:1228-1259`) and
                    `LossCalculator.calculate_physical_losses` on synthetic vectorial mode lists for 3 geometries,
                    both directions, 2 wavelengths.  The mode lists are stored with the outputs.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_geometry():
    sys.path.insert(0, REF)
    import geometry_unified as G
    out = {}
    rng = np.random.default_rng(7)
    xs = rng.uniform(-30, 30, (2, 400))
    for n in G.MCFGeometry.SUPPORTED_N:
        for variant in ([None, "pentagon_center"] if n == 6 else [None]):
            g = G.MCFGeometry(n, 8.0, 1.5, 1.535, 1.0, variant=variant)
            eps = g.epsilon(xs[0], xs[1])
            out[f"{n}:{variant}"] = dict(positions=g.positions.tolist(), config_type=g.config_type,
                                         domain_radius=g.domain_radius, cladding_radius=g.cladding_radius,
                                         hash=g.hash, V_number=float(g.V_number), k0=float(g.k0),
                                         eps_real_digest=digest(np.real(eps)), eps_imag_digest=digest(np.imag(eps)))
    g = G.MCFGeometry(7, 8.0, 1.2, 1.53, 1.0)
    out["selfcheck"] = dict(V=float(g.V_number), eps00=float(np.real(g.epsilon(np.array([0.0]), np.array([0.0])))[0]),
                            eps_far=float(np.real(g.epsilon(np.array([100.0]), np.array([0.0])))[0]))
    out["eps_sample_points"] = xs.tolist()
    return out


def golden_mesh():
    """Import the reference mesh.py with stand-in modules and run its own recipe."""
    sys.path.insert(0, REF)
    import geometry_unified as G

    class StubMeshTri:
        def __init__(self, p, t):
            self.p = np.asarray(p, dtype=np.float64)
            self.t = np.sort(np.asarray(t), axis=0)

    class StubBasis:
        def __init__(self, mesh, elem):
            self.N = -1

    skfem = types.ModuleType("skfem"); skfem.Basis = StubBasis
    skfem_mesh = types.ModuleType("skfem.mesh"); skfem_mesh.MeshTri = StubMeshTri
    skfem_el = types.ModuleType("skfem.element"); skfem_el.ElementTriP2 = lambda: None
    geometry = types.ModuleType("geometry"); geometry.PhotonicLanternGeometry = G.PhotonicLanternGeometry
    config = types.ModuleType("config")

    class SimulationConfig:
        enable_mesh_cache = False; cache_max_size = 150; mesh_min_points = 0; mesh_target_points = 0
    config.SimulationConfig = SimulationConfig; config.PhysicalConstants = object
    saved = {k: sys.modules.get(k) for k in ("skfem", "skfem.mesh", "skfem.element", "geometry", "config")}
    sys.modules.update({"skfem": skfem, "skfem.mesh": skfem_mesh, "skfem.element": skfem_el,
                        "geometry": geometry, "config": config})
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_mesh", os.path.join(REF, "mesh.py"))
        ref_mesh = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_mesh)
        out = {}
        for n, ref in ((7, 1.0), (19, 1.0), (3, 0.4)):
            g = G.MCFGeometry(n, 8.0 if n != 3 else 6.0, 1.5 if n != 3 else 1.2, 1.535 if n != 3 else 1.53, 1.0)
            mesh, _ = ref_mesh.MeshGenerator._generate_mesh(g, ref, SimulationConfig())
            out[f"{n}:{ref}"] = dict(V=int(mesh.p.shape[1]), T=int(mesh.t.shape[1]), p_digest=digest(mesh.p),
                                     t_digest=digest(mesh.t.astype(np.int64)))
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def golden_oracle():
    import plfem_b200 as P
    from oracle import fem_oracle as O
    out = {}
    cases = {"cfg1": (P.MCFGeometry(7, 8.0, 1.5, 1.535, 1.0, 1.55), 1.0, 10),
             "small3": (P.MCFGeometry(3, 6.0, 1.2, 1.53, 1.0, 1.55), 0.4, 4)}
    for name, (g, ref, nm) in cases.items():
        mesh, _ = P.MeshGenerator.generate(g, refinement=ref)
        modes, raw = O.solve_vectorial_modes(g, mesh, nm, return_raw=True)
        s = raw["system"]
        out[name] = dict(V=int(mesh.p.shape[1]), T=int(mesh.t.shape[1]), N=int(s["basis"].N),
                         N_solve=int(len(s["interior"])), sigma=float(raw["sigma"]),
                         beta_sq=[float(v) for v in raw["beta_sq"]], n_eff=[m["n_eff"] for m in modes],
                         nnz_A_int=int(s["A_int"].nnz), nnz_B_int=int(s["B_int"].nnz),
                         A_int_indptr=digest(s["A_int"].indptr.astype(np.int64)),
                         A_int_indices=digest(s["A_int"].indices.astype(np.int64)),
                         B_int_indptr=digest(s["B_int"].indptr.astype(np.int64)),
                         B_int_indices=digest(s["B_int"].indices.astype(np.int64)),
                         A_abs_sum=float(np.abs(s["A_int"].data).sum()), B_sum=float(s["B_int"].data.sum()))
    return out


def synthetic_vectorial_modes(n_modes, seed):
    """Mode records shaped like the reference's own self-check (`losses.py:1234-1250`)."""
    rng = np.random.default_rng(seed)
    modes = []
    for k in range(n_modes):
        Px = float(rng.uniform(0.3, 0.7))
        Py = 1.0 - Px
        modes.append({"n_eff": float(1.20 - k * 0.003 + rng.normal(0, 1e-4)), "beta": float((2 * np.pi / 1.55) * (1.20 - k * 0.003)),
                      "P_x": Px, "P_y": Py, "PDL_dB": float(10 * np.log10(max(Px, Py) / min(Px, Py))), "polarization": "Hybrid",
                      "confinement": float(rng.uniform(0.55, 0.72)), "core_overlap": 0.60, "div_ratio": 0.02,
                      "is_vectorial": True, "method": "H-field_V18.10"})
    return modes


def golden_losses():
    sys.path.insert(0, REF)
    import geometry_unified as G

    class DesignParams:                      # stand-in for the absent config.PhotonicLanternDesignParameters
        def __init__(self, **kw):
            self.__dict__.update(kw)
    config = types.ModuleType("config"); config.PhotonicLanternDesignParameters = DesignParams
    saved = sys.modules.get("config")
    sys.modules["config"] = config
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_losses", os.path.join(REF, "losses.py"))
        L = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(L)
        out = {}
        m7 = synthetic_vectorial_modes(7, 42)
        out["selfcheck"] = dict(xt=L.EnhancedLossCalculator._calculate_crosstalk(m7), pdl=L.EnhancedLossCalculator._calculate_pdl_vectorial(m7))
        cases = []
        for n_cores, n_modes, seed in ((7, 7, 42), (19, 23, 5), (3, 2, 9)):
            g = G.MCFGeometry(n_cores, 8.0, 1.5, 1.535, 1.0)
            modes = synthetic_vectorial_modes(n_modes, seed)
            for direction in ("mux", "demux"):
                for wl in (1550.0, 1490.0):
                    res = L.LossCalculator.calculate_physical_losses(modes, g, direction, wl)
                    dp = L.LossCalculator._build_design_params(modes, g, wl)
                    cases.append(dict(n_cores=n_cores, seed=seed, n_modes=n_modes, direction=direction, wavelength_nm=wl,
                                      result={k: v for k, v in res.items()},
                                      design={k: (v if isinstance(v, (str, bool)) else float(v)) for k, v in dp.__dict__.items()}))
        out["cases"] = cases
        return out
    finally:
        if saved is None:
            sys.modules.pop("config", None)
        else:
            sys.modules["config"] = saved


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, fn in (("geometry", golden_geometry), ("mesh_recipe", golden_mesh), ("oracle_cfg", golden_oracle), ("losses", golden_losses)):
        if only and name not in only:
            continue
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
