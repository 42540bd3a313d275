"""Parametric sweep over designs, one design per GPU at a time, one gather at the end.

The reference advertises a dataset generator (2 000 LHS designs x 12 layouts x 4 bands,
`README.md:193-202, 226-243`) whose driver is absent from its checkout; what exists is the
per-design hot path.  Designs are independent, so the multi-GPU shape is: shard the design list
round-robin over ranks (one process per GPU), solve locally, and exchange the fixed-width float64
records ONCE (`all_gather`).  There is no data-path collective inside a solve.

Record layout (86 float64 slots, the count `README.md:49` names; the README's own table sums to 85,
slot 85 is spare).  The 12 loss columns come from `losses.LossCalculator` (the reference's closed forms for
vectorial modes, `losses.py:742-826`, both directions) evaluated on the host from the mode records.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from .config import IPDipCauchy
from .geometry import MCFGeometry, SAMPLING_WEIGHTS
from .mesh import MeshGenerator

POL_CODE = {"TE-like": 0.0, "HE-like": 1.0, "Hybrid": 2.0, "EH-like": 3.0, "TM-like": 4.0}
N_PER_MODE = 7

RECORD_FIELDS: List[str] = (
    ["sample_id", "success", "solver_type", "solver_time_s"]                                     # 4 metadata
    + ["n_cores", "core_radius_um", "pitch_um", "variant", "has_central_core", "n_peripheral", "R_ring_um",
       "pitch_ratio", "cladding_radius_um", "domain_radius_um", "packing_efficiency", "n_vertices",
       "n_dofs"]                                                                                 # 13 geometry
    + ["V_number", "n_core_lambda", "n_clad", "wavelength_nm", "NA", "k0"]                       # 6 materials/optics
    + ["n_modes_found", "n_eff_mean", "n_eff_max", "n_eff_min", "n_eff_spread", "confinement_mean",
       "confinement_min", "confinement_max", "div_ratio_mean", "sigma_shift"]                    # 10 global modal
    + ["PDL_mean_dB", "PDL_max_dB", "n_hybrid_modes", "n_te_like_modes", "n_tm_like_modes"]      # 5 polarization
    + [f"loss_{k}" for k in ("IL_mux_dB", "MDL_mux_dB", "PDL_mux_dB", "XT_mux_dB", "IL_demux_dB", "MDL_demux_dB",
                             "PDL_demux_dB", "XT_demux_dB", "radiation_dB_per_m", "taper_dB", "mmf_dB",
                             "polymer_dB")]                                                      # 12 losses
    + [f"{name}_mode_{k}" for k in range(N_PER_MODE) for name in ("n_eff", "conf", "PDL", "pol", "div")]  # 35
    + ["spare"]
)
N_RECORD = len(RECORD_FIELDS)
assert N_RECORD == 86


def design_record(sample_id: int, design: Dict, geometry, mesh, modes: Sequence[Dict], stats: Optional[Dict],
                  seconds: float, success: bool = True) -> np.ndarray:
    r = np.full(N_RECORD, np.nan)
    f = {k: i for i, k in enumerate(RECORD_FIELDS)}
    r[f["sample_id"]], r[f["success"]], r[f["solver_type"]], r[f["solver_time_s"]] = sample_id, float(success), 1.0, seconds
    g = geometry
    if g is not None:
        r[f["n_cores"]], r[f["core_radius_um"]], r[f["pitch_um"]] = g.n_cores, g.r_core, g.pitch
        r[f["variant"]] = 1.0 if getattr(g, "config_type", "") == "pentagon_center_6" else 0.0
        r[f["has_central_core"]], r[f["n_peripheral"]], r[f["R_ring_um"]] = float(g.has_central_core), g.n_peripheral, g.R_ring
        r[f["pitch_ratio"]], r[f["cladding_radius_um"]], r[f["domain_radius_um"]] = g.pitch_ratio, g.cladding_radius, g.domain_radius
        r[f["packing_efficiency"]] = g.packing_efficiency
        r[f["V_number"]], r[f["n_core_lambda"]], r[f["n_clad"]] = g.V_number, g.n_core, g.n_clad
        r[f["wavelength_nm"]], r[f["k0"]] = 1000.0 * g.wavelength, g.k0
        r[f["NA"]] = math.sqrt(max(g.n_core ** 2 - g.n_clad ** 2, 0.0))
    if mesh is not None:
        r[f["n_vertices"]] = mesh.p.shape[1]
    if stats:
        r[f["n_dofs"]] = stats.get("n_dofs", np.nan)
        r[f["sigma_shift"]] = stats.get("sigma", np.nan)
    if success and modes:
        ne = np.array([m["n_eff"] for m in modes]); cf = np.array([m["confinement"] for m in modes])
        pdl = np.array([m["PDL_dB"] for m in modes]); dv = np.array([m["div_ratio"] for m in modes])
        pol = [m["polarization"] for m in modes]
        r[f["n_modes_found"]], r[f["n_eff_mean"]], r[f["n_eff_max"]], r[f["n_eff_min"]] = len(modes), ne.mean(), ne.max(), ne.min()
        r[f["n_eff_spread"]] = ne.max() - ne.min()
        r[f["confinement_mean"]], r[f["confinement_min"]], r[f["confinement_max"]] = cf.mean(), cf.min(), cf.max()
        r[f["div_ratio_mean"]] = dv.mean()
        r[f["PDL_mean_dB"]], r[f["PDL_max_dB"]] = pdl.mean(), pdl.max()
        r[f["n_hybrid_modes"]] = sum(p == "Hybrid" for p in pol)
        r[f["n_te_like_modes"]] = sum(p == "TE-like" for p in pol)
        r[f["n_tm_like_modes"]] = sum(p == "TM-like" for p in pol)
        for k, m in enumerate(modes[:N_PER_MODE]):
            r[f[f"n_eff_mode_{k}"]], r[f[f"conf_mode_{k}"]], r[f[f"PDL_mode_{k}"]] = m["n_eff"], m["confinement"], m["PDL_dB"]
            r[f[f"pol_mode_{k}"]], r[f[f"div_mode_{k}"]] = POL_CODE.get(m["polarization"], np.nan), m["div_ratio"]
        if g is not None:
            from .losses import LossCalculator, VectorialLossCalculator
            wl = 1000.0 * g.wavelength
            for direction in ("mux", "demux"):
                L = LossCalculator.calculate_physical_losses(list(modes), g, direction, wl)
                if L.get("success"):
                    r[f[f"loss_IL_{direction}_dB"]], r[f[f"loss_MDL_{direction}_dB"]] = L["IL_dB"], L["MDL_dB"]
                    r[f[f"loss_PDL_{direction}_dB"]], r[f[f"loss_XT_{direction}_dB"]] = L["PDL_dB"], L["crosstalk_dB"]
                    r[f["loss_radiation_dB_per_m"]] = L["radiation_loss_dB_per_m"]
            S = VectorialLossCalculator.calculate_vectorial_losses(list(modes), g, LossCalculator._build_design_params(list(modes), g, wl), "mux", wl)
            if S.get("success"):
                r[f["loss_taper_dB"]], r[f["loss_mmf_dB"]], r[f["loss_polymer_dB"]] = S["IL_taper"], S["IL_MMF"], S["IL_polymer"]
    elif success:
        r[f["n_modes_found"]] = 0
    return r


def band_sweep_designs(n_cores: int = 7, core_radius_um: float = 1.5, pitch_um: float = 8.0,
                       wavelengths_nm=(1490, 1550, 1600, 1650), n_modes: int = 10) -> List[Dict]:
    """Config 3 of BASELINE.json: S/C/L/U bands with the Cauchy index on one geometry (one mesh)."""
    return [dict(n_cores=n_cores, core_radius_um=core_radius_um, pitch_um=pitch_um, wavelength_nm=w,
                 n_core=IPDipCauchy.n(w), n_clad=1.0, n_modes=n_modes, variant=None) for w in wavelengths_nm]


def lhs_designs(n_samples: int, seed: int = 42, wavelengths_nm=(1490, 1550, 1600, 1650)) -> List[Dict]:
    """Config 4: stratified scrambled LHS over (core radius, pitch) per layout, weighted like
    `geometry_unified.py:702-705`, rejected by ``validate()`` (`:351-363`) — SURVEY.md 8(d)."""
    from scipy.stats import qmc
    layouts = sorted(SAMPLING_WEIGHTS)
    w = np.array([SAMPLING_WEIGHTS[n] for n in layouts], dtype=float)
    counts = np.floor(w / w.sum() * n_samples).astype(int)
    counts[np.argmax(w)] += n_samples - counts.sum()
    out: List[Dict] = []
    for n, cnt in zip(layouts, counts):
        variants = [None, "pentagon_center"] if n == 6 else [None]
        sampler = qmc.LatinHypercube(d=2, seed=seed + n)
        tries = 0
        got = 0
        while got < cnt and tries < 20:
            pts = qmc.scale(sampler.random(max(cnt - got, 1) * 2), [0.5, 3.0], [3.0, 15.0])
            tries += 1
            for r_um, p_um in pts:
                if got >= cnt:
                    break
                lam = wavelengths_nm[len(out) % len(wavelengths_nm)]
                d = dict(n_cores=n, core_radius_um=float(r_um), pitch_um=float(p_um), wavelength_nm=lam,
                         n_core=IPDipCauchy.n(lam), n_clad=1.0, n_modes=min(3 * n, 40),
                         variant=variants[got % len(variants)])
                try:
                    ok, _ = design_geometry(d).validate()
                except ValueError:
                    ok = False
                if ok:
                    out.append(d)
                    got += 1
    return out


def design_geometry(d: Dict) -> MCFGeometry:
    return MCFGeometry(d["n_cores"], d["pitch_um"], d["core_radius_um"], d["n_core"], d.get("n_clad", 1.0),
                       d["wavelength_nm"] / 1000.0, variant=d.get("variant"))


def shard(n_items: int, rank: int, world: int) -> List[int]:
    """Static round-robin partition of design indices."""
    return list(range(rank, n_items, world))


def solve_design_gpu(d: Dict, device: int = 0, refinement: float = 1.0):
    """Default per-design worker: mesh on the host, modal solve on the GPU."""
    import time
    from .solver_fem import TrueVectorialMaxwellSolver, sigma_estimate
    g = design_geometry(d)
    mesh, _ = MeshGenerator.generate(g, refinement)
    t0 = time.perf_counter()
    solver = TrueVectorialMaxwellSolver(g, device=device)
    try:
        modes = solver.solve_vectorial_modes(mesh, d.get("n_modes", 10))
    finally:
        solver.close()
    st = dict(solver.last_stats, sigma=sigma_estimate(g), n_dofs=2 * len(modes[0]["Ex_dofs"]) if modes else np.nan)
    return g, mesh, modes, st, time.perf_counter() - t0


def prepare_design(d: Dict, refinement: float = 1.0):
    """Host side of one design: geometry + mesh (Qhull).  Returns ``(geometry, mesh, n_modes)`` or the Exception."""
    try:
        g = design_geometry(d)
        mesh, _ = MeshGenerator.generate(g, refinement)
        return g, mesh, d.get("n_modes", 10)
    except Exception as e:                                  # noqa: BLE001 — a bad design must not stop the forest
        return e


def solve_forest_gpu(ds: Sequence[Dict], pool, refinement: float = 1.0, prepared: Optional[Sequence] = None) -> List:
    """Default worker of the forest mode: meshes on the host, ONE forest solve on the GPU for all of ``ds``.
    ``prepared[j]`` (optional) is the result of `prepare_design(ds[j])` or a Future of it (meshing pipelined on other threads).
    Returns one ``(geometry, mesh, modes, stats, seconds)`` or ``Exception`` per design."""
    import time
    from .solver_fem import sigma_estimate
    ready, out = [], [None] * len(ds)
    for j, d in enumerate(ds):
        p = prepared[j] if prepared is not None else prepare_design(d, refinement)
        p = p.result() if hasattr(p, "result") else p
        if isinstance(p, Exception):
            out[j] = p
        else:
            ready.append((j,) + tuple(p))
    if ready:
        t0 = time.perf_counter()
        res = pool.solve_forest([(g, mesh, n) for _, g, mesh, n in ready])
        secs = (time.perf_counter() - t0) / len(ready)
        stats = list(getattr(pool._local, "last_stats", None) or pool.last_stats)
        for (j, g, mesh, _), modes, st in zip(ready, res, stats):
            if isinstance(modes, Exception):
                out[j] = modes
            else:
                n_dofs = 2 * len(modes[0]["Ex_dofs"]) if modes and modes[0]["Ex_dofs"] is not None else np.nan
                out[j] = (g, mesh, modes, dict(st or {}, sigma=sigma_estimate(g), n_dofs=n_dofs), secs)
    return out


def run_sweep(designs: Sequence[Dict], rank: int = 0, world: int = 1, device: int = 0,
              solve_fn: Optional[Callable] = None, gather: bool = True, forest: int = 0,
              forest_fn: Optional[Callable] = None, workers: int = 0, mesh_threads: int = 4) -> np.ndarray:
    """Solve this rank's shard and return ALL records, (len(designs), 86), on every rank.

    ``forest = B > 0`` is the production mode on GPUs: the shard is cut into forests of ``B`` designs, each solved by
    one `plfem_solve_modes_batch` call.  With the default worker the forests are pipelined: ``mesh_threads`` host threads
    build geometries and Delaunay meshes ahead, ``workers`` forest threads (0 = 6: the device arenas of six contexts hold forests of the largest LHS designs — 19 cores, k = 52 —
    on one B200, `batch.default_workers()` = 12 suits 7-core designs; each with its own CUDA context) take the
    forests in order, so meshing, the host analysis of one forest and the device work of another overlap; designs of a
    forest on an identical mesh (the bands of a wavelength sweep) share one analysis.  A custom
    ``forest_fn(list of designs) -> list of results or Exceptions`` is called forest by forest.
    ``forest = 0`` solves design by design with ``solve_fn``.

    A failed design yields a record with ``success = 0`` and never poisons the others
    (the reference wraps each sample in try/except, `main.py:346,384-386`).
    With ``world > 1`` ``torch.distributed`` must be initialised (NCCL on GPUs, gloo in CPU tests).
    """
    mine = shard(len(designs), rank, world)
    local = np.full((len(mine), N_RECORD), np.nan)

    def put(j0, idx, results):
        for j, i, r in zip(range(j0, j0 + len(idx)), idx, results):
            if isinstance(r, Exception) or r is None:
                local[j] = design_record(i, designs[i], None, None, [], None, 0.0, False)
            else:
                g, mesh, modes, st, secs = r
                local[j] = design_record(i, designs[i], g, mesh, modes, st, secs, True)

    if forest > 0 and forest_fn is None:
        from concurrent.futures import ThreadPoolExecutor
        from .batch import ForestPool
        chunks = [(j0, mine[j0:j0 + forest]) for j0 in range(0, len(mine), forest)]
        with ForestPool(device=device, batch=forest, workers=workers if workers > 0 else 6) as pool, \
                ThreadPoolExecutor(max_workers=max(1, mesh_threads), thread_name_prefix="plfem-mesh") as mesher:
            prepared = {i: mesher.submit(prepare_design, designs[i]) for _, idx in chunks for i in idx}   # in sweep order

            def one(chunk):
                j0, idx = chunk
                try:
                    return solve_forest_gpu([designs[i] for i in idx], pool, prepared=[prepared[i] for i in idx])
                except Exception as e:                      # noqa: BLE001 — the whole forest failed
                    return [e] * len(idx)
            for (j0, idx), results in zip(chunks, pool._pool.map(one, chunks)):
                put(j0, idx, results)
    elif forest > 0:
        for j0 in range(0, len(mine), forest):
            idx = mine[j0:j0 + forest]
            try:
                results = forest_fn([designs[i] for i in idx])
            except Exception as e:                          # noqa: BLE001 — the whole forest failed
                results = [e] * len(idx)
            put(j0, idx, results)
    else:
        solve_fn = solve_fn or (lambda d: solve_design_gpu(d, device))
        for j, i in enumerate(mine):
            try:
                g, mesh, modes, st, secs = solve_fn(designs[i])
                local[j] = design_record(i, designs[i], g, mesh, modes, st, secs, True)
            except Exception:                                   # noqa: BLE001 — record and move on
                local[j] = design_record(i, designs[i], None, None, [], None, 0.0, False)
    if world == 1 or not gather:
        out = np.full((len(designs), N_RECORD), np.nan)
        out[mine] = local
        return out
    return gather_records(local, len(designs), rank, world, device)


def gather_records(local: np.ndarray, n_total: int, rank: int, world: int, device: int = 0) -> np.ndarray:
    """The single collective of the sweep: all_gather of equally padded record blocks."""
    import torch
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    buf = torch.full((per, N_RECORD), float("nan"), dtype=torch.float64, device=dev)
    if len(local):
        buf[: len(local)] = torch.from_numpy(local).to(dev)
    allbuf = torch.empty((world, per, N_RECORD), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(allbuf.view(-1), buf.view(-1))
    allbuf = allbuf.cpu().numpy()
    out = np.full((n_total, N_RECORD), np.nan)
    for r in range(world):
        idx = shard(n_total, r, world)
        out[idx] = allbuf[r, : len(idx)]
    return out


def records_to_csv(records: np.ndarray, path: str):
    with open(path, "w") as fh:
        fh.write(",".join(RECORD_FIELDS) + "\n")
        for row in records:
            fh.write(",".join("" if np.isnan(v) else repr(float(v)) for v in row) + "\n")
