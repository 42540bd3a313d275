#!/usr/bin/env python3
"""Headline benchmark: modal solves/sec of the vectorial H-field P2 FEM path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1|cfg2]

A *step* is ``--workers`` (default 12 for cfg1, 8 for cfg2, 6 for cfg4, 1 for cfg5) FORESTS of ``--inflight`` (default 12) independent
modal solves of the workload's cross-section each (cfg1: 144 solves): the designs of a forest are solved together as one block-diagonal problem by one C-ABI call
(`plfem_solve_modes_batch`), sharing every kernel launch — the sweep's production mode.  A modal
solve is `solve_vectorial_modes`: DOF tables -> assembly -> Dirichlet elimination -> ordering +
factorisation of A - sigma*B -> eigensolve -> per-mode reductions, all of it done per design (nothing
is reused between the designs of a forest).  ``--workers`` host threads each drive their
own forest, so the host-side symbolic analysis of one forest overlaps the device work of another.
The mesh is given (built on the host before timing, as in the reference where `MeshGenerator` runs
first).  ``latency`` in the JSON line is one solve run alone.

* ``value``  : solves/s with the meshes and their DOF tables already resident in HBM
               (`plfem_solve_modes_batch` on existing problems, symbolic analysis NOT reused,
               eigenvectors left on the device).
* ``e2e``    : the same metric through the public forest API with HOST buffers in and out
               (`ForestPool.solve_iter` on (geometry, NumPy mesh, n_modes) jobs: DOF tables, mesh upload,
               forest solve, eigenvectors + metrics copied back, mode records built and consumed);
               ``latency`` is `TrueVectorialMaxwellSolver(g).solve_vectorial_modes(mesh, n)` alone.
* ``host``   : process CPU time per solve inside the two timed regions (what limits N > 1 on one node).
* N > 1      : one process per GPU (torchrun), every rank solves the same workload (weak scaling,
               independent designs, no data-path collective); the 86-slot records are exchanged with
               ONE all_gather inside the timed region; time = max over ranks.
* ``--impl reference`` : the CPU path (oracle port: NumPy restatement of scikit-fem + the real SciPy
               eigsh/SuperLU) on the host cores, one design per worker process.

A forest's working set (GBs of front pools) exceeds the 126 MB L2; L2 is also flushed before the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[0] — the configuration the metric is quoted on ("7-core PL P2 mesh")
    "cfg1": dict(name="cfg1: 7-core hexagonal_1plus6_7 PL, r=1.5um, pitch 8um, n_core 1.535/air, 1550nm, n_modes=10",
                 n_cores=7, pitch=8.0, r=1.5, n_core=1.535, lam=1.55, n_modes=10),
    # BASELINE.json configs[1]
    "cfg2": dict(name="cfg2: 19-core hex_1plus6plus12 MCF, r=1.5um, pitch 8um, Cauchy IP-Dip/air, 1550nm, n_modes=40",
                 n_cores=19, pitch=8.0, r=1.5, n_core=None, lam=1.55, n_modes=40),
    # BASELINE.json configs[3]: a fixed sample of the stratified-LHS dataset (all 12 layouts, mixed mesh sizes and k)
    "cfg4": dict(name="cfg4: fixed sample of the 2,000-design stratified LHS (12 MCF layouts, r 0.5-3um, pitch 3-15um, 4 bands, "
                      "n_modes=min(3 N_cores, 40)), designs sorted into forests by mesh size"),
    # BASELINE.json configs[4]: the ~2M-unknown structured stress mesh (real arithmetic: the reference takes Re(eps))
    "cfg5": dict(name="cfg5: structured stress mesh, 500x500 cells split on one diagonal over a 64um square, cfg-1 cores, 1550nm, "
                      "n_modes=10 (dim 1,996,002)", n_cores=7, pitch=8.0, r=1.5, n_core=1.535, lam=1.55, n_modes=10, cells=500),
}


def make_case(name):
    import plfem_b200 as P
    w = WORKLOADS[name]
    n_core = w["n_core"] if w["n_core"] is not None else P.IPDipCauchy.n(1000 * w["lam"])
    g = P.MCFGeometry(w["n_cores"], w["pitch"], w["r"], n_core, 1.0, w["lam"])
    if "cells" in w:
        mesh = P.MeshTri.init_structured(w["cells"], w["cells"], 32.0)
    else:
        mesh, _ = P.MeshGenerator.generate(g, 1.0)
    return w, g, mesh


def make_jobs(name, n_jobs, forest):
    """The (geometry, mesh, n_modes) jobs of ONE step and how they are grouped into forests.  cfg1/cfg2/cfg5: n_jobs copies of
    the configuration's design.  cfg4: the first n_jobs designs of the seeded LHS sample, sorted by mesh size and cut into
    forests of `forest` designs — designs of similar size converge and finish together, which keeps the lockstep idle share low
    (size-sorted partition)."""
    import plfem_b200 as P
    if name != "cfg4":
        w, g, mesh = make_case(name)
        return w, [(g, mesh, w["n_modes"])] * n_jobs
    from plfem_b200 import sweep
    designs = sweep.lhs_designs(max(n_jobs * 2, 48), seed=42)
    designs = [designs[i] for i in np.random.default_rng(7).permutation(len(designs))[:n_jobs]]      # every layout, fixed order
    jobs = []
    for d in designs:
        g = sweep.design_geometry(d)
        mesh, _ = P.MeshGenerator.generate(g, 1.0)
        jobs.append((g, mesh, d["n_modes"]))
    jobs.sort(key=lambda j: j[1].p.shape[1] * (j[2] + 12))
    return WORKLOADS[name], jobs


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc, self.device = [], None, device

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.startswith("Active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
_ONE_THREAD = {"OMP_NUM_THREADS": "1", "OPENBLAS_NUM_THREADS": "1", "MKL_NUM_THREADS": "1", "NUMEXPR_NUM_THREADS": "1"}


def _blas_threads():
    """Threads the BLAS pools of THIS process actually run with (threadpoolctl), for the record."""
    try:
        from threadpoolctl import threadpool_info
        return sorted({int(p.get("num_threads", 0)) for p in threadpool_info()}) or [1]
    except Exception:
        return None


def _cpu_solve(args):
    """One modal solve with the oracle port.  The caller pins the BLAS/OpenMP pools to ONE thread per process through
    the environment BEFORE this process imports NumPy (`run_reference` sets it in the parent of the spawned workers,
    `cpu_baseline_sample` runs a subprocess): SuperLU and ARPACK are single-threaded, a BLAS pool per worker process only
    oversubscribes the host."""
    name, faithful = args
    from oracle import fem_oracle as O
    w, g, mesh = make_case(name)
    t = time.perf_counter()
    modes = O.solve_vectorial_modes(g, mesh, w["n_modes"], faithful_cost=faithful)
    return time.perf_counter() - t, len(modes), _blas_threads()


def cpu_baseline_sample(name: str, n_solves: int = 2):
    """Oracle port timed on one host core: the reference's own work per solve (epsilon re-evaluated in
    each of the 180 form calls like scikit-fem does, SuperLU + ARPACK through SciPy).  Runs in a fresh
    subprocess whose environment pins every BLAS/OpenMP pool to one thread before NumPy is imported."""
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "r = [bench._cpu_solve((%r, True)) for _ in range(%d)]; print(json.dumps(r))" % (ROOT, name, n_solves))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **_ONE_THREAD), capture_output=True, text=True, check=True)
    res = json.loads(out.stdout.strip().splitlines()[-1])
    times = [r[0] for r in res]
    return {"value": 1.0 / statistics.mean(times), "unit": "solves/s", "cores": 1, "kind": "port", "blas_threads": res[0][2],
            "sample": f"{n_solves} full modal solves of {name} with oracle/fem_oracle.py (NumPy restatement of scikit-fem "
                      f"assembly + real scipy eigsh/SuperLU), one process, one thread, {statistics.mean(times):.2f} s each, "
                      f"host has {os.cpu_count()} cores"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    name = args.workload if args.workload in ("cfg1", "cfg2") else "cfg1"
    cores = max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    w, g, mesh = make_case(name)
    os.environ.update(_ONE_THREAD)          # inherited by the spawned workers BEFORE they import NumPy
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        threads = pool.map(_cpu_solve, [(name, True)] * cores)[0][2] if args.warmup == 0 else None
        for _ in range(args.warmup):
            threads = pool.map(_cpu_solve, [(name, True)] * cores)[0][2]
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_solve, [(name, True)] * cores)
        dt = time.perf_counter() - t0
    value = cores * args.steps / dt
    line = {"impl": "reference", "metric": "modal_solves_per_sec", "value": value, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "mesh": {"V": int(mesh.p.shape[1]), "T": int(mesh.t.shape[1])},
                       "step": f"{cores} independent modal solves, one per worker process"},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": "port", "blas_threads_per_process": threads,
                             "sample": f"each step = {cores} concurrent full modal solves, one single-threaded process per host core "
                                       f"({cores} of {os.cpu_count()} cores usable; OMP/OPENBLAS/MKL_NUM_THREADS=1 set before the workers "
                                       "import NumPy; SuperLU/ARPACK are single-threaded)"},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from plfem_b200 import _cabi
    from plfem_b200.batch import ForestPool, default_workers
    from plfem_b200.solver_fem import TrueVectorialMaxwellSolver, sigma_estimate
    from plfem_b200.sweep import gather_records, N_RECORD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL_DEBUG stays as the launcher set it; NCCL's log goes to stderr so that stdout carries only the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")

    # designs per forest / forests in flight: cfg5 is ONE 2M-unknown design per step (a forest of one, one worker)
    B = 1 if args.workload == "cfg5" else max(1, args.inflight)
    # forests in flight: cfg1's contexts take ~4 GB each; the 19-core / LHS forests (k = 52: a 164-vector basis, larger fronts)
    # take 10-15 GB each, twelve of them do not fit one B200 (measured: cfg4 fails to allocate with 12, cfg2 runs with 10)
    NW = 1 if args.workload == "cfg5" else (args.workers if args.workers > 0 else
                                            {"cfg1": default_workers(), "cfg2": 8, "cfg4": 6}.get(args.workload, 6))
    w, jobs = make_jobs(args.workload, B * NW, B)
    forests = [jobs[i * B:(i + 1) * B] for i in range(NW)]          # forest i of a step is worker i's
    g, mesh, n_modes = forests[-1][-1]                              # the largest design (cfg4: sorted) for latency / sizes
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=f"cuda:{local}")

    def sync_all():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize(local)

    # ---- value: a step = ONE forest of B designs per worker; meshes + DOF tables resident, everything else inside ------
    K = args.steps
    pool = ForestPool(device=local, batch=B, workers=NW, want_vectors=False, share_analysis=False)
    resident = {}                                          # worker index -> (problems, materials, shifts, ks) of its forest

    def prepare(_pool, c, wi):
        if wi not in resident:                             # created during warm-up, outside the timed region
            pbs = [_cabi.Problem(j[1], c) for j in forests[wi]]
            mk = [_cabi.material_struct(j[0]) for j in forests[wi]]
            resident[wi] = (pbs, [m for m, _ in mk], [sigma_estimate(j[0]) for j in forests[wi]],
                            [min(j[2] + 12, 2 * pb.n_interior - 4) for j, pb in zip(forests[wi], pbs)], mk)
        return resident[wi]

    def forest_resident(_pool, c, wi, reps=1):
        pbs, mats, sig, ks, _keep = prepare(_pool, c, wi)
        out = []
        for _ in range(reps):
            out.append(_cabi.solve_modes_batch(c, pbs, mats, sig, ks, tol=_cabi.EIG_TOL, maxiter=12000, want_vectors=False,
                                               reuse_symbolic=False))
        return out

    widx = iter(range(NW))
    lock = threading.Lock()

    def my_index():
        with lock:
            return next(widx)
    tl = threading.local()

    def on_worker(p_, c, reps):
        if not hasattr(tl, "wi"):
            tl.wi = my_index()                             # a worker thread keeps its forest (and its resident problems)
        return forest_resident(p_, c, tl.wi, reps)

    for _ in range(args.warmup):                           # every worker thread gets its context and problems
        pool.on_every_worker(lambda p_, c: on_worker(p_, c, 1))
    NF = K * NW
    launches, phase = 0, {n: 0.0 for n in ("ms_symbolic", "ms_symbolic_wall", "ms_assemble", "ms_factor", "ms_lanczos", "ms_metrics", "ms_total")}
    records = np.full((NF * B, N_RECORD), np.nan)
    idle_num = idle_den = 0.0
    flush_l2()
    sync_all()
    with ClockSampler(local) as clocks:
        t0, c0 = time.perf_counter(), time.process_time()
        outs = pool.on_every_worker(lambda p_, c: on_worker(p_, c, K))   # K forests per worker thread, NW threads side by side
        torch.cuda.synchronize(local)
        t_value, cpu_value = time.perf_counter() - t0, time.process_time() - c0
        i = 0
        for per_worker in outs:
            for forest in per_worker:
                st = forest[0][4]
                launches += st.kernel_launches
                for n in phase:
                    phase[n] += getattr(st, n)
                for j, f in enumerate(forest):
                    assert f[5] == 0, "a design failed"
                    records[i * B + j, 0], records[i * B + j, 1], records[i * B + j, 3] = (rank * NF + i) * B + j, 1.0, st.ms_total * 1e-3 / B
                    idle_num += st.batch_block_ops - f[4].n_block_op      # block steps a design was carried after converging
                    idle_den += st.batch_block_ops
                i += 1
        if world > 1:          # the sweep's single collective, inside the timed region
            t0 = time.perf_counter()
            allrec = gather_records(records, world * NF * B, rank, world, local)
            torch.cuda.synchronize(local)
            t_value += time.perf_counter() - t0
            assert allrec.shape == (world * NF * B, N_RECORD)
    stats = st.as_dict()
    host_threads = pool.host_threads
    n_int = {wi: [pb.n_interior for pb in resident[wi][0]] for wi in resident}      # for the byte counts of the end-to-end run
    resident.clear()
    pool.close()                 # the contexts of the value run (streams, device arenas with their front pools) go back before the
    #                              end-to-end pool creates its own: two pools of twelve contexts each do not fit for the larger workloads

    # ---- e2e: public API, host buffers in, mode records (with eigenvectors) out -----------------------------
    # share_analysis off: the designs of this synthetic step sit on ONE mesh, a real sweep's do not — every design pays its own analysis
    pool_e = ForestPool(device=local, batch=B, workers=NW, want_vectors=True, share_analysis=False)
    step_jobs = [j for f in forests for j in f]

    # whole steps through the same pipeline, as ONE stream of forests like the timed region (the same number of forests in
    # flight): contexts, device arenas and the pool of page-locked result blocks reach their steady state before timing
    for modes in pool_e.solve_iter(step_jobs * args.warmup):
        assert not isinstance(modes, Exception), modes
    flush_l2()
    sync_all()
    with ClockSampler(local) as clocks2:
        t0, c0 = time.perf_counter(), time.process_time()
        n_rec = 0
        for modes in pool_e.solve_iter(step_jobs * K):     # records consumed as they arrive (a dataset writer would
            assert not isinstance(modes, Exception), modes     # reduce each to its 86-slot row here)
            n_rec += len(modes) > 0
        torch.cuda.synchronize(local)
        t_e2e, cpu_e2e = time.perf_counter() - t0, time.process_time() - c0
        assert n_rec == B * NF
    pool_e.close()

    # ---- latency: one solve alone through the public API ---------------------------------------------
    lat = []
    lat_stats = {}
    if rank == 0:
        _cabi.load().plfem_set_host_threads(0)
        for i in range(3 + min(args.steps, 10 if args.workload != "cfg5" else 2)):
            flush_l2()
            t0 = time.perf_counter()
            s1 = TrueVectorialMaxwellSolver(g, device=local)
            s1.solve_vectorial_modes(mesh, n_modes)
            lat_stats = dict(s1.last_stats)
            s1.close()
            lat.append(time.perf_counter() - t0)
        lat = lat[3:]
    ctx0 = _cabi.Context.get(local)
    pbs0 = [_cabi.Problem(j[1], ctx0) for j in forests[-1]]
    n_solve = pbs0[-1].n_interior
    k = min(n_modes + 12, 2 * n_solve - 4)
    h2d = sum(j[1].p.nbytes + j[1].t.astype(np.int64).nbytes + 8 * (3 * j[0].n_cores + 4) for j in step_jobs)
    d2h = 0
    for f, wi in zip(forests, range(NW)):
        for j, ni in zip(f, n_int[wi]):
            kk = min(j[2] + 12, 2 * ni - 4)
            d2h += 8 * (kk + kk * 2 * ni + kk * _cabi.NMETRICS)

    if world > 1:
        tt = torch.tensor([t_value, t_e2e], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_value, t_e2e = (float(v) for v in tt.cpu())
        lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(lt)
        launches = int(lt.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline of a forest (CUDA events on the library's stream, L2 flushed per repetition) ----
    mk0 = [_cabi.material_struct(j[0]) for j in forests[-1]]
    fo = _cabi.solve_modes_batch(ctx0, pbs0, [m for m, _ in mk0], [sigma_estimate(j[0]) for j in forests[-1]],
                                 [min(j[2] + 12, 2 * pb.n_interior - 4) for j, pb in zip(forests[-1], pbs0)], want_vectors=False)   # leaves plan + factors on the device
    fstats = fo[-1][4].as_dict()
    # the sweeps are timed in the schedule the timed region ran with (a pool with several workers asks its contexts for one
    # launch per level, a single solve uses the dataflow launch); the other schedule is measured too and reported beside it
    pooled = NW > 1 and not os.environ.get("PLFEM_SWEEP")
    prof_alone = None
    if pooled:
        prof_alone, _ = ctx0.profile_last(repeat=10)
        ctx0.set_sweep_schedule(_cabi.Context.SWEEPS_PER_LEVEL)
    prof, nb_prof = ctx0.profile_last(repeat=10 if args.workload != "cfg5" else 3)
    sweep_schedule = ctx0.sweep_schedule
    if pooled:
        ctx0.set_sweep_schedule(-1)
    nblk = fstats["batch_block_ops"]
    n_sweeps = nblk * (1 + int(fstats["refine_steps"]))        # block-LDL^T solves per operator application: 1 + refinement steps
    fkey, bkey = "forward_sweep_4rhs", "backward_sweep_4rhs"
    share = {k_: 0.0 for k_ in prof}
    share.update({fkey: n_sweeps * prof[fkey][0], bkey: n_sweeps * prof[bkey][0], "factorize": prof["factorize"][0],
                  "assemble": prof["assemble"][0], "spmm_B": nblk * prof["spmm_B"][0],
                  "spmv_K_residual": nblk * int(fstats["refine_steps"]) * prof["spmv_K_residual"][0]})
    # the dominant kernel of the step: the largest time share ("factorize" is a phase of five kernels, reported apart); shares
    # within 2 % of the largest are a tie in this measurement (the two sweep directions differ by less than their run-to-run
    # noise) and go to the candidate that moves more algorithmic bytes
    cand = [k_ for k_ in share if k_ != "factorize"]
    top = max(share[k_] for k_ in cand)
    dom = max((k_ for k_ in cand if share[k_] >= 0.98 * top), key=lambda k_: prof[k_][1])
    kernels = {}
    for name, (ms, nbytes) in prof.items():
        gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else None
        kernels[name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak if gbs else None,
                         "est_ms_per_forest": share[name]}
    # FP64 side of the factorisation: flops of the plan (inverse + W + Schur) over the measured factorisation time
    fl = sum(f[4].factor_flops for f in fo)
    kernels["factorize"]["fp64_tflops"] = fl / (prof["factorize"][0] * 1e-3) / 1e12
    # against the FP64 ceilings measured on this pool's B200 (scripts/micro/fp64_peak.cu, profiles/r02_fp64_peak.txt): DMMA
    # mma.sync m8n8k4 37.0, CUDA-core FMA 34.1, cuBLAS DGEMM 35.1 TFLOP/s.  The phase is inverse + extend-add + two GEMMs + packs
    kernels["factorize"]["fp64_peak_tflops"] = 37.0
    kernels["factorize"]["fp64_peak_source"] = "measured: mma.sync.m8n8k4.f64 chain, profiles/r02_fp64_peak.txt"
    kernels["factorize"]["frac_of_fp64_peak"] = kernels["factorize"]["fp64_tflops"] / 37.0
    kernels["factorize"]["factor_gflop"] = fl / 1e9
    fused = sweep_schedule.startswith("dataflow")
    # launches of one sweep: the TMA-streamed bottom subtrees + ONE dataflow launch for every level above them (or, with
    # PLFEM_SWEEP=levels, one launch per elimination-tree level)
    launches_per = (2 if fused else fstats["n_levels"]) if "sweep" in dom else 1
    ms_dom, bytes_dom = prof[dom]
    traffic = None          # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same forest size)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        tr = tj["sweeps"].get(dom) if tj.get("workload", "cfg1") == args.workload else None
        if tr:      # whole-sweep DRAM bytes of the captured forest, scaled to this forest (traffic is proportional to the
            #         designs) and spread over the same launches `achieved` is quoted per
            traffic = tr["dram_bytes"] * nb_prof / float(tj.get("designs", 12)) / launches_per
    except Exception:
        pass
    above = ("level_forward_kernel<NR> / level_backward_kernel<NR> (every front above them: four-warp tasks, one dataflow launch with tickets and "
             "per-front counters)" if fused else "level_forward_kernel<NR> / level_backward_kernel<NR> (one launch per level above them)")
    roofline = {"kernel": ("stream_forward_kernel<NR> / stream_backward_kernel<NR> (bottom subtrees, one warp each, TMA-streamed) + " + above)
                          if "sweep" in dom else {"factorize": "invert_kernel + gemm_w_dmma_kernel + gemm_schur_dmma_kernel + extend_add_kernel",
                                                  "assemble": "assemble_rows_kernel", "spmm_B": "spmm_b_kernel", "spmv_K_residual": "resid_k_kernel"}[dom],
                "name": dom, "bound": "hbm", "achieved": bytes_dom / (ms_dom * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_dom / (ms_dom * 1e-3) / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "per_launch": {"launches_per_sweep": launches_per, "avg_launch_us": 1e3 * ms_dom / launches_per,
                               "algorithmic_bytes_per_launch": bytes_dom / launches_per, "designs_per_launch": nb_prof},
                "note": "one sweep = one TMA-streamed launch for the bottom subtrees + the launch(es) for the fronts above them, all carrying "
                        "the fronts of every design of the forest; algorithmic bytes = factor entries of the left block columns (8 B each) + the "
                        "right-hand sides; CUDA events on the library's stream around the whole sweep, L2 flushed before each timed sweep"}

    if prof_alone is not None:      # the same sweep as ONE dataflow launch above the subtrees (what a solve alone on the device runs)
        ms_a, by_a = prof_alone[dom]
        roofline["single_solve_schedule"] = {"sweeps": "dataflow launch", "ms": ms_a, "achieved": by_a / (ms_a * 1e-3) / 1e9,
                                             "frac": by_a / (ms_a * 1e-3) / 1e9 / hbm_peak}
        if traffic is not None:
            roofline["traffic_note"] = "ncu --set full capture of the dataflow schedule (profiles/r02_traffic.json), spread over this schedule's launches"
    cpu = cpu_baseline_sample(args.workload, 2) if (world == 1 and args.workload in ("cfg1", "cfg2")) else None
    nst = args.steps
    line = {"metric": "modal_solves_per_sec", "value": world * B * NF / t_value, "unit": "solves/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_value / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "mesh": {"V": int(mesh.p.shape[1]), "T": int(mesh.t.shape[1]), "N_p2": int(pbs0[-1].N),
                                                       "dim": 2 * n_solve, "recipe": "structured cells" if args.workload == "cfg5" else "reference point recipe, refinement 1.0, flat hull triangles dropped",
                                                       "note": "largest design of the step" if args.workload == "cfg4" else "every design of the step"},
                       "lockstep_idle_fraction": idle_num / max(idle_den, 1.0),
                       "step": f"{NW} forests (one per host thread / context) of {B} independent modal solves each = {NW * B} solves; the designs "
                               f"of a forest (each with its own symbolic analysis, assembly, factorisation, eigensolve and reductions) share "
                               f"every kernel launch, the {NW} forests in flight overlap host analysis and device work",
                       "designs_per_forest": B, "forests_in_flight": NW, "solves_per_step": NW * B,
                       "analysis_shared_between_designs": False,
                       "k": k, "lanczos": "thick-restart block Lanczos, 4 vectors per operator application, basis 3k, designs in lockstep", "tol": _cabi.EIG_TOL,
                       "start_vector": "ones (+3 fixed pseudo-random)", "refine_steps": int(fstats["refine_steps"]),
                       "sweeps": sweep_schedule + " above the bottom subtrees (timed region and roofline; a single solve uses the dataflow launch)",
                       "l2": f"inputs larger than L2: one forest streams {B * fstats['factor_entries'] * 8 / 1e6:.0f} MB of factor panels per sweep "
                             f"(front pools {B * fstats['front_pool_doubles'] * 8 / 1e9:.2f} GB); L2 also flushed (512 MiB write) before the timed region",
                       "timing": "wall clock around the K steps (forests) submitted to the worker threads, cuda synchronize + barrier on both sides, max over ranks",
                       "SimulationConfig": {"mesh_min_points": 0, "mesh_target_points": 0}},
            "clocks": clocks.summary(),
            "e2e": {"value": world * B * NF / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / args.steps, "n_modes_returned": len(modes), "clocks": clocks2.summary(),
                    "path": "ForestPool.solve_iter on NumPy meshes: DOF tables + mesh upload, forest solve, eigenvectors + reductions copied back into page-locked result arrays, mode records built and consumed one by one"},
            "gpu_launches": int(launches),
            "latency": {"ms_per_solve_alone_e2e": 1e3 * statistics.mean(lat), "solves_per_s": 1.0 / statistics.mean(lat),
                        "phases_ms": {n: lat_stats[n] for n in ("ms_symbolic", "ms_assemble", "ms_factor", "ms_lanczos", "ms_metrics")}},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "phases_ms_per_forest": {n: v / NF for n, v in phase.items()},
            "host": {"cores": os.cpu_count(), "host_threads_per_forest": host_threads,
                     "cpu_ms_per_solve_value": 1e3 * cpu_value / (B * NF), "cpu_ms_per_solve_e2e": 1e3 * cpu_e2e / (B * NF),
                     "note": "process CPU time of rank 0 inside the two timed regions / designs solved"},
            "solver": {kk: stats[kk] for kk in ("nconv", "n_op", "n_block_op", "n_restart", "n_fronts", "n_levels", "max_front_nodes", "factor_entries",
                                                "front_pool_doubles", "factor_flops", "max_residual", "batch_size", "batch_block_ops", "probe_rho")},
            "kernels": kernels}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS),
                    help="cfg1 (default, the configuration the metric is quoted on), cfg2, cfg4 (heterogeneous LHS sample), cfg5 (2M unknowns)")
    ap.add_argument("--inflight", type=int, default=12, help="designs per forest (= per step)")
    ap.add_argument("--workers", type=int, default=0, help="host threads / contexts, each working on its own forest "
                    "(0 = 12 for cfg1 (batch.default_workers()), 8 for cfg2, 6 for cfg4: device memory)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
