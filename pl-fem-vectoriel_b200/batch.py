"""Throughput mode: many modal solves per GPU (the dataset sweep of `README.md:226-243`).

One 7-core cross-section is ~45k unknowns: its kernels are small grids chained level by level
(~2000 dependent launches per solve), so a single solve leaves most of a B200 idle.  Two ways to
fill it:

``ForestPool`` (production mode)
    Designs are grouped into *forests* of ``batch`` designs.  A forest is solved by ONE C-ABI call
    (``plfem_solve_modes_batch``) as one block-diagonal problem: the elimination trees of its designs
    share the level-batched factorisation and sweep launches and their block-Lanczos iterations
    advance in lockstep, so a forest costs about as many launches as one design.  ``workers``
    host threads (default 2), each with its own context (CUDA stream + device arena), work on
    different forests, so the host-side symbolic analysis of one forest overlaps the device work of
    another.  Results are identical to one-at-a-time solves.

``SolverPool``
    One design per host thread and stream (no lockstep): kept for latency-oriented callers that
    want each result as soon as it is ready.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Sequence

from . import _cabi
from .solver_fem import TrueVectorialMaxwellSolver, modes_from_solution, sigma_estimate


def default_workers(cores: int = 0) -> int:
    """Forests in flight per GPU.  A worker thread analyses its forest on the host (about 4 ms per config-1 design), drives it
    on the device and builds its records; with the sweeps launched level by level (nothing waits on the device) more forests
    in flight keep filling the GPU until about twelve.  Measured on one B200, config 1, solves/s device-resident / end to end:
    16 host cores: 6 workers 480 / 416, 8: 517 / 492, 10: 525 / 511, 12: 529 / 526, 15: 530 / 529;
    restricted to 4 cores (a rank's share of a 32-core 8-GPU node): 9 workers 417-431 / 372-390, 12: 454 / 390.
    Each context keeps its device arena (about 4 GB for forests of twelve 7-core designs, 10-15 GB for 19-core designs with
    k = 52): callers with large designs pass fewer workers; an arena that cannot allocate trims every arena's cache first."""
    return 12


def usable_cores() -> int:
    """Host cores this process may count on: its share of the node under torchrun (one rank per GPU), never more than its
    affinity mask allows."""
    n = os.cpu_count() or 1
    if hasattr(os, "sched_getaffinity"):
        n = min(n, len(os.sched_getaffinity(0)))
    return max(1, n // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))


class ForestPool:
    """``solve_many(jobs)`` with jobs = ``(geometry, mesh, n_modes_target)`` -> list of mode lists."""

    def __init__(self, device: int = 0, batch: int = 8, workers: int = 2, want_vectors: bool = True, host_threads: int = 0,
                 share_analysis: bool = True):
        self.device, self.batch, self.workers = int(device), max(1, int(batch)), max(1, int(workers))
        self.want_vectors = want_vectors
        #: designs of a forest that sit on an identical mesh (the bands of a wavelength sweep) share ONE ordering / front plan
        self.share_analysis = bool(share_analysis)
        lib = _cabi.load()
        # host threads per forest: the cores this process may count on (its share of the node under torchrun), spread
        # over the forests in flight with some oversubscription — hundreds of runnable threads per core cost more than
        # the idle cores they could fill
        cores = usable_cores()
        # (a forest of ONE design — the 2M-unknown stress mesh — runs its own dissection on all of them: halves on separate threads)
        share = -(-3 * cores // (2 * self.workers))
        self.host_threads = host_threads or max(1, min(self.batch if self.batch > 1 else cores // self.workers, share))
        lib.plfem_set_host_threads(self.host_threads)
        self._local = threading.local()
        self._contexts: list = []                     # every context a worker thread created (destroyed in close())
        self._ctx_lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="plfem-forest")
        self._build = ThreadPoolExecutor(max_workers=max(1, min(self.batch, cores)), thread_name_prefix="plfem-dof")
        self.last_stats: List[dict] = []

    def _ctx(self):
        if not hasattr(self._local, "ctx"):
            self._local.ctx = _cabi.Context(self.device)     # a fresh context: own stream + arena
            if self.workers > 1:
                # several forests in flight on this GPU: sweeps as one launch per level — a dataflow launch keeps CTAs waiting
                # on the device, which costs the other forests' kernels their slots (measured 502 vs 456 solves/s)
                self._local.ctx.set_sweep_schedule(_cabi.Context.SWEEPS_PER_LEVEL)
            with self._ctx_lock:
                self._contexts.append(self._local.ctx)
        return self._local.ctx

    # -- one forest ---------------------------------------------------------------------------------
    def solve_forest(self, jobs: Sequence[tuple], problems=None, return_raw: bool = False):
        """Solve ``jobs`` as ONE forest on the calling thread's context.  ``problems`` (optional) are resident
        ``_cabi.Problem`` objects of that context for the jobs' meshes; otherwise they are created (DOF tables on
        host threads, mesh upload) and released here."""
        ctx = self._ctx()
        own = problems is None
        if own:
            problems = list(self._build.map(lambda j: _cabi.Problem(j[1], ctx), jobs))
        try:
            # A design whose host-side preparation fails (its epsilon callable raises, k out of range ...) is reported alone
            # and left out of the forest; the others are solved as usual (the C layer does the same for device-side
            # failures through statuses[b]).
            out: list = [None] * len(jobs)
            stats: list = [None] * len(jobs)
            live, mats, keep, sigmas, ks = [], [], [], [], []
            for j, ((geo, mesh, n_modes), pb) in enumerate(zip(jobs, problems)):
                try:
                    m, k_ = TrueVectorialMaxwellSolver(geo, device=self.device, ctx=ctx)._material(pb)
                    sg, kk = sigma_estimate(geo), min(n_modes + 12, 2 * pb.n_interior - 4)
                    if kk < 1:
                        raise ValueError("mesh too small for the requested number of modes")
                except Exception as e:                      # noqa: BLE001
                    out[j] = e
                    continue
                live.append(j); mats.append(m); keep.append(k_); sigmas.append(sg); ks.append(kk)
            if live:
                res = _cabi.solve_modes_batch(ctx, [problems[j] for j in live], mats, sigmas, ks, tol=_cabi.EIG_TOL,
                                              maxiter=12000, want_vectors=self.want_vectors, reuse_symbolic=self.share_analysis)
                for j, (vals, vecs, met, ncore, st, status) in zip(live, res):
                    geo, pb = jobs[j][0], problems[j]
                    stats[j] = st.as_dict()
                    if status != 0:
                        out[j] = _cabi.PlfemError(status, "design failed inside a forest")
                        continue
                    try:      # e.g. no eigenvalue inside the n_eff window: the reference raises here too (`solver_fem.py:228`)
                        guided, raw, frac = modes_from_solution(geo, pb.n_interior, vals, vecs, met, ncore)
                    except Exception as e:                  # noqa: BLE001
                        out[j] = e
                        continue
                    out[j] = (guided, dict(beta_sq=vals, evecs=vecs, metrics=met, modes_raw=raw, frac_core=frac,
                                           stats=stats[j])) if return_raw else guided
            self.last_stats = stats
            self._local.last_stats = stats          # per worker thread: `last_stats` is whichever forest finished last
            return out
        finally:
            if own:
                for pb in problems:
                    pb.close()

    # -- many forests, pipelined over the worker threads -------------------------------------------------
    def solve_many(self, jobs: Sequence[tuple]) -> List[list]:
        """Mode lists in job order.  A design that fails inside its forest raises here (after all forests ran)."""
        chunks = [jobs[i:i + self.batch] for i in range(0, len(jobs), self.batch)]
        out: List[list] = []
        for part in self._pool.map(self.solve_forest, chunks):
            out.extend(part)
        for r in out:
            if isinstance(r, Exception):
                raise r
        return out

    def solve_iter(self, jobs: Sequence[tuple]):
        """Like ``solve_many`` but yields the mode list (or the Exception) of one job at a time, in job order, while
        later forests are still being solved — a consumer that reduces each result (a dataset record) and drops it
        keeps only a few forests of eigenvectors alive, and their page-locked buffers are recycled.  Forests are handed to
        the worker threads as the consumer takes results: at most ``workers + max(4, workers / 2)`` are submitted and not yet
        consumed — every worker busy plus some finished forests waiting behind a slower one that is due first (results come in
        job order; a window of ``workers + 2`` was measured to starve six workers: 441 -> 385 solves/s) — so finished forests
        cannot pile up without bound and the page-locked pool stays below its cap."""
        from collections import deque
        chunks = [jobs[i:i + self.batch] for i in range(0, len(jobs), self.batch)]
        pending: deque = deque()
        nxt = 0
        while nxt < len(chunks) or pending:
            while nxt < len(chunks) and len(pending) < self.workers + max(4, self.workers // 2):
                pending.append(self._pool.submit(self.solve_forest, chunks[nxt]))
                nxt += 1
            part = pending.popleft().result()
            while part:
                yield part.pop(0)

    def map_forests(self, fn: Callable, items: Sequence):
        """Run ``fn(pool, ctx, item)`` on the worker threads (bench hook: resident problems per context)."""
        return list(self._pool.map(lambda it: fn(self, self._ctx(), it), items))

    def on_every_worker(self, fn: Callable):
        """Run ``fn(pool, ctx)`` once on EACH worker thread (warm-up: contexts, resident problems)."""
        gate = threading.Barrier(self.workers)

        def one(_):
            gate.wait()                                   # every task is held until all threads have taken one
            return fn(self, self._ctx())
        return list(self._pool.map(one, range(self.workers)))

    def close(self):
        self._pool.shutdown(wait=True)
        self._build.shutdown(wait=True)
        with self._ctx_lock:
            ctxs, self._contexts = self._contexts, []
        for c in ctxs:                                 # stream, events and the device arena (GBs per forest) go back
            c.close()
        _cabi.load().plfem_set_host_threads(0)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class SolverPool:
    def __init__(self, device: int = 0, workers: int = 4, host_threads_per_solve: int = 0):
        self.device, self.workers = int(device), int(workers)
        lib = _cabi.load()
        cores = usable_cores()
        lib.plfem_set_host_threads(host_threads_per_solve or max(1, min(8, cores // max(self.workers, 1))))
        self._local = threading.local()
        self._contexts: list = []
        self._ctx_lock = threading.Lock()
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="plfem")

    def _ctx(self):
        if not hasattr(self._local, "ctx"):
            self._local.ctx = _cabi.Context(self.device)     # a fresh context: own stream + arena
            with self._ctx_lock:
                self._contexts.append(self._local.ctx)
        return self._local.ctx

    def _solve(self, job):
        geometry, mesh, n_modes = job
        s = TrueVectorialMaxwellSolver(geometry, device=self.device, ctx=self._ctx())
        try:
            return s.solve_vectorial_modes(mesh, n_modes)
        finally:
            s.close()

    def solve_many(self, jobs: Sequence[tuple]) -> List[list]:
        """jobs: (geometry, mesh, n_modes_target) triples -> list of mode lists, in order."""
        return list(self._pool.map(self._solve, jobs))

    def map(self, fn: Callable, items: Sequence):
        """Run ``fn(ctx, item)`` on the pool's threads with each thread's own context."""
        return list(self._pool.map(lambda it: fn(self._ctx(), it), items))

    def close(self):
        self._pool.shutdown(wait=True)
        with self._ctx_lock:
            ctxs, self._contexts = self._contexts, []
        for c in ctxs:
            c.close()
        _cabi.load().plfem_set_host_threads(0)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
