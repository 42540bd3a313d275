"""Throughput mode: several modal solves in flight on one GPU.

One 7-core cross-section is ~45k unknowns: its kernels are small grids chained level by level, so a
single solve leaves most of a B200 idle and alternates with host-side symbolic analysis.  A
``SolverPool`` keeps ``workers`` host threads, each with its own C-ABI context (own CUDA stream and
device-memory arena); ctypes releases the GIL inside the library, so the symbolic analysis of one
design overlaps the factorisation and Lanczos sweeps of the others, and kernels from different
streams share the SMs.  Results are identical to one-at-a-time solves (every solve is deterministic
and independent).
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Sequence

from . import _cabi
from .solver_fem import TrueVectorialMaxwellSolver


class SolverPool:
    def __init__(self, device: int = 0, workers: int = 4, host_threads_per_solve: int = 0):
        self.device, self.workers = int(device), int(workers)
        lib = _cabi.load()
        cores = os.cpu_count() or 1
        lib.plfem_set_host_threads(host_threads_per_solve or max(1, min(8, cores // max(self.workers, 1))))
        self._local = threading.local()
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="plfem")

    def _ctx(self):
        if not hasattr(self._local, "ctx"):
            self._local.ctx = _cabi.Context(self.device)     # a fresh context: own stream + arena
            # persistent operator kernels of all workers must be co-resident: share the 8 CTA slots per SM
            self._local.ctx.set_coop_ctas(max(1, min(4, 8 // max(self.workers, 1))))
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
        _cabi.load().plfem_set_host_threads(0)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
