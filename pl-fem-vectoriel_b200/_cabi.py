"""ctypes binding of ``libplfem.so`` (the C ABI declared in ``include/plfem.h``).

There is no CPU implementation behind this module: if the shared library is
missing it is built with nvcc, and if that fails — or no CUDA device is usable
when a context is requested — the error propagates.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libplfem.so"

c_i32, c_i64, c_f64 = C.c_int32, C.c_int64, C.c_double
p_i32, p_i64, p_f64 = C.POINTER(c_i32), C.POINTER(c_i64), C.POINTER(c_f64)

STATUS = {0: "OK", 1: "CUDA", 2: "INVALID", 3: "DEGENERATE", 4: "NOT_READY", 5: "NO_CONVERGENCE",
          6: "SINGULAR", 7: "INTERNAL"}
MATRIX = dict(A=0, B=1, Dxx=2, Dyy=3, Dxy=4, M_inv=5, Kxx=6, Kyy=7, Kxy=8, Kyx=9, M=10, A_int=11, B_int=12)
NMETRICS = 8
#: Residual tolerance handed to the block Lanczos.  The reference passes tol = 1e-7 to ARPACK (`solver_fem.py:197`), which
#: tests convergence only at restarts and in practice returns Ritz vectors converged far beyond it (measured on config 2:
#: the divergence energy v^T D v, which amplifies eigenvector errors through the 1e9-size entries of sliver elements,
#: agrees with a tol = 1e-13 run to 5e-13).  This solver tests every second block step and stops right at the tolerance,
#: so it is asked for one decade more (costs about one block step) to keep such derived quantities within 5e-6 of eigsh.
EIG_TOL = 1e-8

#: every symbol include/plfem.h declares (checked by tests/test_cabi.py)
SYMBOLS = ["plfem_ctx_create", "plfem_ctx_destroy", "plfem_last_error", "plfem_version",
           "plfem_problem_create", "plfem_problem_destroy", "plfem_problem_set_dirichlet", "plfem_problem_info", "plfem_problem_dofs",
           "plfem_quad_points", "plfem_assemble", "plfem_export_csr", "plfem_spmv_csr", "plfem_solve_modes",
           "plfem_plan_sizes", "plfem_plan_export", "plfem_debug_symeig", "plfem_debug_symeig_tail", "plfem_profile_kernels", "plfem_set_host_threads", "plfem_ctx_set_sweep_schedule", "plfem_ctx_sweep_schedule",
           "plfem_debug_solve", "plfem_solve_modes_batch", "plfem_profile_last", "plfem_host_alloc", "plfem_host_free"]


class MeshInfo(C.Structure):
    _fields_ = [(n, c_i64) for n in ("V", "T", "E", "N", "n_boundary", "n_interior", "nnz_scalar", "n_degenerate")]


class Material(C.Structure):
    _fields_ = [("cores_xy", p_f64), ("cores_r", p_f64), ("n_cores", c_i32), ("eps_core", c_f64),
                ("eps_clad", c_f64), ("k0", c_f64), ("alpha_p", c_f64), ("eps_at_quad", p_f64),
                ("scalar_mode", c_i32), ("scalar_shift", c_f64)]


class SolveOpts(C.Structure):
    _fields_ = [("sigma", c_f64), ("k", c_i32), ("ncv", c_i32), ("tol", c_f64), ("maxiter", c_i32),
                ("v0", p_f64), ("leaf_nodes", c_i32), ("max_sn_nodes", c_i32), ("reuse_symbolic", c_i32),
                ("refine", c_i32), ("block", c_i32)]


class SolveStats(C.Structure):
    _fields_ = [("nconv", c_i32), ("n_op", c_i32), ("n_restart", c_i32), ("n_fronts", c_i32),
                ("n_levels", c_i32), ("max_front_nodes", c_i32), ("factor_entries", c_i64),
                ("front_pool_doubles", c_i64), ("factor_flops", c_f64), ("max_residual", c_f64),
                ("ms_symbolic", C.c_float), ("ms_assemble", C.c_float), ("ms_factor", C.c_float),
                ("ms_lanczos", C.c_float), ("ms_metrics", C.c_float), ("ms_total", C.c_float),
                ("kernel_launches", c_i32), ("n_block_op", c_i32), ("batch_size", c_i32), ("batch_block_ops", c_i32),
                ("ms_symbolic_wall", C.c_float), ("refine_steps", c_i32), ("probe_rho", c_f64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class PlfemError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"plfem [{STATUS.get(status, status)}]: {message}")
        self.status = status


_lib = None
_lock = threading.Lock()


def load():
    """Load (building first if needed) the CUDA library.  Never falls back to anything else."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if not LIB_PATH.exists() or _build.needs_build():     # never load a library older than its sources
            try:
                _build.build()
            except RuntimeError:
                if not LIB_PATH.exists():
                    raise                                     # no nvcc and no library: nothing to run
        lib = C.CDLL(str(LIB_PATH))
        vp = C.c_void_p
        lib.plfem_version.restype = C.c_char_p
        lib.plfem_last_error.restype = C.c_char_p
        lib.plfem_last_error.argtypes = [vp]
        lib.plfem_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        lib.plfem_ctx_destroy.argtypes = [vp]
        lib.plfem_ctx_destroy.restype = None
        lib.plfem_problem_create.argtypes = [vp, p_f64, p_i64, c_i64, c_i64, C.POINTER(vp)]
        lib.plfem_problem_destroy.argtypes = [vp]
        lib.plfem_problem_destroy.restype = None
        lib.plfem_problem_set_dirichlet.argtypes = [vp, C.c_int]
        lib.plfem_problem_info.argtypes = [vp, C.POINTER(MeshInfo)]
        lib.plfem_problem_dofs.argtypes = [vp, p_i64, p_f64, p_i64, p_i64]
        lib.plfem_quad_points.argtypes = [vp, p_f64]
        lib.plfem_assemble.argtypes = [vp, C.POINTER(Material)]
        lib.plfem_export_csr.argtypes = [vp, C.c_int, p_i64, p_i64, p_i64, p_i64, p_f64]
        lib.plfem_spmv_csr.argtypes = [vp, c_i64, c_i64, p_i64, p_i64, p_f64, p_f64, p_f64, C.c_int,
                                       C.POINTER(C.c_float)]
        lib.plfem_solve_modes.argtypes = [vp, C.POINTER(Material), C.POINTER(SolveOpts), p_f64, p_f64, p_f64,
                                          p_i32, C.POINTER(SolveStats)]
        lib.plfem_solve_modes_batch.argtypes = [vp, c_i32, C.POINTER(vp), C.POINTER(Material), C.POINTER(SolveOpts),
                                                C.POINTER(p_f64), C.POINTER(p_f64), C.POINTER(p_f64), p_i32,
                                                C.POINTER(SolveStats), p_i32]
        lib.plfem_plan_sizes.argtypes = [vp, c_i32, c_i32, p_i64]
        lib.plfem_plan_export.argtypes = [vp] + [p_i32] * 9 + [p_i64]
        lib.plfem_debug_symeig.argtypes = [c_i32, p_f64, p_f64]
        lib.plfem_debug_symeig_tail.argtypes = [c_i32, p_f64, p_f64, c_i32, p_f64]
        lib.plfem_debug_solve.argtypes = [vp, c_f64, p_f64, p_f64, C.c_int]
        lib.plfem_set_host_threads.argtypes = [C.c_int]
        lib.plfem_set_host_threads.restype = None
        lib.plfem_ctx_set_sweep_schedule.argtypes = [vp, C.c_int]
        lib.plfem_ctx_sweep_schedule.argtypes = [vp]
        lib.plfem_profile_kernels.argtypes = [vp, C.POINTER(Material), c_f64, C.c_int, p_f64, p_f64]
        lib.plfem_profile_last.argtypes = [vp, C.c_int, p_f64, p_f64, p_i32]
        lib.plfem_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
        lib.plfem_host_free.argtypes = [vp]
        lib.plfem_host_free.restype = None
        _lib = lib
        return lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


class PinnedPool:
    """Page-locked host memory for the big result arrays (eigenvectors), carved out of a few large slabs.

    ``empty(shape)`` returns a float64 NumPy array backed by a piece of a pinned slab; when the array (and every view of it —
    the mode records hold views) is garbage-collected the piece goes back to the slab's free list and merges with its free
    neighbours.  Device->host copies into such arrays run at link speed and touch no fresh pages.  Slabs are 512 MiB (or the
    request, if larger): an LHS sweep returns blocks of many sizes (k = 14 ... 52 modes on 20k ... 120k unknowns), and one
    cudaHostAlloc per size class — milliseconds each, serialised with the device's other work — kept happening inside the
    sweep.  Above ``cap_bytes`` of slabs the pool hands out ordinary ``np.empty`` arrays, so a caller that keeps thousands of
    results alive does not pin all of host memory."""

    SLAB = 512 << 20
    ALIGN = 4096

    def __init__(self, cap_bytes: int = 8 << 30):
        self.cap, self.reserved, self.in_use = int(cap_bytes), 0, 0
        self.free: list = []            # [address, size, slab id], sorted by address; neighbours of one slab are merged
        self.n_slabs = 0
        self.lock = threading.Lock()

    def _take(self, size: int):
        for i, (addr, sz, slab) in enumerate(self.free):          # first fit
            if sz >= size:
                if sz == size:
                    del self.free[i]
                else:
                    self.free[i] = [addr + size, sz - size, slab]
                return addr, slab
        return None

    def _give_back(self, addr: int, size: int, slab: int):
        import bisect
        with self.lock:
            self.in_use -= size
            i = bisect.bisect_left(self.free, [addr, 0, 0])
            self.free.insert(i, [addr, size, slab])
            if i + 1 < len(self.free) and self.free[i + 1][2] == slab and addr + size == self.free[i + 1][0]:
                self.free[i][1] += self.free[i + 1][1]
                del self.free[i + 1]
            if i > 0 and self.free[i - 1][2] == slab and self.free[i - 1][0] + self.free[i - 1][1] == addr:
                self.free[i - 1][1] += self.free[i][1]
                del self.free[i]

    def empty(self, shape) -> np.ndarray:
        import bisect
        import weakref
        n = int(np.prod(shape))
        size = -(-8 * max(n, 1) // self.ALIGN) * self.ALIGN
        with self.lock:
            got = self._take(size)
            if got is None:
                slab_bytes = max(self.SLAB, size)
                if self.reserved + slab_bytes > self.cap:
                    return np.empty(shape)
                self.reserved += slab_bytes             # reserved before the (slow) allocation, outside the lock below
        if got is None:
            h = C.c_void_p()
            if load().plfem_host_alloc(slab_bytes, C.byref(h)) != 0 or not h.value:
                with self.lock:
                    self.reserved -= slab_bytes
                return np.empty(shape)
            with self.lock:
                slab = self.n_slabs
                self.n_slabs += 1
                if slab_bytes > size:
                    bisect.insort(self.free, [h.value + size, slab_bytes - size, slab])
                got = (h.value, slab)
        with self.lock:
            self.in_use += size
        addr, slab = got
        buf = (c_f64 * n).from_address(addr)
        weakref.finalize(buf, self._give_back, addr, size, slab)      # runs when the last array / view over buf dies
        return np.ctypeslib.as_array(buf).reshape(shape)


PINNED = PinnedPool()


class Context:
    """One CUDA device + stream (``plfem_ctx``).  Contexts are cached per device."""
    _cache: dict = {}

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        st = self.lib.plfem_ctx_create(int(device), C.byref(h))
        if st != 0:
            raise PlfemError(st, f"cannot create a CUDA context on device {device} "
                                 "(this package has no CPU path)")
        self.handle, self.device = h, int(device)
        import weakref
        self._problems = weakref.WeakSet()      # live problems of this context (closed with it: their buffers are in its arena)

    @classmethod
    def get(cls, device: int = 0) -> "Context":
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    SWEEPS_DATAFLOW, SWEEPS_PER_LEVEL = 0, 1

    def set_sweep_schedule(self, schedule: int):
        """``SWEEPS_DATAFLOW`` (default: one persistent launch per sweep direction, best for a solve alone on the device) or
        ``SWEEPS_PER_LEVEL`` (one launch per elimination-tree level, best with several forests in flight); before the first solve."""
        self.check(self.lib.plfem_ctx_set_sweep_schedule(self.handle, int(schedule)))

    @property
    def sweep_schedule(self) -> str:
        return "one launch per level" if self.lib.plfem_ctx_sweep_schedule(self.handle) == 1 else "dataflow launch"

    def check(self, st: int):
        if st != 0:
            raise PlfemError(st, self.lib.plfem_last_error(self.handle).decode(errors="replace"))

    def profile_last(self, repeat: int = 20):
        """({item: (avg ms, algorithmic bytes)}, batch size) of the device state the last solve on this context left
        (a single design or a forest), measured with CUDA events on the library stream, L2 flushed per repetition."""
        ms = np.zeros(len(Problem.PROFILE_ITEMS)); nbytes = np.zeros(len(Problem.PROFILE_ITEMS)); nb = c_i32()
        self.check(self.lib.plfem_profile_last(self.handle, int(repeat), _ptr(ms, p_f64), _ptr(nbytes, p_f64), C.byref(nb)))
        return {k: (float(a), float(b)) for k, a, b in zip(Problem.PROFILE_ITEMS, ms, nbytes)}, nb.value

    def spmv_csr(self, M, x, repeat: int = 1):
        """y = M @ x on the device for a SciPy CSR matrix; returns (y, avg ms per launch)."""
        M = M.tocsr()
        indptr = np.ascontiguousarray(M.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(M.indices, dtype=np.int64)
        data = np.ascontiguousarray(M.data, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(M.shape[0])
        ms = C.c_float()
        self.check(self.lib.plfem_spmv_csr(self.handle, M.shape[0], M.nnz, _ptr(indptr, p_i64), _ptr(indices, p_i64),
                                           _ptr(data, p_f64), _ptr(x, p_f64), _ptr(y, p_f64), repeat, C.byref(ms)))
        return y, ms.value

    def close(self):
        """Destroy the context: stream, events, pinned staging and every block of its device arena.  Problems created on
        it must be closed first (their buffers live in the arena)."""
        h, self.handle = getattr(self, "handle", None), None
        if h:
            for pb in list(self._problems):
                pb.close()
            self.lib.plfem_ctx_destroy(h)
            for k, v in list(Context._cache.items()):
                if v is self:
                    del Context._cache[k]

    def __del__(self):
        pass  # cached contexts live for the process; pools close the ones they create (ForestPool.close)


def material_struct(geometry, alpha_p: float = 1.0, eps_at_quad=None):
    """(Material, keep-alive tuple) from a geometry duck-type (`geometry_unified.py:200-201`)."""
    pos = np.ascontiguousarray(np.atleast_2d(np.asarray(geometry.positions, dtype=np.float64)))
    rad = np.ascontiguousarray(np.asarray(geometry.core_radii, dtype=np.float64))
    m = Material()
    m.cores_xy, m.cores_r, m.n_cores = _ptr(pos, p_f64), _ptr(rad, p_f64), len(rad)
    m.eps_core, m.eps_clad = float(geometry.n_core ** 2), float(geometry.n_clad ** 2)
    m.k0, m.alpha_p = float(geometry.k0), float(alpha_p)
    keep = [pos, rad]
    if eps_at_quad is not None:
        e = np.ascontiguousarray(eps_at_quad, dtype=np.float64)
        m.eps_at_quad = _ptr(e, p_f64)
        keep.append(e)
    return m, keep


class Problem:
    """One mesh: DOF tables, patterns, front plan and device buffers (``plfem_problem``)."""

    def __init__(self, mesh, ctx: "Context | None" = None, host_only: bool = False):
        self.lib = load()
        self.ctx = None if host_only else (ctx or Context.get(0))
        p = np.ascontiguousarray(mesh.p, dtype=np.float64)
        t = np.ascontiguousarray(mesh.t, dtype=np.int64)
        if p.ndim != 2 or p.shape[0] != 2 or t.ndim != 2 or t.shape[0] != 3:
            raise ValueError("mesh.p must be (2,V) and mesh.t (3,T)")
        h = C.c_void_p()
        st = self.lib.plfem_problem_create(self.ctx.handle if self.ctx else None, _ptr(p, p_f64), _ptr(t, p_i64),
                                           p.shape[1], t.shape[1], C.byref(h))
        if st != 0:
            msg = self.lib.plfem_last_error(self.ctx.handle).decode() if self.ctx else "invalid mesh"
            raise PlfemError(st, msg)
        self.handle = h
        if self.ctx is not None:
            self.ctx._problems.add(self)
        info = MeshInfo()
        self.lib.plfem_problem_info(h, C.byref(info))
        self.info = info
        self.V, self.T, self.N = info.V, info.T, info.N
        self.n_interior = info.n_interior

    def _check(self, st):
        if st != 0:
            if self.ctx is None:
                raise PlfemError(st, "host-only problem")
            self.ctx.check(st)

    def set_dirichlet(self, on: bool):
        """``False``: keep every DOF (natural boundary condition of the reference's scalar solver, `solver_fem.py:245-276`)."""
        self._check(self.lib.plfem_problem_set_dirichlet(self.handle, int(bool(on))))
        self.lib.plfem_problem_info(self.handle, C.byref(self.info))
        self.n_interior = self.info.n_interior

    def dofs(self):
        ed = np.empty((6, self.T), dtype=np.int64)
        loc = np.empty((2, self.N))
        bnd = np.empty(self.info.n_boundary, dtype=np.int64)
        itr = np.empty(self.info.n_interior, dtype=np.int64)
        self._check(self.lib.plfem_problem_dofs(self.handle, _ptr(ed, p_i64), _ptr(loc, p_f64), _ptr(bnd, p_i64),
                                                _ptr(itr, p_i64)))
        return ed, loc, bnd, itr

    def quad_points(self):
        xy = np.empty((2, self.T, 6))
        self._check(self.lib.plfem_quad_points(self.handle, _ptr(xy, p_f64)))
        return xy

    def assemble(self, material: Material):
        self._check(self.lib.plfem_assemble(self.handle, C.byref(material)))

    def export_csr(self, which: str):
        from scipy.sparse import csr_matrix
        wid = MATRIX[which]
        rows, nnz = c_i64(), c_i64()
        self._check(self.lib.plfem_export_csr(self.handle, wid, C.byref(rows), C.byref(nnz), None, None, None))
        indptr = np.empty(rows.value + 1, dtype=np.int64)
        indices = np.empty(nnz.value, dtype=np.int64)
        data = np.empty(nnz.value)
        self._check(self.lib.plfem_export_csr(self.handle, wid, C.byref(rows), C.byref(nnz), _ptr(indptr, p_i64),
                                              _ptr(indices, p_i64), _ptr(data, p_f64)))
        idx_t = np.int32 if max(rows.value, nnz.value) < 2 ** 31 else np.int64
        return csr_matrix((data, indices.astype(idx_t), indptr.astype(idx_t)), shape=(rows.value, rows.value))

    def solve_modes(self, material: Material, sigma: float, k: int, ncv: int = 0, tol: float = EIG_TOL,
                    maxiter: int = 12000, v0=None, want_vectors: bool = True, leaf_nodes: int = 0,
                    max_sn_nodes: int = 0, reuse_symbolic: bool = False, refine: int = 0, block: int = 0):
        n2 = 2 * self.n_interior
        o = SolveOpts(sigma=float(sigma), k=int(k), ncv=int(ncv), tol=float(tol), maxiter=int(maxiter),
                      leaf_nodes=int(leaf_nodes), max_sn_nodes=int(max_sn_nodes),
                      reuse_symbolic=int(bool(reuse_symbolic)), refine=int(refine), block=int(block))
        if v0 is not None:
            v0 = np.ascontiguousarray(v0, dtype=np.float64)
            if v0.shape != (n2,):
                raise ValueError(f"v0 must have length {n2}")
            o.v0 = _ptr(v0, p_f64)
        vals = np.empty(k)
        vecs = PINNED.empty((k, n2)) if want_vectors else None
        met = np.empty((k, NMETRICS))
        ncore = c_i32()
        stats = SolveStats()
        self._check(self.lib.plfem_solve_modes(self.handle, C.byref(material), C.byref(o), _ptr(vals, p_f64),
                                               _ptr(vecs, p_f64), _ptr(met, p_f64), C.byref(ncore), C.byref(stats)))
        return vals, vecs, met, ncore.value, stats

    def debug_solve(self, sigma: float, b, refine: int = 1):
        """x = (A - sigma B)^-1 b with the factors of the last solve (test hook)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        self._check(self.lib.plfem_debug_solve(self.handle, float(sigma), _ptr(b, p_f64), _ptr(x, p_f64), int(refine)))
        return x

    PROFILE_ITEMS = ("assemble", "factorize", "forward_sweep", "backward_sweep", "spmm_B", "spmv_K_residual",
                     "forward_sweep_4rhs", "backward_sweep_4rhs")

    def profile_kernels(self, material: Material, sigma: float, repeat: int = 20) -> dict:
        """{item: (avg ms, algorithmic bytes)} measured with CUDA events on the library stream."""
        ms = np.zeros(len(self.PROFILE_ITEMS))
        nb = np.zeros(len(self.PROFILE_ITEMS))
        self._check(self.lib.plfem_profile_kernels(self.handle, C.byref(material), float(sigma), int(repeat),
                                                   _ptr(ms, p_f64), _ptr(nb, p_f64)))
        return {k: (float(a), float(b)) for k, a, b in zip(self.PROFILE_ITEMS, ms, nb)}

    def plan(self, leaf_nodes: int = 0, max_sn_nodes: int = 0) -> dict:
        """Front plan as NumPy arrays (host logic; needs no GPU)."""
        sz = np.zeros(6, dtype=np.int64)
        self._check(self.lib.plfem_plan_sizes(self.handle, leaf_nodes, max_sn_nodes, _ptr(sz, p_i64)))
        n, nf, nl, ns, nc, _ = (int(v) for v in sz)
        a = dict(perm=np.empty(n, np.int32), first=np.empty(nf, np.int32), s=np.empty(nf, np.int32),
                 parent=np.empty(nf, np.int32), level=np.empty(nf, np.int32), sptr=np.empty(nf + 1, np.int32),
                 strct=np.empty(ns, np.int32), cmap_ptr=np.empty(nf + 1, np.int32), cmap=np.empty(nc, np.int32))
        foff = np.empty(nf + 1, np.int64)
        self._check(self.lib.plfem_plan_export(self.handle, *[_ptr(a[k], p_i32) for k in
                                               ("perm", "first", "s", "parent", "level", "sptr", "strct", "cmap_ptr",
                                                "cmap")], _ptr(foff, p_i64)))
        a.update(foff=foff, n=n, nfronts=nf, nlevels=nl)
        return a

    def close(self):
        if getattr(self, "handle", None):
            self.lib.plfem_problem_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_modes_batch(ctx: "Context", problems, materials, sigmas, ks, tol: float = EIG_TOL, maxiter: int = 12000,
                      want_vectors: bool = True, ncv: int = 0, refine: int = 0, reuse_symbolic: bool = False,
                      leaf_nodes: int = 0, max_sn_nodes: int = 0):
    """Forest solve (``plfem_solve_modes_batch``): the designs share every kernel launch.

    Returns one ``(eigvals, evecs | None, metrics, n_core_dofs, stats, status)`` per design; ``status`` is 0 or the
    ``plfem_status`` of that design alone (the other designs are unaffected)."""
    nb = len(problems)
    if not (nb == len(materials) == len(sigmas) == len(ks)) or nb == 0:
        raise ValueError("problems, materials, sigmas and ks must have the same non-zero length")
    lib = ctx.lib
    mats = (Material * nb)(*materials)
    opts = (SolveOpts * nb)()
    for b in range(nb):
        opts[b] = SolveOpts(sigma=float(sigmas[b]), k=int(ks[b]), ncv=int(ncv), tol=float(tol), maxiter=int(maxiter),
                            leaf_nodes=int(leaf_nodes), max_sn_nodes=int(max_sn_nodes),
                            reuse_symbolic=int(bool(reuse_symbolic)), refine=int(refine), block=0)
    vals = [np.empty(int(k)) for k in ks]
    vecs = [PINNED.empty((int(k), 2 * pb.n_interior)) if want_vectors else None for k, pb in zip(ks, problems)]
    mets = [np.empty((int(k), NMETRICS)) for k in ks]
    ncore = np.zeros(nb, dtype=np.int32)
    status = np.zeros(nb, dtype=np.int32)
    stats = (SolveStats * nb)()
    handles = (C.c_void_p * nb)(*[pb.handle for pb in problems])
    pv = (p_f64 * nb)(*[_ptr(a, p_f64) for a in vals])
    pe = (p_f64 * nb)(*[_ptr(a, p_f64) for a in vecs])
    pm = (p_f64 * nb)(*[_ptr(a, p_f64) for a in mets])
    ctx.check(lib.plfem_solve_modes_batch(ctx.handle, nb, handles, mats, opts, pv, pe, pm, _ptr(ncore, p_i32), stats,
                                          _ptr(status, p_i32)))
    out = []
    for b in range(nb):
        st = SolveStats.from_buffer_copy(stats[b])
        out.append((vals[b], vecs[b], mets[b], int(ncore[b]), st, int(status[b])))
    return out


def symeig(a: np.ndarray):
    """Test hook for the restart eigensolver: returns (w ascending, eigenvectors in columns)."""
    lib = load()
    n = a.shape[0]
    buf = np.asfortranarray(a, dtype=np.float64).copy(order="F")
    w = np.empty(n)
    st = lib.plfem_debug_symeig(n, buf.ctypes.data_as(p_f64), _ptr(w, p_f64))
    if st != 0:
        raise PlfemError(st, "symeig")
    return w, buf


def symeig_tail(a: np.ndarray, p: int):
    """Test hook for the convergence checks' eigensolver: (w ascending, last p rows of the eigenvector matrix, shape (p, n))."""
    lib = load()
    n = a.shape[0]
    buf = np.asfortranarray(a, dtype=np.float64).copy(order="F")
    w = np.empty(n)
    tail = np.empty((p, n), order="F")
    st = lib.plfem_debug_symeig_tail(n, buf.ctypes.data_as(p_f64), _ptr(w, p_f64), p, tail.ctypes.data_as(p_f64))
    if st != 0:
        raise PlfemError(st, "symeig_tail")
    return w, tail
