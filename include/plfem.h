/* plfem.h — C ABI of the B200-native H-field P2 FEM mode solver.
 *
 * The reference (KhaoulaAguech/pl-fem-vectoriel) has no FFI: its boundary for this path is the
 * Python class `TrueVectorialMaxwellSolver` (solver_fem.py:113-239), which delegates the
 * arithmetic to scikit-fem and SciPy.  This header is what a binding for that class calls
 * instead; each entry point names the reference lines it replaces.  Plain pointers and sizes
 * only, `int` status returns (0 = ok), no exceptions cross the boundary; the message of the last
 * failure is available from plfem_last_error().
 *
 * Conventions
 *   - arrays are C-contiguous like NumPy: p is (2,V) float64, t is (3,T) int64;
 *   - "reference ordering" of an interior vector is [Hx(interior...), Hy(interior...)]
 *     (solver_fem.py:181);
 *   - all output buffers are caller-owned HOST memory unless the name ends in `_dev`.
 *   - a context is bound to one CUDA device and one stream; contexts are independent and may be
 *     driven from different host threads.
 */
#ifndef PLFEM_H
#define PLFEM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plfem_ctx plfem_ctx;
typedef struct plfem_problem plfem_problem;

enum plfem_status {
  PLFEM_OK = 0,
  PLFEM_ERR_CUDA = 1,         /* a CUDA runtime call failed or no usable device */
  PLFEM_ERR_INVALID = 2,      /* bad argument (null pointer, negative size, index out of range) */
  PLFEM_ERR_DEGENERATE = 3,   /* the mesh holds zero-area triangles (scikit-fem would divide by zero) */
  PLFEM_ERR_NOT_READY = 4,    /* call order violated (e.g. export before assemble) */
  PLFEM_ERR_NO_CONVERGENCE = 5, /* eigensolver hit maxiter (scipy: ArpackNoConvergence) */
  PLFEM_ERR_SINGULAR = 6,     /* shifted operator numerically singular (scipy: "Factor is exactly singular") */
  PLFEM_ERR_INTERNAL = 7
};

/* ---- context ---------------------------------------------------------------------------------- */
int plfem_ctx_create(int device, plfem_ctx** out);
void plfem_ctx_destroy(plfem_ctx* ctx);
const char* plfem_last_error(const plfem_ctx* ctx);
/* host threads the symbolic analysis of ONE call (a solve, or a whole forest) may use (0 = default: min(cores, 8) for a
 * single solve, every core for a forest); lower it when several contexts / processes share the host */
void plfem_set_host_threads(int n);
/* How the forward / backward sweeps run above the bottom subtrees: PLFEM_SWEEPS_DATAFLOW (default) = one persistent launch per
 * direction whose tasks wait on per-front counters — fastest for a solve that has the device to itself; PLFEM_SWEEPS_PER_LEVEL =
 * one launch per elimination-tree level — nothing waits on the device, better when several contexts keep forests in flight on
 * one GPU (the forest pool of the host package asks for it).  Results are bit-identical.  -1 restores the default; the
 * environment variable PLFEM_SWEEP=levels|dataflow overrides every context (A/B runs).  Set it before the first solve. */
#define PLFEM_SWEEPS_DATAFLOW 0
#define PLFEM_SWEEPS_PER_LEVEL 1
int plfem_ctx_set_sweep_schedule(plfem_ctx* ctx, int schedule);
/* the schedule in effect for this context (environment override included) */
int plfem_ctx_sweep_schedule(const plfem_ctx* ctx);
/* abi / build identification: "plfem <version> sm_100a" */
const char* plfem_version(void);

/* ---- mesh + DOF tables:  Basis(mesh, ElementTriP2())  (solver_fem.py:126), get_dofs (:179-180) -- */
typedef struct {
  int64_t V, T, E, N;          /* vertices, triangles, facets, scalar P2 DOFs (N = V + E) */
  int64_t n_boundary, n_interior;
  int64_t nnz_scalar;          /* structural nnz of one scalar N x N matrix */
  int64_t n_degenerate;        /* zero-area triangles found */
} plfem_mesh_info;

/* ctx may be NULL: a host-only problem (DOF tables, front plan) that needs no GPU; assembly and
 * solves on it return PLFEM_ERR_INVALID */
int plfem_problem_create(plfem_ctx* ctx, const double* p, const int64_t* t, int64_t V, int64_t T,
                         plfem_problem** out);
void plfem_problem_destroy(plfem_problem* pb);
/* on = 1 (default): the solve eliminates the boundary DOFs (solver_fem.py:179-184).  on = 0: every DOF is kept, the
 * natural boundary condition of ScalarHelmholtzSolver (solver_fem.py:245-276), which never calls get_dofs() */
int plfem_problem_set_dirichlet(plfem_problem* pb, int on);
int plfem_problem_info(const plfem_problem* pb, plfem_mesh_info* info);
/* element_dofs (6,T) int64, doflocs (2,N) float64, boundary (n_boundary) int64, interior (n_interior)
 * int64; any pointer may be NULL to skip it */
int plfem_problem_dofs(const plfem_problem* pb, int64_t* element_dofs, double* doflocs,
                       int64_t* boundary, int64_t* interior);

/* ---- material:  geometry.epsilon(x, y)  (geometry_unified.py:325-347) ------------------------- */
typedef struct {
  const double* cores_xy;      /* (n_cores, 2) core centres [um] */
  const double* cores_r;       /* (n_cores)    core radii   [um] */
  int32_t n_cores;
  double eps_core, eps_clad;   /* Re eps = n_core^2, n_clad^2 (evaluated by the caller) */
  double k0;                   /* 2 pi / lambda [1/um] */
  double alpha_p;              /* divergence penalty, reference uses 1.0 (solver_fem.py:158) */
  const double* eps_at_quad;   /* optional (T,6) Re eps sampled by the caller at plfem_quad_points();
                                  when non-NULL it overrides the disc model */
  int32_t scalar_mode;         /* 0: vectorial H-field system (solver_fem.py:122-169).  1: the scalar Helmholtz pencil of
                                  ScalarHelmholtzSolver (solver_fem.py:245-276), (K - k0^2 M_eps, M), carried in the Hx block;
                                  the Hy block holds (scalar_shift * M, M), whose eigenvalues all equal scalar_shift */
  double scalar_shift;         /* scalar mode only: pick it far from sigma so that the Hy copy is never wanted */
} plfem_material;

/* global coordinates of the 6 quadrature points of every element, out_xy is (2,T,6) float64 */
int plfem_quad_points(const plfem_problem* pb, double* out_xy);

/* ---- assembly:  9 x asm(form, basis) + block build  (solver_fem.py:131-167) -------------------- */
enum plfem_matrix {
  PLFEM_MAT_A = 0,      /* 2N x 2N  [[Kxx+aDxx-k0^2 M, Kxy+aDxy],[Kyx+aDxy^T, Kyy+aDyy-k0^2 M]] */
  PLFEM_MAT_B = 1,      /* 2N x 2N  blockdiag(M_inv, M_inv) */
  PLFEM_MAT_DXX = 2, PLFEM_MAT_DYY = 3, PLFEM_MAT_DXY = 4, PLFEM_MAT_MINV = 5,  /* N x N, as returned */
  PLFEM_MAT_KXX = 6, PLFEM_MAT_KYY = 7, PLFEM_MAT_KXY = 8, PLFEM_MAT_KYX = 9, PLFEM_MAT_M = 10,
  PLFEM_MAT_A_INT = 11, /* A[idx,:][:,idx], idx = [interior, interior+N]  (solver_fem.py:181-182) */
  PLFEM_MAT_B_INT = 12
};
/* assemble all scalar matrices on the device over the full N x N pattern */
int plfem_assemble(plfem_problem* pb, const plfem_material* mat);
/* CSR export with SciPy's structure rules (element-level zeros dropped by asm, exact-zero sums
 * dropped by sparse +/-, sorted indices).  Call with data == NULL to get nnz first.
 * indptr has rows+1 entries (int64), indices nnz (int64), data nnz (float64). */
int plfem_export_csr(plfem_problem* pb, int which, int64_t* rows, int64_t* nnz, int64_t* indptr,
                     int64_t* indices, double* data);

/* ---- CSR SpMV (scipy csr_matvec; solver_fem.py:214 and the M-product inside eigsh) ------------- */
/* y = M x on the device for an exported matrix; x, y are HOST vectors of length rows.  `repeat`
 * launches are timed with CUDA events; avg milliseconds per launch is returned in *ms. */
int plfem_spmv_csr(plfem_ctx* ctx, int64_t rows, int64_t nnz, const int64_t* indptr, const int64_t* indices,
                   const double* data, const double* x, double* y, int repeat, float* ms);

/* ---- modal solve:  Dirichlet elimination + eigsh + per-mode reductions (solver_fem.py:179-225) - */
typedef struct {
  double sigma;        /* shift (solver_fem.py:187-193, computed by the caller) */
  int32_t k;           /* eigenpairs wanted = min(n_modes+12, 2 N_solve - 4)  (solver_fem.py:196) */
  int32_t ncv;         /* Lanczos basis size; 0 = default: scipy's max(2k+1, 20) for block = 1, 3k for block Lanczos */
  double tol;          /* 1e-7 in the reference */
  int32_t maxiter;     /* restarts allowed, 12000 in the reference */
  const double* v0;    /* optional start vector, reference ordering, length 2 N_solve; NULL = ones */
  int32_t leaf_nodes;  /* nested-dissection leaf size (0 = default) */
  int32_t max_sn_nodes;/* supernode width limit in nodes (0 = default) */
  int32_t reuse_symbolic; /* 1 = the ordering / front plan may be reused: the one of the previous solve on this problem, or, inside
                             a forest, the one of another design with reuse_symbolic = 1 on an identical mesh (the bands of a
                             wavelength sweep share their mesh); 0 = analyse this design on its own */
  int32_t refine;      /* iterative-refinement steps per operator application: 0 = chosen by probing one raw solve
                          (1 for the reference's meshes), n > 0 = n, -1 = none.  Environment PLFEM_RELAX_AT=x (default
                          0 = off) applies one step fewer once every wanted Ritz pair is within x of convergence */
  int32_t block;       /* Lanczos block size: 0 = default (4 vectors per operator application), 1 = single vector */
} plfem_solve_opts;

typedef struct {
  int32_t nconv, n_op, n_restart;   /* converged pairs, operator applications, restarts */
  int32_t n_fronts, n_levels, max_front_nodes;
  int64_t factor_entries;           /* doubles read by one forward+backward sweep / 2 */
  int64_t front_pool_doubles;
  double factor_flops;
  double max_residual;              /* normwise backward error: max_i ||A x - lambda B x||_2 / ((||A||_F + |lambda| ||B||_F) ||x||_2) */
  float ms_symbolic, ms_assemble, ms_factor, ms_lanczos, ms_metrics, ms_total; /* host wall / CUDA events */
  int32_t kernel_launches;
  int32_t n_block_op;               /* sequential operator applications (= n_op / block size) */
  /* forest (plfem_solve_modes_batch): ms_assemble .. ms_metrics, ms_total and kernel_launches are those of the WHOLE
   * batch (the designs share every launch); ms_symbolic is this design's own host analysis */
  int32_t batch_size;               /* designs solved together (1 for plfem_solve_modes) */
  int32_t batch_block_ops;          /* lockstep operator applications of the batch */
  float ms_symbolic_wall;           /* host wall time of the analysis of all designs (parallel threads) */
  int32_t refine_steps;             /* refinement steps per operator application actually used */
  double probe_rho;                 /* |dx|/|x| of the first refinement correction of a raw block-LDL^T solve (this design) */
} plfem_solve_stats;

/* Per-mode reductions of solver_fem.py:212-220, computed on the l2-normalised (vx, vy):
 * metrics is (k, 8): [div_energy, sum_e_core, sum_e, Px_core, Py_core, Px_all, Py_all, norm2_raw] */
#define PLFEM_NMETRICS 8
int plfem_solve_modes(plfem_problem* pb, const plfem_material* mat, const plfem_solve_opts* opts,
                      double* eigvals,    /* (k) ascending, = beta^2 */
                      double* evecs,      /* (k, 2 N_solve) reference ordering, l2-normalised; may be NULL */
                      double* metrics,    /* (k, PLFEM_NMETRICS) */
                      int32_t* core_dof_count, /* number of interior DOFs inside a core (solver_fem.py:200-203) */
                      plfem_solve_stats* stats);

/* ---- forest of designs: the sweep's production mode (dataset generation, README.md:226-243) -----------------
 * nb independent (mesh, material, shift) designs are solved as ONE block-diagonal problem: their fronts share the
 * level-batched factorisation and sweep launches and their block-Lanczos iterations advance in lockstep, so a batch
 * costs about as many (latency-bound) launches as a single design.  Results per design are identical to
 * plfem_solve_modes on that design alone.  All problems must have been created on `ctx`.  Array arguments have nb
 * entries; evecs[b] may be NULL.  Returns PLFEM_OK when the batch ran; statuses[b] tells whether design b succeeded
 * (a singular shift or non-convergence of one design does not affect the others), plfem_last_error() describes the
 * first failed design.  opts[0].refine/.block and the largest ncv / maxiter apply to the whole batch. */
int plfem_solve_modes_batch(plfem_ctx* ctx, int32_t nb, plfem_problem* const* pbs, const plfem_material* mats,
                            const plfem_solve_opts* opts, double* const* eigvals, double* const* evecs,
                            double* const* metrics, int32_t* core_dof_counts, plfem_solve_stats* stats,
                            int32_t* statuses);

/* ---- page-locked host memory for result buffers --------------------------------------------------------------
 * The eigenvectors of a forest are >100 MB; copying them into pageable memory runs at a few GB/s (driver staging
 * copies + first-touch page faults).  Buffers obtained here are page-locked (cudaHostAlloc, portable), so the
 * device->host copies of plfem_solve_modes / plfem_solve_modes_batch run at link speed.  Free with plfem_host_free. */
int plfem_host_alloc(size_t bytes, void** out);
void plfem_host_free(void* p);

/* ---- measurement hook for bench.py: per-kernel device times (CUDA events on the library's stream, L2
 * flushed before each repetition) and the algorithmic bytes of the same items; needs a prior solve.
 * index: 0 assembly (K1), 1 front load + factorisation, 2 forward sweep, 3 backward sweep,
 *        4 B product (SpMM, 2 right-hand sides), 5 K residual SpMV, 6 / 7 forward / backward sweep with 4 right-hand sides */
#define PLFEM_NPROFILE 8
int plfem_profile_kernels(plfem_problem* pb, const plfem_material* mat, double sigma, int repeat,
                          double* out_ms, double* out_bytes);
/* the same on whatever the last solve of this context left on the device — a single design or a forest (every
 * launch then carries all designs; *batch_size tells how many).  The problems of that solve must still exist. */
int plfem_profile_last(plfem_ctx* ctx, int repeat, double* out_ms, double* out_bytes, int32_t* batch_size);

/* ---- debug / test hooks (host logic checks that need no GPU) ----------------------------------- */
/* sizes: [n, nfronts, nlevels, strct_len, cmap_len, nchild] */
int plfem_plan_sizes(plfem_problem* pb, int32_t leaf_nodes, int32_t max_sn_nodes, int64_t sizes[6]);
int plfem_plan_export(plfem_problem* pb, int32_t* perm, int32_t* first, int32_t* s, int32_t* parent,
                      int32_t* level, int32_t* sptr, int32_t* strct, int32_t* cmap_ptr, int32_t* cmap,
                      int64_t* foff);

/* x = (A - sigma B)^-1 b with the factors of the last plfem_solve_modes (same sigma), reference ordering */
int plfem_debug_solve(plfem_problem* pb, double sigma, const double* b, double* x, int refine);
/* dense symmetric eigensolver used at Lanczos restarts: a is n*n column-major, overwritten by eigenvectors */
int plfem_debug_symeig(int32_t n, double* a, double* w);
/* the convergence checks' variant: eigenvalues (ascending) and only the last p rows of the eigenvector matrix, tail[j*p + r] = Z(n-p+r, j) */
int plfem_debug_symeig_tail(int32_t n, const double* a, double* w, int32_t p, double* tail);

#ifdef __cplusplus
}
#endif
#endif /* PLFEM_H */
