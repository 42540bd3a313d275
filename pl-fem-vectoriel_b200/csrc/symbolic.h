// Host-side symbolic analysis for the H-field P2 mode solver.
//
// Replaces, for this path, what the reference gets from
//   * scikit-fem  Basis(mesh, ElementTriP2())            (solver_fem.py:126)  -> DofTables
//   * scikit-fem  basis.get_dofs().all() / setdiff1d     (solver_fem.py:179-180)
//   * SuperLU's COLAMD ordering + symbolic factorisation inside scipy eigsh
//     (solver_fem.py:197)                                                     -> FrontPlan
// Nothing numerical happens here: integers and coordinates only.
#pragma once
#include <cstdint>
#include <vector>

namespace plfem {

struct DofTables {
  int64_t V = 0, T = 0, E = 0, N = 0;
  std::vector<int32_t> edofs;      // [6*T]  element-major: edofs[6*e + k]
  std::vector<int32_t> facets;     // [2*E]  facets[2*f + {0,1}] = (min vertex, max vertex), lexicographic
  std::vector<int32_t> t2f;        // [3*T]  element-major
  std::vector<double> doflocs;     // [2*N]  x[0..N), y[0..N)
  std::vector<int32_t> boundary;   // sorted boundary DOFs
  std::vector<int32_t> interior;   // sorted interior DOFs (= setdiff1d(arange(N), boundary))
  std::vector<int32_t> n2e_ptr;    // [N+1] node -> elements (ascending element id)
  std::vector<int32_t> n2e;        // [6*T]
  int64_t n_degenerate = 0;        // elements with det J == 0
};

// Build DOF tables from p (2,V) row-major and t (3,T) row-major (numpy C order).
void build_dof_tables(const double* p, const int64_t* t, int64_t V, int64_t T, DofTables& out);

// Sparsity pattern of a scalar P2 matrix restricted/renumbered by newid (old node -> new id, -1 = dropped).
struct Pattern {
  int32_t n = 0;
  std::vector<int32_t> rowptr;      // [n+1]
  std::vector<int32_t> col;         // sorted within each row (new numbering)
  std::vector<int32_t> old_of_new;  // [n]
  std::vector<int32_t> new_of_old;  // [N]
};
void build_pattern(const DofTables& d, const std::vector<int32_t>& new_of_old, int32_t n_new, Pattern& out);

// worker threads the symbolic phase may use per call (0 = default: min(cores, 8) or $PLFEM_HOST_THREADS)
void set_host_threads(int n);
// the same for the calling thread only (0 = follow the global setting)
void set_host_threads_local(int n);
// threads one call (a solve, or the analysis of a whole forest) may occupy: the plfem_set_host_threads value,
// $PLFEM_HOST_THREADS, or every hardware thread
int host_thread_budget();

struct SymbolicOptions {
  int leaf_nodes = 24;   // stop dissecting below this many nodes
  int max_sn_nodes = 64; // split separators into chains of supernodes of at most this many nodes
  int search_min_nodes = 512; // subsets at least this large try 4 cut directions, smaller ones 1
};

// Multifrontal plan on the interior nodes, in nested-dissection order.
// Every front f owns the contiguous new ids [first[f], first[f]+s[f]) and has an update set
// strct[sptr[f]..sptr[f+1]) of later ids (sorted).  Fronts are numbered in post-order.
struct FrontPlan {
  int32_t n = 0;                    // interior nodes
  int32_t nfronts = 0;
  std::vector<int32_t> perm;        // [n] new id -> interior index (position in DofTables::interior)
  std::vector<int32_t> first, s, parent, level;
  std::vector<int32_t> sptr, strct; // update sets
  std::vector<int32_t> sn_of;       // [n] new id -> front
  std::vector<int64_t> foff;        // [nfronts+1] offset (in doubles) of each front's (2nf x 2nf) matrix
  // children
  std::vector<int32_t> cptr, child; // [nfronts+1], children lists
  std::vector<int32_t> cmap_ptr;    // [nfronts+1] offset into cmap for front f's own struct (as a child)
  std::vector<int32_t> cmap;        // position of each strct entry of f inside parent's index list
  // level schedule (level 0 = fronts without children)
  int32_t nlevels = 0;
  std::vector<int32_t> lptr, lfront; // fronts grouped by level
  // statistics
  int64_t factor_entries = 0;       // doubles kept for the solve phase: sum (2s)^2 + (2s)(2u)
  double factor_flops = 0.0;        // inverse + W + Schur
  int32_t max_front = 0, max_s = 0; // in nodes
};

// dof: the mesh's DOF tables (the dissection runs on its P1 vertex graph and is lifted to the P2 nodes; node
//      neighbourhoods are read from the element tables); x, y: coordinates of the interior nodes (interior index order)
// adj: only when front_plan_needs_adjacency(dof) (the P2-graph dissection, PLFEM_DISSECT_P2=1): pattern of the interior
//      nodes in *interior index* numbering (identity order over DofTables::interior); otherwise may be NULL
bool front_plan_needs_adjacency(const DofTables& dof);
void build_front_plan(const DofTables& dof, const Pattern* adj, const double* x, const double* y, const SymbolicOptions& opt,
                      FrontPlan& out);

// Forest of several independent designs as ONE plan / ONE block-diagonal pattern: node ids, front ids and all
// offsets of design b are shifted behind those of designs 0..b-1; level l of the result is the union of the
// designs' levels l.  node_off / front_off get nb+1 entries.  `perm` of the result keeps each design's LOCAL
// interior indices, `old_of_new` of the merged pattern each design's LOCAL DOF ids.
void merge_front_plans(const std::vector<const FrontPlan*>& parts, FrontPlan& out, std::vector<int32_t>& node_off,
                       std::vector<int32_t>& front_off);

}  // namespace plfem
