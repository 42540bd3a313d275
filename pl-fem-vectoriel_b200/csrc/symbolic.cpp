// Host-side symbolic analysis: DOF numbering, sparsity pattern, nested dissection, front plan.
// See symbolic.h for what each piece replaces in the reference.
#include "symbolic.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>
#include <atomic>
#include <cstring>
#include <thread>
#include <type_traits>
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace plfem {

// ------------------------------------------------------------------------------------------------
// DOF tables (scikit-fem ElementTriP2 numbering)
// ------------------------------------------------------------------------------------------------
void build_dof_tables(const double* p, const int64_t* t, int64_t V, int64_t T, DofTables& d) {
  d.V = V; d.T = T;
  const double* px = p;
  const double* py = p + V;
  const int64_t* t0 = t; const int64_t* t1 = t + T; const int64_t* t2 = t + 2 * T;
  for (int64_t e = 0; e < T; ++e)
    for (const int64_t* tr : {t0, t1, t2})
      if (tr[e] < 0 || tr[e] >= V) throw std::runtime_error("mesh.t holds a vertex index outside [0, V)");

  // facets: local edges (0,1),(1,2),(0,2); bucket by min vertex, sort each bucket by max vertex
  static const int LE[3][2] = {{0, 1}, {1, 2}, {0, 2}};
  std::vector<int32_t> emin(3 * T), emax(3 * T);
  std::vector<int32_t> cnt(V + 1, 0);
  for (int64_t e = 0; e < T; ++e) {
    const int64_t v[3] = {t0[e], t1[e], t2[e]};
    for (int k = 0; k < 3; ++k) {
      int64_t a = v[LE[k][0]], b = v[LE[k][1]];
      if (a > b) std::swap(a, b);
      emin[3 * e + k] = (int32_t)a; emax[3 * e + k] = (int32_t)b;
      cnt[a + 1]++;
    }
  }
  for (int64_t v = 0; v < V; ++v) cnt[v + 1] += cnt[v];
  // half-edges grouped by min vertex as (max vertex, half-edge id) in one 64-bit key; a bucket (~6 entries) is sorted by
  // counting, for every key, the smaller ones (no data-dependent branch: comparison sorts of such buckets spend their
  // time in mispredictions)
  std::vector<int64_t> slot(3 * T);
  {
    std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
    for (int64_t h = 0; h < 3 * T; ++h) slot[fill[emin[h]]++] = ((int64_t)emax[h] << 32) | (int64_t)h;
  }
  d.t2f.assign(3 * T, -1);
  d.facets.resize(6 * T);                     // trimmed to 2 E below
  std::vector<int32_t> fcount(3 * T + 1);
  int32_t* fac = d.facets.data();
  int32_t* fc = fcount.data();
  int32_t* t2f = d.t2f.data();
  int64_t nf = 0;
  for (int64_t v = 0; v < V; ++v) {
    int32_t b = cnt[v], e = cnt[v + 1];
    const int32_t m = e - b;
    if (m > 32) std::sort(slot.begin() + b, slot.begin() + e);
    else if (m > 1) {
      int64_t key[32];
      for (int32_t i = 0; i < m; ++i) key[i] = slot[b + i];
      for (int32_t i = 0; i < m; ++i) {
        int32_t r = 0;
        for (int32_t j = 0; j < m; ++j) r += key[j] < key[i];
        slot[b + r] = key[i];
      }
    }
    int32_t last = -1;
    for (int32_t i = b; i < e; ++i) {
      const int32_t h = (int32_t)(slot[i] & 0xffffffff), hmax = (int32_t)(slot[i] >> 32);
      // a new facet starts where the max vertex changes (stores without a branch: a repeated facet rewrites its own entry)
      const int32_t isnew = hmax != last;
      last = hmax;
      nf += isnew;
      const int64_t f = nf - 1;
      fac[2 * f] = (int32_t)v; fac[2 * f + 1] = hmax;
      fc[f] = (isnew ? 0 : fc[f]) + 1;
      // slot h = 3*e + k  ->  element-major t2f
      t2f[h] = (int32_t)f;
    }
  }
  d.facets.resize(2 * nf); fcount.resize(nf);
  d.E = nf;
  d.N = V + d.E;

  d.edofs.resize(6 * T);
  for (int64_t e = 0; e < T; ++e) {
    d.edofs[6 * e + 0] = (int32_t)t0[e]; d.edofs[6 * e + 1] = (int32_t)t1[e]; d.edofs[6 * e + 2] = (int32_t)t2[e];
    for (int k = 0; k < 3; ++k) d.edofs[6 * e + 3 + k] = (int32_t)(V + d.t2f[3 * e + k]);
  }

  // boundary DOFs: both vertices and the facet DOF of every facet owned by exactly one element
  std::vector<uint8_t> isb(d.N, 0);
  for (int64_t f = 0; f < d.E; ++f)
    if (fcount[f] == 1) { isb[d.facets[2 * f]] = 1; isb[d.facets[2 * f + 1]] = 1; isb[V + f] = 1; }
  d.boundary.clear(); d.interior.clear();
  for (int64_t i = 0; i < d.N; ++i) (isb[i] ? d.boundary : d.interior).push_back((int32_t)i);

  // DOF locations: affine image of the reference nodes; loop order (local node outer, element
  // inner) makes later writes win exactly like the fancy-index assignment it restates.
  static const double RX[6] = {0.0, 1.0, 0.0, 0.5, 0.5, 0.0};
  static const double RY[6] = {0.0, 0.0, 1.0, 0.0, 0.5, 0.5};
  d.doflocs.assign(2 * d.N, 0.0);
  d.n_degenerate = 0;
  for (int k = 0; k < 6; ++k)
    for (int64_t e = 0; e < T; ++e) {
      const double a00 = px[t1[e]] - px[t0[e]], a01 = px[t2[e]] - px[t0[e]];
      const double a10 = py[t1[e]] - py[t0[e]], a11 = py[t2[e]] - py[t0[e]];
      const int32_t g = d.edofs[6 * e + k];
      d.doflocs[g] = a00 * RX[k] + a01 * RY[k] + px[t0[e]];
      d.doflocs[d.N + g] = a10 * RX[k] + a11 * RY[k] + py[t0[e]];
      if (k == 0 && a00 * a11 - a01 * a10 == 0.0) d.n_degenerate++;
    }

  // node -> elements (ascending element id)
  d.n2e_ptr.assign(d.N + 1, 0);
  for (int64_t i = 0; i < 6 * T; ++i) d.n2e_ptr[d.edofs[i] + 1]++;
  for (int64_t i = 0; i < d.N; ++i) d.n2e_ptr[i + 1] += d.n2e_ptr[i];
  d.n2e.resize(6 * T);
  {
    std::vector<int32_t> fill(d.n2e_ptr.begin(), d.n2e_ptr.end() - 1);
    for (int64_t e = 0; e < T; ++e)
      for (int k = 0; k < 6; ++k) d.n2e[fill[d.edofs[6 * e + k]]++] = (int32_t)e;
  }
}

namespace {
std::atomic<int> g_host_threads{0};
thread_local int tl_host_threads = 0;   // per-thread override: a forest analyses its designs on separate threads

int host_threads() {
  if (tl_host_threads > 0) return tl_host_threads;
  const int set = g_host_threads.load(std::memory_order_relaxed);
  if (set > 0) return set;
  static const int n = [] {
    if (const char* e = std::getenv("PLFEM_HOST_THREADS")) return std::max(1, atoi(e));
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(hc ? hc : 1u, 8u));
  }();
  return n;
}

template <class F>
void parallel_for(int64_t n, int64_t grain, F&& fn) {
  const int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n / std::max<int64_t>(grain, 1)));
  if (nt <= 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const int64_t chunk = (n + nt - 1) / nt;
  for (int t = 1; t < nt; ++t) th.emplace_back([&, t] { fn(std::min(n, t * chunk), std::min(n, (t + 1) * chunk)); });
  fn(0, std::min(n, chunk));
  for (auto& x : th) x.join();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Sparsity pattern in an arbitrary renumbering
// ------------------------------------------------------------------------------------------------
void set_host_threads(int n) { g_host_threads.store(std::max(0, n)); }
void set_host_threads_local(int n) { tl_host_threads = std::max(0, n); }
int host_thread_budget() {
  const int set = g_host_threads.load(std::memory_order_relaxed);
  if (set > 0) return set;
  if (const char* e = std::getenv("PLFEM_HOST_THREADS")) return std::max(1, atoi(e));
  return (int)std::max(1u, std::thread::hardware_concurrency());
}

void build_pattern(const DofTables& d, const std::vector<int32_t>& new_of_old, int32_t n_new, Pattern& out) {
  out.n = n_new;
  out.new_of_old = new_of_old;
  out.old_of_new.assign(n_new, -1);
  for (int64_t o = 0; o < d.N; ++o)
    if (new_of_old[o] >= 0) out.old_of_new[new_of_old[o]] = (int32_t)o;
  out.rowptr.assign(n_new + 1, 0);
  // rows are independent: each worker fills a private column buffer for its row range
  const int nt = host_threads();
  std::vector<std::vector<int32_t>> bufs(nt);
  std::vector<int64_t> lo(nt + 1, n_new);
  const int64_t chunk = ((int64_t)n_new + nt - 1) / nt;
  for (int t = 0; t <= nt; ++t) lo[t] = std::min<int64_t>(n_new, t * chunk);
  auto work = [&](int t) {
    std::vector<int32_t>& col = bufs[t];
    col.reserve((size_t)(lo[t + 1] - lo[t]) * 12);
    std::vector<int32_t> stamp(n_new, -1);
    for (int32_t r = (int32_t)lo[t]; r < (int32_t)lo[t + 1]; ++r) {
      const int32_t o = out.old_of_new[r];
      const size_t b0 = col.size();
      for (int32_t q = d.n2e_ptr[o]; q < d.n2e_ptr[o + 1]; ++q) {
        const int32_t* ed = &d.edofs[6 * (int64_t)d.n2e[q]];
        for (int k = 0; k < 6; ++k) {
          const int32_t c = new_of_old[ed[k]];
          if (c >= 0 && stamp[c] != r) { stamp[c] = r; col.push_back(c); }
        }
      }
      std::sort(col.begin() + b0, col.end());
      out.rowptr[r + 1] = (int32_t)(col.size() - b0);
    }
  };
  if (nt > 1 && n_new > 4096) {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  } else {
    lo.assign(nt + 1, n_new); lo[0] = 0;
    work(0);
  }
  for (int32_t r = 0; r < n_new; ++r) out.rowptr[r + 1] += out.rowptr[r];
  out.col.resize(out.rowptr[n_new]);
  for (int t = 0; t < nt; ++t)
    if (!bufs[t].empty()) std::copy(bufs[t].begin(), bufs[t].end(), out.col.begin() + out.rowptr[lo[t]]);
}

// ------------------------------------------------------------------------------------------------
// Nested dissection + front plan
// ------------------------------------------------------------------------------------------------
namespace {

// The same order by six 11-bit passes over the whole 64-bit keys (all histograms from one read, uniform digits skipped): for
// large inputs, see below.
void argsort_doubles_wide(const std::vector<double>& key, std::vector<int32_t>& order) {
  const int32_t n = (int32_t)key.size();
  constexpr int B = 11, R = 1 << B, NP = 6;                 // 6 x 11 = 66 bits >= 64
  std::vector<uint64_t> k(n), k2(n);
  std::vector<int32_t> a(n), b(n);
  std::vector<int32_t> cnt((size_t)NP * (R + 1), 0);
  for (int32_t i = 0; i < n; ++i) {
    uint64_t u; std::memcpy(&u, &key[i], 8);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    k[i] = u; a[i] = i;
    for (int pass = 0; pass < NP; ++pass) cnt[(size_t)pass * (R + 1) + ((u >> (B * pass)) & (R - 1)) + 1]++;
  }
  for (int pass = 0; pass < NP; ++pass) {
    int32_t* c = cnt.data() + (size_t)pass * (R + 1);
    const int sh = B * pass;
    bool single = false;
    for (int j = 1; j <= R; ++j) if (c[j] == n) { single = true; break; }
    if (single) continue;
    for (int j = 0; j < R; ++j) c[j + 1] += c[j];
    for (int32_t i = 0; i < n; ++i) { const int32_t d = c[(k[i] >> sh) & (R - 1)]++; k2[d] = k[i]; b[d] = a[i]; }
    k.swap(k2); a.swap(b);
  }
  order.swap(a);
}

// order[r] = index of the r-th smallest key; stable (ties keep ascending index).  LSD radix (three 11-bit digits, the three
// histograms from one read) on the upper 32 bits of the order-preserving integer image of the doubles, then one insertion
// pass over the full keys: entries that agree in their upper 32 bits (relative distance < 2^-20: exact ties of a symmetric
// mesh, hardly anything else) arrive in index order and leave sorted by (key, index) — the order a stable sort of the full
// keys gives.
// Large inputs take the plain 64-bit passes: on a structured grid nearly every entry sits in a run of equal upper halves, and the
// insertion pass then reads the full keys through the index — two cache misses per entry once they have outgrown the cache.
void argsort_doubles(const std::vector<double>& key, std::vector<int32_t>& order) {
  const int32_t n = (int32_t)key.size();
  if (n > 100000) { argsort_doubles_wide(key, order); return; }
  std::vector<uint64_t> k(n);
  std::vector<uint32_t> hk(n), hk2(n);
  std::vector<int32_t> a(n), b(n);
  constexpr int B = 11, R = 1 << B;
  std::vector<int32_t> cnt(3 * (R + 1), 0);
  int32_t* c0 = cnt.data(); int32_t* c1 = c0 + R + 1; int32_t* c2 = c1 + R + 1;
  for (int32_t i = 0; i < n; ++i) {
    uint64_t u; std::memcpy(&u, &key[i], 8);
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    k[i] = u;
    const uint32_t h = (uint32_t)(u >> 32);
    hk[i] = h; a[i] = i;
    c0[(h & (R - 1)) + 1]++; c1[((h >> B) & (R - 1)) + 1]++; c2[(h >> (2 * B)) + 1]++;
  }
  for (int pass = 0; pass < 3; ++pass) {
    int32_t* c = cnt.data() + pass * (R + 1);
    const int sh = B * pass;
    const int nb = pass == 2 ? (1 << (32 - 2 * B)) : R;
    bool single = false;
    for (int j = 1; j <= nb; ++j) if (c[j] == n) { single = true; break; }
    if (single) continue;
    for (int j = 0; j < nb; ++j) c[j + 1] += c[j];
    for (int32_t i = 0; i < n; ++i) { const int32_t d = c[(hk[i] >> sh) & (R - 1)]++; hk2[d] = hk[i]; b[d] = a[i]; }
    hk.swap(hk2); a.swap(b);
  }
  for (int32_t i = 1; i < n; ++i) {
    if (hk[i] != hk[i - 1]) continue;
    const int32_t v = a[i];
    const uint64_t kv = k[v];
    int32_t j = i;
    while (j > 0 && hk[j - 1] == hk[i] && (k[a[j - 1]] > kv || (k[a[j - 1]] == kv && a[j - 1] > v))) { a[j] = a[j - 1]; --j; }
    a[j] = v;
  }
  order.swap(a);
}

// A forest of tree nodes in flat storage: the own list and the child list of a node are slices of two pools — the analysis
// creates thousands of small nodes per design, and an allocation per list was a fifth of its time (more with a dozen analyses
// side by side in one allocator).
struct Forest {
  struct Node {
    int32_t own_b, own_n;   // slice of `own`: interior indices, elimination order inside the node
    int32_t ch_b, ch_n;     // slice of `child`: tree node ids (within the same Forest)
    int od;                 // separators: projection direction along which the own list is ordered (-1: leaf)
  };
  std::vector<Node> nodes;
  std::vector<int32_t> own, child;
  // (o and c must not point into this forest's pools)
  int32_t add(const int32_t* o, int32_t no, const int32_t* c, int32_t nc, int od) {
    nodes.push_back(Node{(int32_t)own.size(), no, (int32_t)child.size(), nc, od});
    own.insert(own.end(), o, o + no);
    child.insert(child.end(), c, c + nc);
    return (int32_t)nodes.size() - 1;
  }
  const int32_t* own_of(int32_t t) const { return own.data() + nodes[t].own_b; }
  const int32_t* children_of(int32_t t) const { return child.data() + nodes[t].ch_b; }
  // appends the nodes of `o`; returns the offset its node ids got
  int32_t absorb(const Forest& o) {
    const int32_t off = (int32_t)nodes.size(), oo = (int32_t)own.size(), co = (int32_t)child.size();
    for (Node nd : o.nodes) { nd.own_b += oo; nd.ch_b += co; nodes.push_back(nd); }
    own.insert(own.end(), o.own.begin(), o.own.end());
    for (int32_t c : o.child) child.push_back(c + off);
    return off;
  }
};

struct Dissector {
  static constexpr int ND = 4;  // projection directions: x, y, x+y, x-y
  const Pattern& adj;
  const double* x;
  const double* y;
  const SymbolicOptions& opt;
  // Per node: its position in each direction list (lists[d][pos4[v].r[d]] == v).  A subset under examination occupies the
  // same range [off, off + n) of every list, so "w belongs to the subset" is a range test on w's position — the partition
  // below moves the nodes AND writes their new positions, which keeps the positions current for every node that is still
  // to be dissected (nodes of other subsets lie in other ranges; a separator gets INT32_MAX for good).
  struct alignas(16) Rk { int32_t r[4]; };
  std::vector<Rk> pos4;
  struct alignas(16) Ext { int32_t hi[4], lo[4]; };
  std::vector<Ext> ext;        // per node: extreme positions among its neighbours inside the subset (and itself)
  std::vector<int32_t> lists[ND];  // the node set of the current call, sorted along each direction
  bool split_chains = true;        // cut separators into chains of <= max_sn_nodes supernodes here (false: the caller does)
  struct Scratch { std::vector<int32_t> tmp, hist, sep, mm; std::vector<double> fac, cost; };   // sep: a stack over the recursion

  Dissector(const Pattern& a, const double* x_, const double* y_, const SymbolicOptions& o)
      : adj(a), x(x_), y(y_), opt(o), pos4(a.n, Rk{{0, 0, 0, 0}}), ext(a.n) {
    const int32_t n = a.n;
    auto one = [&](int d) {
      std::vector<double> key(n);
      for (int32_t v = 0; v < n; ++v) key[v] = proj(d, v);
      argsort_doubles(key, lists[d]);
    };
    if (host_threads() >= ND && n > 4096) {
      std::vector<std::thread> th;
      for (int d = 1; d < ND; ++d) th.emplace_back(one, d);
      one(0);
      for (auto& t : th) t.join();
    } else {
      for (int d = 0; d < ND; ++d) one(d);
    }
    // (on one thread: the four positions of a node share a cache line — threads filling a lane each would fight over it)
    for (int d = 0; d < ND; ++d) {
      const int32_t* Ld = lists[d].data();
      for (int32_t r = 0; r < n; ++r) pos4[Ld[r]].r[d] = r;
    }
  }

  double proj(int d, int32_t v) const {
    switch (d) { case 0: return x[v]; case 1: return y[v]; case 2: return x[v] + y[v]; default: return x[v] - y[v]; }
  }

  // Dissect the node set stored (in ND different orders) at lists[d][off .. off+n).
  // Appends tree nodes to `F` and the ids of the nodes heading the resulting sub-forest to `heads` (which doubles as the
  // stack of the recursion: the heads the two halves append are the children of this call's separator).
  void dissect(int32_t off, int32_t n, Forest& F, std::vector<int32_t>& heads, Scratch& sc, int par_budget) {
    if (n == 0) return;
    int32_t* L0 = lists[0].data() + off;
    if (n <= opt.leaf_nodes) {
      heads.push_back(F.add(L0, n, nullptr, 0, -1));
      return;
    }
    // Candidate cuts: ND directions x every split position in the middle 40%.  Along a direction a
    // node of rank r is in the left-boundary separator of split h iff r < h <= (highest neighbour
    // rank), so the separator sizes for EVERY h come from the histograms of the extreme neighbour ranks:
    // left boundary of h = h - #{hi < h}, right boundary = #{lo < h} - h.  Pick the
    // (direction, h, side) with the smallest separator, mildly penalising imbalance.
    const int ndir = (n >= opt.search_min_nodes) ? ND : 2;
    double best_cost = 1e300; int best_dir = 0; int32_t best_h = n / 2; bool best_left = true;
    const int32_t h0 = std::max<int32_t>(1, (int32_t)(0.3 * n)), h1 = std::min<int32_t>(n - 1, (int32_t)(0.7 * n));
    struct Cand { double cost = 1e300; int32_t h = 0; bool left = true; };
    Cand cand[ND];
    {
      sc.hist.assign((size_t)2 * ndir * n, 0);      // per direction: n counters of the highest, n of the lowest neighbour positions
      // ONE pass over the adjacency: the extreme neighbour positions along all directions at once.
      // (small subsets examine two directions only: most of the nodes sit in such subsets, and the pass over their
      // neighbours is the hot loop of the whole analysis)
      // Branch-free over the neighbours: the four positions of a node are one 128-bit lane set, a neighbour outside the subset
      // contributes 0 to the maxima (positions are >= 0 and a node's own takes part) and INT32_MAX to the minima.  The range
      // test runs per lane ((p - off) <u n as a signed compare of biased values): a node outside the subset is outside the
      // range in EVERY list this subset looks at (lanes 2, 3 of a node stop being updated below its last four-direction
      // ancestor and stay inside that ancestor's half — disjoint from any subset that still reads them).  The
      // extremes are kept per node (ext[v]): the separator of the chosen cut is read off them below without a second pass
      // over the adjacency.
      auto neighbour_pass = [&](auto nd_tag) {
        constexpr int NDIR = decltype(nd_tag)::value;
        const __m128i big = _mm_set1_epi32(INT32_MAX);
        const __m128i bias = _mm_set1_epi32((int32_t)(0x80000000u - (uint32_t)off)), nb = _mm_set1_epi32((int32_t)(0x80000000u + (uint32_t)n));
        const int32_t* rp = adj.rowptr.data(); const int32_t* cl = adj.col.data();
        const Rk* ps = pos4.data();
        int32_t* nh[ND]; int32_t* nl_[ND];
        for (int d = 0; d < NDIR; ++d) { nh[d] = sc.hist.data() + (size_t)(2 * d) * n - off; nl_[d] = sc.hist.data() + (size_t)(2 * d + 1) * n - off; }
        for (int32_t i = 0; i < n; ++i) {
          const int32_t v = L0[i];
          const __m128i rv = _mm_load_si128((const __m128i*)&ps[v]);
          __m128i hi = rv, lo = rv;
          for (int32_t q = rp[v], qe = rp[v + 1]; q < qe; ++q) {
            const __m128i rw = _mm_load_si128((const __m128i*)&ps[cl[q]]);
            const __m128i in = _mm_cmpgt_epi32(nb, _mm_add_epi32(rw, bias));
            hi = _mm_max_epi32(hi, _mm_and_si128(rw, in));
            lo = _mm_min_epi32(lo, _mm_blendv_epi8(big, rw, in));
          }
          _mm_store_si128((__m128i*)&ext[v].hi, hi); _mm_store_si128((__m128i*)&ext[v].lo, lo);
          const int32_t* h4 = ext[v].hi; const int32_t* l4 = ext[v].lo;
          for (int d = 0; d < NDIR; ++d) { nh[d][h4[d]]++; nl_[d][l4[d]]++; }
        }
      };
      if (ndir == ND) neighbour_pass(std::integral_constant<int, ND>{}); else neighbour_pass(std::integral_constant<int, 2>{});
      // the imbalance factor of a split position is the same along every direction: once per call (a loop of divisions the
      // compiler vectorises; same IEEE operations as one by one)
      sc.fac.resize(h1 - h0 + 1);
      {
        double* fac = sc.fac.data();
        const double dn = n;
        for (int32_t h = h0; h <= h1; ++h) fac[h - h0] = 1.0 + 1.5 * std::fabs(2.0 * h / dn - 1.0);
      }
      // Per direction: the boundary sizes of every split position (two running sums), their costs four at a time, then the
      // FIRST position attaining the smallest cost — what comparing them one by one with "<" selects, without a
      // data-dependent branch per position (the sizes hover around their minimum: such a branch mispredicts all along).
      const int32_t nh_ = h1 - h0 + 1;
      sc.mm.resize(nh_ + 4); sc.cost.resize(nh_ + 4);
      for (int d = 0; d < ndir; ++d) {
        const int32_t* nh = sc.hist.data() + (size_t)(2 * d) * n; const int32_t* nlo = nh + n;
        const double* fac = sc.fac.data();
        int32_t* mm = sc.mm.data();            // min(cl, cr) + 1, sign bit of the word below = "left boundary is the smaller one"
        double* cost = sc.cost.data();
        int32_t chi = 0, clo = 0;
        for (int32_t h = 1; h < h0; ++h) { chi += nh[h - 1]; clo += nlo[h - 1]; }
        for (int32_t h = h0; h <= h1; ++h) {
          chi += nh[h - 1]; clo += nlo[h - 1];
          const int32_t cl = h - chi, cr = clo - h;       // members of the left / right boundary of split h
          mm[h - h0] = cl <= cr ? cl + 1 : -(cr + 1);     // (both >= 0: the sign carries the side)
        }
        __m256d vmin = _mm256_set1_pd(1e300);
        int32_t i = 0;
        for (; i + 4 <= nh_; i += 4) {
          const __m128i m4 = _mm_abs_epi32(_mm_loadu_si128((const __m128i*)(mm + i)));
          const __m256d c4 = _mm256_mul_pd(_mm256_cvtepi32_pd(m4), _mm256_loadu_pd(fac + i));
          _mm256_storeu_pd(cost + i, c4);
          vmin = _mm256_min_pd(vmin, c4);
        }
        alignas(32) double lane[4];
        _mm256_store_pd(lane, vmin);
        double best = std::min(std::min(lane[0], lane[1]), std::min(lane[2], lane[3]));
        for (; i < nh_; ++i) { cost[i] = (double)std::abs(mm[i]) * fac[i]; best = std::min(best, cost[i]); }
        int32_t at = 0;
        while (at < nh_ - 1 && cost[at] != best) ++at;
        Cand c;
        c.cost = best; c.h = h0 + at; c.left = mm[at] > 0;
        cand[d] = c;
      }
    }
    for (int d = 0; d < ndir; ++d)
      if (cand[d].cost < best_cost) { best_cost = cand[d].cost; best_dir = d; best_h = cand[d].h; best_left = cand[d].left; }
    const int32_t h = best_h;
    const int32_t* Lb = lists[best_dir].data() + off;
    const size_t sep_mark = sc.sep.size();
    if (best_left) {          // left half: members with a neighbour of rank >= h
      for (int32_t i = 0; i < h; ++i) { const int32_t v = Lb[i]; if (ext[v].hi[best_dir] >= off + h) sc.sep.push_back(v); }
    } else {                  // right half: members with a neighbour of rank < h
      for (int32_t i = h; i < n; ++i) { const int32_t v = Lb[i]; if (ext[v].lo[best_dir] < off + h) sc.sep.push_back(v); }
    }
    // stable three-way partition of every direction list: [left | right | (separator: dropped)], the new positions written
    // as the nodes move.  The side of a node is its position along the chosen direction, so that list goes last.
    const int32_t nsep = (int32_t)(sc.sep.size() - sep_mark);
    const int32_t nl = best_left ? h - nsep : h, nr = n - nsep - nl;
    for (int32_t q = 0; q < nsep; ++q) pos4[sc.sep[sep_mark + q]] = Rk{{INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX}};
    sc.tmp.resize(n);
    for (int dd = 0; dd < ndir; ++dd) {         // a subset that examines two directions has only smaller subsets below it
      const int d = dd == ndir - 1 ? best_dir : (dd < best_dir ? dd : dd + 1);
      int32_t* Ld = lists[d].data() + off;
      int32_t* tmp = sc.tmp.data();
      const int32_t cut = off + h;
      int32_t a = 0, b = 0;
      for (int32_t i = 0; i < n; ++i) {
        const int32_t v = Ld[i];
        int32_t* pv = pos4[v].r;
        const int32_t pb = pv[best_dir];
        if (pb == INT32_MAX) continue;
        if (pb < cut) { Ld[a] = v; pv[d] = off + a; ++a; } else { tmp[b] = v; pv[d] = off + nl + b; ++b; }
      }
      std::copy(tmp, tmp + b, Ld + a);
    }
    const size_t kid_mark = heads.size();
    if (par_budget > 1 && std::min(nl, nr) >= 1024) {
      // the two halves are independent: left half on a new thread with its own forest and scratch.  (Each side reads the
      // positions of neighbours that belong to the other side while that side rewrites them: whatever it reads lies in the
      // other half's range of the lists or is the separator mark — "not a member" either way.)
      Forest FL; std::vector<int32_t> hl;
      std::thread th([&] { Scratch s2; dissect(off, nl, FL, hl, s2, par_budget / 2); });
      dissect(off + nl, nr, F, heads, sc, par_budget - par_budget / 2);
      th.join();
      const int32_t shift = F.absorb(FL);
      for (int32_t& hd : hl) hd += shift;
      heads.insert(heads.begin() + kid_mark, hl.begin(), hl.end());      // left heads first
    } else {
      dissect(off, nl, F, heads, sc, 1);
      dissect(off + nl, nr, F, heads, sc, 1);
    }
    if (nsep == 0) return;  // disconnected halves: no front of its own, the children's heads go up
    // order the separator along the cut so that chain links are spatially compact
    // (a caller that splits the chains itself sorts the lifted separators: nothing to order here)
    const int od = (best_dir == 0) ? 1 : (best_dir == 1 ? 0 : (best_dir == 2 ? 3 : 2));
    int32_t* sep = sc.sep.data() + sep_mark;
    if (split_chains)
      std::sort(sep, sep + nsep, [&](int32_t a, int32_t b) {
        const double pa = proj(od, a), pb = proj(od, b);
        return pa < pb || (pa == pb && a < b);
      });
    const int32_t ns = nsep;
    const int32_t nchunks = split_chains ? (ns + opt.max_sn_nodes - 1) / opt.max_sn_nodes : 1;
    int32_t prev = -1, pos = 0;
    for (int32_t k = 0; k < nchunks; ++k) {
      const int32_t len = ns / nchunks + (k < ns % nchunks ? 1 : 0);
      prev = k == 0 ? F.add(sep + pos, len, heads.data() + kid_mark, (int32_t)(heads.size() - kid_mark), od)
                    : F.add(sep + pos, len, &prev, 1, od);
      pos += len;
    }
    sc.sep.resize(sep_mark);
    heads.resize(kid_mark);
    heads.push_back(prev);
  }
};

}  // namespace

namespace {

// Nested dissection of the P1 VERTEX graph, lifted to the P2 nodes.  The vertex graph has 4x fewer nodes and half the
// degree of the P2 node graph, so the dissection costs ~1/7; a vertex separator S lifts to the P2 separator
// S + {edge nodes with both ends in S}: an edge node (a, b) goes to the tree node of its DEEPER endpoint (the endpoints of
// a mesh edge are adjacent, so their tree nodes lie on one root path), which keeps every pair of P2 nodes that share an
// element on a common root path — the tree stays a valid elimination tree of the P2 graph.  Edge nodes without an
// interior endpoint (chords between boundary vertices) are eliminated last, in a tree node above all roots.
void dissect_vertex_graph(const DofTables& d, const std::vector<int32_t>& interior, const double* x, const double* y,
                          const SymbolicOptions& opt, Forest& out, std::vector<int32_t>& out_roots) {
  const int64_t V = d.V, N = d.N;
  const int32_t n = (int32_t)interior.size();    // interior P2 nodes; interior[i] = DOF id of interior index i
  static const bool timing = std::getenv("PLFEM_TIMING") != nullptr;
  auto clk = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = clk();
  std::vector<int32_t> int_of(N, -1);
  for (int32_t i = 0; i < n; ++i) int_of[interior[i]] = i;
  // interior vertices and their graph (two vertices are adjacent iff they share an element)
  std::vector<int32_t> vid(V, -1), vdof;
  for (int64_t v = 0; v < V; ++v) if (int_of[v] >= 0) { vid[v] = (int32_t)vdof.size(); vdof.push_back((int32_t)v); }
  const int32_t nv = (int32_t)vdof.size();
  // two vertices share an element iff they share a mesh edge: the adjacency comes straight from the facet table (two
  // counting passes over the edges; rows are unsorted and carry no self entry — the dissection only asks for the extreme
  // neighbour ranks, and a node's own rank always takes part)
  Pattern vadj;
  vadj.n = nv;
  vadj.rowptr.assign(nv + 1, 0);
  for (int64_t f = 0; f < d.E; ++f) {
    const int32_t a = vid[d.facets[2 * f]], b = vid[d.facets[2 * f + 1]];
    if (a >= 0 && b >= 0) { vadj.rowptr[a + 1]++; vadj.rowptr[b + 1]++; }
  }
  for (int32_t a = 0; a < nv; ++a) vadj.rowptr[a + 1] += vadj.rowptr[a];
  vadj.col.resize(vadj.rowptr[nv]);
  {
    std::vector<int32_t> fill(vadj.rowptr.begin(), vadj.rowptr.end() - 1);
    for (int64_t f = 0; f < d.E; ++f) {
      const int32_t a = vid[d.facets[2 * f]], b = vid[d.facets[2 * f + 1]];
      if (a >= 0 && b >= 0) { vadj.col[fill[a]++] = b; vadj.col[fill[b]++] = a; }
    }
  }
  std::vector<double> vx(nv), vy(nv);
  for (int32_t a = 0; a < nv; ++a) { vx[a] = x[int_of[vdof[a]]]; vy[a] = y[int_of[vdof[a]]]; }
  SymbolicOptions vopt = opt;
  vopt.leaf_nodes = std::max(1, opt.leaf_nodes / 4);           // a vertex brings ~3 edge nodes along
  vopt.search_min_nodes = std::max(1, opt.search_min_nodes / 4);
  Forest vf; std::vector<int32_t> vroots;
  const double t1 = clk();
  double t1b;
  {
    Dissector D(vadj, vx.data(), vy.data(), vopt);
    t1b = clk();
    D.split_chains = false;
    Dissector::Scratch sc;
    D.dissect(0, nv, vf, vroots, sc, host_threads());
  }
  const double t2 = clk();
  // depth of every tree node, tree node of every vertex
  const int32_t nt = (int32_t)vf.nodes.size();
  std::vector<int32_t> depth(nt, 0), tn_of(nv, -1);
  {
    std::vector<int32_t> st(vroots.begin(), vroots.end());
    while (!st.empty()) {
      const int32_t t = st.back(); st.pop_back();
      const int32_t* ch = vf.children_of(t);
      for (int32_t q = 0; q < vf.nodes[t].ch_n; ++q) { depth[ch[q]] = depth[t] + 1; st.push_back(ch[q]); }
    }
    for (int32_t t = 0; t < nt; ++t) { const int32_t* o = vf.own_of(t); for (int32_t q = 0; q < vf.nodes[t].own_n; ++q) tn_of[o[q]] = t; }
  }
  // lift: own lists in interior P2 indices (one pool, a slice per tree node: its vertices, then its edge nodes by facet id)
  std::vector<int32_t> lptr(nt + 1, 0), lown, orphans, tn_of_facet(d.E, -1);
  for (int32_t t = 0; t < nt; ++t) lptr[t + 1] = vf.nodes[t].own_n;
  for (int64_t f = 0; f < d.E; ++f) {
    if (int_of[V + f] < 0) continue;
    const int32_t a = vid[d.facets[2 * f]], b = vid[d.facets[2 * f + 1]];
    const int32_t ta = a >= 0 ? tn_of[a] : -1, tb = b >= 0 ? tn_of[b] : -1;
    int32_t t = ta;
    if (ta < 0 || (tb >= 0 && depth[tb] > depth[ta])) t = tb;
    if (t < 0) { orphans.push_back(int_of[V + f]); continue; }
    tn_of_facet[f] = t; lptr[t + 1]++;
  }
  for (int32_t t = 0; t < nt; ++t) lptr[t + 1] += lptr[t];
  lown.resize(lptr[nt]);
  {
    std::vector<int32_t> fill(lptr.begin(), lptr.end() - 1);
    for (int32_t t = 0; t < nt; ++t) { const int32_t* o = vf.own_of(t); for (int32_t q = 0; q < vf.nodes[t].own_n; ++q) lown[fill[t]++] = int_of[vdof[o[q]]]; }
    for (int64_t f = 0; f < d.E; ++f) if (tn_of_facet[f] >= 0) lown[fill[tn_of_facet[f]]++] = int_of[V + f];
  }
  // order separators along their cut, split them into chains of <= max_sn_nodes supernodes
  auto proj = [&](int dd, int32_t v) { switch (dd) { case 0: return x[v]; case 1: return y[v]; case 2: return x[v] + y[v]; default: return x[v] - y[v]; } };
  // (keys gathered once per node and sorted as (key, node) pairs: the same order as comparing proj(odir, .) with ties by node
  //  id; the tree nodes are independent — large trees (the 2M-unknown mesh: 60 k of them) sort on the call's host threads)
  parallel_for(nt, 4096, [&](int64_t b, int64_t e) {
    std::vector<std::pair<double, int32_t>> keyed;
    for (int64_t t = b; t < e; ++t) {
      const int odir = vf.nodes[t].od;
      if (odir < 0) continue;
      int32_t* o = lown.data() + lptr[t];
      const int32_t no = lptr[t + 1] - lptr[t];
      keyed.resize(no);
      for (int32_t i = 0; i < no; ++i) keyed[i] = {proj(odir, o[i]), o[i]};
      std::sort(keyed.begin(), keyed.end());
      for (int32_t i = 0; i < no; ++i) o[i] = keyed[i].second;
    }
  });
  out = Forest(); out_roots.clear();
  out.nodes.reserve((size_t)nt + nt / 4 + 16); out.own.reserve((size_t)n); out.child.reserve((size_t)nt + nt / 4 + 16);
  std::vector<int32_t> head(nt, -1);          // tree node of `out` heading (last chain link of) vertex-tree node t
  // children before parents: process in reverse DFS order
  std::vector<int32_t> order; order.reserve(nt);
  {
    std::vector<int32_t> st(vroots.begin(), vroots.end());
    while (!st.empty()) {
      const int32_t t = st.back(); st.pop_back(); order.push_back(t);
      const int32_t* ch = vf.children_of(t);
      for (int32_t q = 0; q < vf.nodes[t].ch_n; ++q) st.push_back(ch[q]);
    }
    std::reverse(order.begin(), order.end());
  }
  std::vector<int32_t> kids;
  for (int32_t t : order) {
    const int32_t* o = lown.data() + lptr[t];
    kids.clear();
    { const int32_t* ch = vf.children_of(t); for (int32_t q = 0; q < vf.nodes[t].ch_n; ++q) kids.push_back(head[ch[q]]); }
    const int odir = vf.nodes[t].od;
    // chains of <= max_sn_nodes supernodes (a leaf larger than the limit — many edge nodes — is split as well)
    const int32_t ns = lptr[t + 1] - lptr[t];
    const int32_t nchunks = std::max(1, (ns + opt.max_sn_nodes - 1) / opt.max_sn_nodes);
    int32_t prev = -1, pos = 0;
    for (int32_t k = 0; k < nchunks; ++k) {
      const int32_t len = ns / nchunks + (k < ns % nchunks ? 1 : 0);
      prev = k == 0 ? out.add(o + pos, len, kids.data(), (int32_t)kids.size(), odir) : out.add(o + pos, len, &prev, 1, odir);
      pos += len;
    }
    head[t] = prev;
  }
  for (int32_t r : vroots) out_roots.push_back(head[r]);
  if (timing) fprintf(stderr, "[plfem] vertex dissection: graph %.2f ms, presort %.2f ms, dissect %.2f ms, lift %.2f ms (nv=%d)\n", t1 - t0, t1b - t1, t2 - t1b, clk() - t2, nv);
  if (!orphans.empty()) {
    // one more node above all roots; keep the supernode limit: chain if needed (the last max_sn_nodes orphans first)
    std::vector<int32_t> ch(out_roots);
    int32_t left = (int32_t)orphans.size();
    while (left > opt.max_sn_nodes) {
      const int32_t id = out.add(orphans.data() + left - opt.max_sn_nodes, opt.max_sn_nodes, ch.data(), (int32_t)ch.size(), -1);
      left -= opt.max_sn_nodes;
      ch.assign(1, id);
    }
    out_roots.assign(1, out.add(orphans.data(), left, ch.data(), (int32_t)ch.size(), -1));
  }
}

}  // namespace

bool front_plan_needs_adjacency(const DofTables& dof) {
  static const bool p2_graph = [] { const char* e = std::getenv("PLFEM_DISSECT_P2"); return e && e[0] == '1'; }();
  return p2_graph || dof.V == 0;
}

void build_front_plan(const DofTables& dof, const Pattern* adj_p, const double* x, const double* y, const SymbolicOptions& opt, FrontPlan& P) {
  const int32_t n = (int32_t)dof.interior.size();
  P = FrontPlan();
  P.n = n;
  static const bool timing = std::getenv("PLFEM_TIMING") != nullptr;
  const bool p2_graph = front_plan_needs_adjacency(dof);
  if (p2_graph && !adj_p) throw std::runtime_error("front plan: the P2-graph dissection needs the adjacency pattern");
  auto clk = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tA = clk();
  double tB = tA;
  std::vector<int32_t> roots;
  Forest forest;
  if (p2_graph) {       // dissect the P2 node graph itself (the first implementation; kept for comparison)
    Dissector D(*adj_p, x, y, opt);
    tB = clk();
    Dissector::Scratch sc;
    D.dissect(0, n, forest, roots, sc, host_threads());
  } else {
    dissect_vertex_graph(dof, dof.interior, x, y, opt, forest, roots);
  }
  const double tC = clk();

  // post-order numbering of tree nodes -> fronts
  const int32_t nt = (int32_t)forest.nodes.size();
  std::vector<int32_t> front_of(nt, -1), order; order.reserve(nt);
  {
    std::vector<std::pair<int32_t, size_t>> st;
    for (int32_t r : roots) {
      st.emplace_back(r, 0);
      while (!st.empty()) {
        auto& [v, k] = st.back();
        if (k < (size_t)forest.nodes[v].ch_n) { int32_t ch = forest.children_of(v)[k++]; st.emplace_back(ch, 0); }
        else { front_of[v] = (int32_t)order.size(); order.push_back(v); st.pop_back(); }
      }
    }
  }
  const int32_t nf = (int32_t)order.size();
  P.nfronts = nf;
  P.first.resize(nf); P.s.resize(nf); P.parent.assign(nf, -1); P.level.assign(nf, 0);
  P.perm.resize(n); P.sn_of.resize(n);
  std::vector<int32_t> new_of(n, -1);
  int32_t next = 0;
  for (int32_t f = 0; f < nf; ++f) {
    const Forest::Node& tn = forest.nodes[order[f]];
    const int32_t* own = forest.own_of(order[f]);
    const int32_t* kids = forest.children_of(order[f]);
    P.first[f] = next; P.s[f] = tn.own_n;
    for (int32_t q = 0; q < tn.own_n; ++q) { const int32_t v = own[q]; P.perm[next] = v; new_of[v] = next; P.sn_of[next] = f; ++next; }
    for (int32_t q = 0; q < tn.ch_n; ++q) P.parent[front_of[kids[q]]] = f;
  }
  if (next != n) throw std::runtime_error("nested dissection lost nodes");
  P.cptr.assign(nf + 1, 0);
  for (int32_t f = 0; f < nf; ++f) if (P.parent[f] >= 0) P.cptr[P.parent[f] + 1]++;
  for (int32_t f = 0; f < nf; ++f) P.cptr[f + 1] += P.cptr[f];
  P.child.resize(P.cptr[nf]);
  {
    std::vector<int32_t> fill(P.cptr.begin(), P.cptr.end() - 1);
    for (int32_t f = 0; f < nf; ++f) if (P.parent[f] >= 0) P.child[fill[P.parent[f]]++] = f;
  }

  const double tC1 = clk();
  // update sets, bottom-up (post-order guarantees children first).  Neighbours of a node = the nodes of its elements
  // (read from the element tables: the node adjacency pattern is never built on this path).
  P.sptr.assign(nf + 1, 0);
  P.strct.clear(); P.strct.reserve((size_t)n * 8);
  // The update set of a front is collected in a two-level bitmap over the new ids (a word of the upper level has a bit per
  // word of the lower one) and read back in ascending order: no duplicates to test for, no sort (sorting the ~30-entry
  // sets front by front was half of this pass).
  std::vector<uint64_t> bm0(((size_t)n + 63) / 64 + 1, 0), bm1((bm0.size() + 63) / 64 + 1, 0);
  int32_t w1lo = INT32_MAX, w1hi = -1;
  auto bm_add = [&](int32_t c) {
    const int32_t w = c >> 6, w1 = w >> 6;             // (both levels marked without asking: no data-dependent branch)
    bm1[w1] |= 1ull << (w & 63); w1lo = std::min(w1lo, w1); w1hi = std::max(w1hi, w1);
    bm0[w] |= 1ull << (c & 63);
  };
  // a whole word of the lower level at once (the children's sets arrive sorted: their entries are gathered per word in a
  // register — back-to-back updates of one word in memory wait for each other's store)
  auto bm_add_word = [&](int32_t w, uint64_t m) {
    const int32_t w1 = w >> 6;
    bm1[w1] |= 1ull << (w & 63); w1lo = std::min(w1lo, w1); w1hi = std::max(w1hi, w1);
    bm0[w] |= m;
  };
  std::vector<int32_t> drained((size_t)n + 1);
  auto bm_drain = [&](std::vector<int32_t>& out) {
    int32_t* o = drained.data();
    for (int32_t w1 = w1lo; w1 <= w1hi; ++w1) {
      uint64_t m1 = bm1[w1];
      bm1[w1] = 0;
      while (m1) {
        const int32_t w = (w1 << 6) + __builtin_ctzll(m1);
        m1 &= m1 - 1;
        uint64_t m0 = bm0[w];
        bm0[w] = 0;
        while (m0) { *o++ = (w << 6) + __builtin_ctzll(m0); m0 &= m0 - 1; }
      }
    }
    out.insert(out.end(), drained.data(), o);
    w1lo = INT32_MAX; w1hi = -1;
  };
  // Without the adjacency pattern: an element is a clique of its 6 nodes, and in a valid elimination forest they lie on one
  // root path — so it is enough to hand the element to the front of its FIRST eliminated node (the later nodes reach the
  // fronts further up through the children's update sets).  One pass over the elements instead of one over every node's
  // element ring (6x fewer look-ups).
  std::vector<int32_t> enew, eptr, elist;
  if (!adj_p) {
    std::vector<int32_t> new_of_dof(dof.N, -1);
    for (int32_t r = 0; r < n; ++r) new_of_dof[dof.interior[P.perm[r]]] = r;
    const int64_t T = dof.T;
    enew.resize(6 * T);
    std::vector<int32_t> efront(T, -1);
    eptr.assign(nf + 1, 0);
    for (int64_t e = 0; e < T; ++e) {
      int32_t mn = INT32_MAX;
      for (int k = 0; k < 6; ++k) {
        const int32_t c = new_of_dof[dof.edofs[6 * e + k]];
        enew[6 * e + k] = c;
        if (c >= 0 && c < mn) mn = c;
      }
      if (mn != INT32_MAX) { efront[e] = P.sn_of[mn]; eptr[efront[e] + 1]++; }
    }
    for (int32_t f = 0; f < nf; ++f) eptr[f + 1] += eptr[f];
    elist.resize(eptr[nf]);
    std::vector<int32_t> fill(eptr.begin(), eptr.end() - 1);
    for (int64_t e = 0; e < T; ++e) if (efront[e] >= 0) elist[fill[efront[e]]++] = (int32_t)e;
  }
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t last = P.first[f] + P.s[f] - 1;
    const size_t b = P.strct.size();
    if (adj_p) {
      for (int32_t r = P.first[f]; r <= last; ++r) {
        const int32_t v = P.perm[r];
        for (int32_t q = adj_p->rowptr[v]; q < adj_p->rowptr[v + 1]; ++q) {
          const int32_t c = new_of[adj_p->col[q]];
          if (c > last) bm_add(c);
        }
      }
    } else {
      for (int32_t q = eptr[f]; q < eptr[f + 1]; ++q) {
        const int32_t* en = &enew[6 * (int64_t)elist[q]];
        for (int k = 0; k < 6; ++k) if (en[k] > last) bm_add(en[k]);
      }
    }
    for (int32_t q = P.cptr[f]; q < P.cptr[f + 1]; ++q) {
      const int32_t ch = P.child[q];
      int32_t k = P.sptr[ch];
      const int32_t ke = P.sptr[ch + 1];
      if (k < ke && P.strct[k] < P.first[f]) throw std::runtime_error("front plan: child update set escapes its parent");
      while (k < ke && P.strct[k] <= last) ++k;           // (sorted: the parent's own nodes come first)
      while (k < ke) {
        const int32_t w = P.strct[k] >> 6;
        uint64_t m = 0;
        do { m |= 1ull << (P.strct[k] & 63); ++k; } while (k < ke && (P.strct[k] >> 6) == w);
        bm_add_word(w, m);
      }
      P.level[f] = std::max(P.level[f], P.level[ch] + 1);
    }
    bm_drain(P.strct);
    P.sptr[f + 1] = (int32_t)P.strct.size();
    if (P.parent[f] < 0 && P.sptr[f + 1] != (int32_t)b) throw std::runtime_error("front plan: root front has an update set");
  }

  const double tC2 = clk();
  // child -> parent position maps
  P.cmap_ptr.assign(nf + 1, 0);
  for (int32_t f = 0; f < nf; ++f) P.cmap_ptr[f + 1] = P.cmap_ptr[f] + (P.parent[f] >= 0 ? P.sptr[f + 1] - P.sptr[f] : 0);
  P.cmap.resize(P.cmap_ptr[nf]);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t pa = P.parent[f];
    if (pa < 0) continue;
    const int32_t pf = P.first[pa], ps = P.s[pa];
    const int32_t* pst = P.strct.data() + P.sptr[pa];
    const int32_t pu = P.sptr[pa + 1] - P.sptr[pa];
    int32_t w = 0;
    for (int32_t k = P.sptr[f], o = P.cmap_ptr[f]; k < P.sptr[f + 1]; ++k, ++o) {
      const int32_t c = P.strct[k];
      if (c < pf + ps) { P.cmap[o] = c - pf; continue; }
      while (w < pu && pst[w] < c) ++w;
      if (w >= pu || pst[w] != c) throw std::runtime_error("front plan: child index missing from parent");
      P.cmap[o] = ps + w;
    }
  }

  const double tD = clk();
  if (timing) fprintf(stderr, "[plfem] front plan: presort %.2f ms, dissect %.2f ms, fronts/maps %.2f ms (numbering %.2f, update sets %.2f, cmap %.2f)\n", tB - tA, tC - tB, tD - tC, tC1 - tC, tC2 - tC1, tD - tC2);
  // storage offsets, statistics, level schedule
  P.foff.assign(nf + 1, 0);
  for (int32_t f = 0; f < nf; ++f) {
    const int64_t s2 = 2 * (int64_t)P.s[f], u2 = 2 * (int64_t)(P.sptr[f + 1] - P.sptr[f]);
    const int64_t m = s2 + u2;
    P.foff[f + 1] = P.foff[f] + m * m;
    P.factor_entries += s2 * s2 + s2 * u2;
    P.factor_flops += 2.0 * s2 * s2 * s2 + 2.0 * s2 * s2 * u2 + 2.0 * s2 * u2 * u2;
    P.max_front = std::max<int32_t>(P.max_front, (int32_t)(m / 2));
    P.max_s = std::max(P.max_s, P.s[f]);
    P.nlevels = std::max(P.nlevels, P.level[f] + 1);
  }
  P.lptr.assign(P.nlevels + 1, 0);
  for (int32_t f = 0; f < nf; ++f) P.lptr[P.level[f] + 1]++;
  for (int32_t l = 0; l < P.nlevels; ++l) P.lptr[l + 1] += P.lptr[l];
  P.lfront.resize(nf);
  {
    std::vector<int32_t> fill(P.lptr.begin(), P.lptr.end() - 1);
    for (int32_t f = 0; f < nf; ++f) P.lfront[fill[P.level[f]]++] = f;
  }
}

// ------------------------------------------------------------------------------------------------
// Forest batching: several designs as one block-diagonal problem
// ------------------------------------------------------------------------------------------------
void merge_front_plans(const std::vector<const FrontPlan*>& parts, FrontPlan& M, std::vector<int32_t>& node_off,
                       std::vector<int32_t>& front_off) {
  const int nb = (int)parts.size();
  M = FrontPlan();
  node_off.assign(nb + 1, 0); front_off.assign(nb + 1, 0);
  int64_t n_tot = 0, nf_tot = 0, ns_tot = 0, nc_tot = 0, nch_tot = 0;
  for (int b = 0; b < nb; ++b) {
    const FrontPlan& P = *parts[b];
    n_tot += P.n; nf_tot += P.nfronts; ns_tot += (int64_t)P.strct.size(); nc_tot += (int64_t)P.cmap.size(); nch_tot += (int64_t)P.child.size();
    node_off[b + 1] = (int32_t)n_tot; front_off[b + 1] = (int32_t)nf_tot;
    M.nlevels = std::max(M.nlevels, P.nlevels);
    M.max_front = std::max(M.max_front, P.max_front); M.max_s = std::max(M.max_s, P.max_s);
    M.factor_entries += P.factor_entries; M.factor_flops += P.factor_flops;
  }
  if (n_tot >= (int64_t(1) << 30) || ns_tot >= (int64_t(1) << 30)) throw std::runtime_error("batch too large for 32-bit plan indices");
  M.n = (int32_t)n_tot; M.nfronts = (int32_t)nf_tot;
  // every array sized once and filled through plain pointers (copies with an offset added: loops the compiler vectorises)
  M.perm.resize(n_tot); M.sn_of.resize(n_tot);
  M.first.resize(nf_tot); M.s.resize(nf_tot); M.parent.resize(nf_tot); M.level.resize(nf_tot);
  M.sptr.resize(nf_tot + 1); M.cptr.resize(nf_tot + 1); M.cmap_ptr.resize(nf_tot + 1); M.foff.resize(nf_tot + 1);
  M.sptr[0] = 0; M.cptr[0] = 0; M.cmap_ptr[0] = 0; M.foff[0] = 0;
  M.strct.resize(ns_tot); M.cmap.resize(nc_tot); M.child.resize(nch_tot);
  auto shifted = [](const int32_t* src, int64_t cnt, int32_t add, int32_t* dst) { for (int64_t i = 0; i < cnt; ++i) dst[i] = src[i] + add; };
  int32_t so = 0, co = 0, mo = 0;
  int64_t po = 0;
  for (int b = 0; b < nb; ++b) {
    const FrontPlan& P = *parts[b];
    const int32_t no = node_off[b], fo = front_off[b], nf = P.nfronts;
    std::copy(P.perm.begin(), P.perm.end(), M.perm.begin() + no);
    shifted(P.sn_of.data(), P.n, fo, M.sn_of.data() + no);
    shifted(P.first.data(), nf, no, M.first.data() + fo);
    std::copy(P.s.begin(), P.s.end(), M.s.begin() + fo);
    std::copy(P.level.begin(), P.level.end(), M.level.begin() + fo);
    {
      const int32_t* pp = P.parent.data();
      int32_t* mp = M.parent.data() + fo;
      for (int32_t f = 0; f < nf; ++f) mp[f] = pp[f] >= 0 ? pp[f] + fo : -1;
    }
    shifted(P.sptr.data() + 1, nf, so, M.sptr.data() + fo + 1);
    shifted(P.cptr.data() + 1, nf, co, M.cptr.data() + fo + 1);
    shifted(P.cmap_ptr.data() + 1, nf, mo, M.cmap_ptr.data() + fo + 1);
    {
      const int64_t* pf = P.foff.data() + 1;
      int64_t* mf = M.foff.data() + fo + 1;
      for (int32_t f = 0; f < nf; ++f) mf[f] = pf[f] + po;
    }
    shifted(P.strct.data(), (int64_t)P.strct.size(), no, M.strct.data() + so);
    shifted(P.child.data(), (int64_t)P.child.size(), fo, M.child.data() + co);
    std::copy(P.cmap.begin(), P.cmap.end(), M.cmap.begin() + mo);
    so += (int32_t)P.strct.size(); co += (int32_t)P.child.size(); mo += (int32_t)P.cmap.size();
    po += P.foff[nf];
  }
  M.lptr.assign(M.nlevels + 1, 0);
  for (int32_t f = 0; f < M.nfronts; ++f) M.lptr[M.level[f] + 1]++;
  for (int32_t l = 0; l < M.nlevels; ++l) M.lptr[l + 1] += M.lptr[l];
  M.lfront.resize(M.nfronts);
  std::vector<int32_t> fill(M.lptr.begin(), M.lptr.end() - 1);
  for (int32_t f = 0; f < M.nfronts; ++f) M.lfront[fill[M.level[f]]++] = f;
}

}  // namespace plfem
