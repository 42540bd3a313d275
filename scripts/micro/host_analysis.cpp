// Micro-benchmark of the host analysis (no GPU): DOF tables, front plan and the merge of a forest's plans from
// pl-fem-vectoriel_b200/csrc/symbolic.cpp, timed on one mesh, with FNV digests of everything they produce — a change of the
// host code that is meant to keep the front plan must print the same digests before and after (the GPU results depend on the
// plan bit for bit; tests/test_cabi_host.py::test_front_plan_is_the_frozen_one pins three of them through the C ABI).
//
//   python scripts/micro/host_analysis_mesh.py cfg1 /tmp/cfg1.mesh           # writes V, T, p (2,V) f64, t (3,T) i64
//   g++ -O3 -std=c++17 -mavx2 -ffp-contract=off -Ipl-fem-vectoriel_b200/csrc scripts/micro/host_analysis.cpp \
//       pl-fem-vectoriel_b200/csrc/symbolic.cpp -o /tmp/host_analysis -lpthread
//   /tmp/host_analysis /tmp/cfg1.mesh [repetitions = 200] [host threads = 1]
//
// Prints the fastest repetition (the build container's timings are noisy: compare minima, or two binaries run alternately).
// Build container, one thread, cfg1 (V = 5 691, T = 11 313): DOF tables 1.0 ms; plan 4.5 ms at the start of round 2's last
// session, 3.2 ms at its end (cfg2: 9.7 -> 7.0); merge of twelve such plans 1.5 -> 0.66 ms.
#include "symbolic.h"

#include <malloc.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
using namespace plfem;

static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static uint64_t fnv(const void* p, size_t n, uint64_t h) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}
template <class T> static uint64_t hv(const std::vector<T>& v, uint64_t h) { const size_t n = v.size(); return fnv(&n, 8, fnv(v.data(), n * sizeof(T), h)); }
static uint64_t plan_digest(const FrontPlan& P) {
  uint64_t h = 1469598103934665603ull;
  h = hv(P.perm, h); h = hv(P.first, h); h = hv(P.s, h); h = hv(P.parent, h); h = hv(P.level, h); h = hv(P.sptr, h); h = hv(P.strct, h);
  h = hv(P.cmap_ptr, h); h = hv(P.cmap, h); h = hv(P.foff, h); h = hv(P.sn_of, h); h = hv(P.cptr, h); h = hv(P.child, h);
  h = hv(P.lptr, h); h = hv(P.lfront, h);
  const int64_t st[6] = {P.n, P.nfronts, P.nlevels, P.max_front, P.max_s, P.factor_entries};
  return fnv(st, sizeof st, h);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s mesh-file [repetitions] [host threads]\n", argv[0]); return 2; }
  // what the library does once per process (api.cu, tune_host_allocator_once): blocks of the analysis come from the heap
  mallopt(M_MMAP_THRESHOLD, 32 << 20); mallopt(M_TRIM_THRESHOLD, 512 << 20); mallopt(M_TOP_PAD, 16 << 20);
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  const int reps = argc > 2 ? std::max(2, atoi(argv[2])) : 200;
  set_host_threads(argc > 3 ? atoi(argv[3]) : 1);
  int64_t hd[2];
  if (fread(hd, 8, 2, f) != 2) return 2;
  const int64_t V = hd[0], T = hd[1];
  std::vector<double> p(2 * V);
  std::vector<int64_t> t(3 * T);
  if (fread(p.data(), 8, 2 * V, f) != (size_t)(2 * V) || fread(t.data(), 8, 3 * T, f) != (size_t)(3 * T)) return 2;
  fclose(f);
  double t_dof = 1e9, t_plan = 1e9, t_merge = 1e9;
  uint64_t h_dof = 0, h_plan = 0, h_merge = 0;
  FrontPlan keep;
  for (int r = 0; r < reps; ++r) {
    DofTables d;
    const double a = now();
    build_dof_tables(p.data(), t.data(), V, T, d);
    const double b = now();
    const int32_t n = (int32_t)d.interior.size();
    std::vector<double> x(n), y(n);
    for (int32_t i = 0; i < n; ++i) { x[i] = d.doflocs[d.interior[i]]; y[i] = d.doflocs[d.N + d.interior[i]]; }
    SymbolicOptions o;
    FrontPlan P;
    const double c0 = now();
    build_front_plan(d, nullptr, x.data(), y.data(), o, P);
    const double c1 = now();
    if (r) { t_dof = std::min(t_dof, b - a); t_plan = std::min(t_plan, c1 - c0); }
    h_dof = 1469598103934665603ull;
    h_dof = hv(d.edofs, h_dof); h_dof = hv(d.facets, h_dof); h_dof = hv(d.t2f, h_dof); h_dof = hv(d.doflocs, h_dof);
    h_dof = hv(d.boundary, h_dof); h_dof = hv(d.interior, h_dof); h_dof = hv(d.n2e_ptr, h_dof); h_dof = hv(d.n2e, h_dof);
    h_dof = fnv(&d.n_degenerate, 8, h_dof);
    h_plan = plan_digest(P);
    if (r + 1 == reps) keep = std::move(P);
  }
  std::vector<const FrontPlan*> parts(12, &keep);
  for (int r = 0; r < std::max(2, reps / 10); ++r) {
    FrontPlan M;
    std::vector<int32_t> no, fo;
    const double a = now();
    merge_front_plans(parts, M, no, fo);
    t_merge = std::min(t_merge, now() - a);
    h_merge = hv(fo, hv(no, plan_digest(M)));
  }
  printf("%s: V %lld T %lld | DOF tables %.3f ms, front plan %.3f ms, merge of 12 plans %.3f ms | digests dof %016llx plan %016llx merge %016llx\n",
         argv[1], (long long)V, (long long)T, t_dof, t_plan, t_merge, (unsigned long long)h_dof, (unsigned long long)h_plan,
         (unsigned long long)h_merge);
  return 0;
}
