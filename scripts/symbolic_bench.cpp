// CPU driver for the host side of the sparse LU analysis (csrc/symbolic_host.hpp): reads a CSR
// pattern + DOF coordinates from a raw binary file, runs the analysis, checks the elimination tree
// and prints timings as one JSON line.
//   file layout: int64 n, int64 nnz, int32 dim, int32 leaf, int32 rp[n+1], int32 ci[nnz],
//                double coords[n*dim], int8 role[n]
//   g++ -O2 -std=c++17 -pthread -I monolithicfemvlfs.jl_b200/csrc scripts/symbolic_bench.cpp -o /tmp/symbolic_bench
#include "symbolic_host.hpp"
#include <cstdint>
#include <cstring>
#include <fstream>
#include <set>

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: symbolic_bench pattern.bin\n"); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  int64_t n = 0, nnz = 0; int32_t dim = 0, leaf = 0;
  f.read((char*)&n, 8); f.read((char*)&nnz, 8); f.read((char*)&dim, 4); f.read((char*)&leaf, 4);
  std::vector<int> rp(n + 1), ci(nnz);
  std::vector<double> xy((size_t)n * dim);
  std::vector<signed char> role(n);
  f.read((char*)rp.data(), (n + 1) * 4); f.read((char*)ci.data(), nnz * 4);
  f.read((char*)xy.data(), (size_t)n * dim * 8); f.read((char*)role.data(), n);
  if (!f) { fprintf(stderr, "short file\n"); return 2; }
  vlfs_sym::Analysis R;
  auto t0 = std::chrono::steady_clock::now();
  try {
    vlfs_sym::analyse(n, rp.data(), ci.data(), role.data(), xy.data(), dim, leaf, R);
  } catch (const std::string& e) { fprintf(stderr, "error: %s\n", e.c_str()); return 1; }
  const double total = vlfs_sym::ms_since(t0);
  // ---- validity of the analysis
  const int nf = (int)R.nodes.size();
  bool ok = true;
  std::vector<char> seen(R.n + R.n_keep, 0);
  for (int64_t i = 0; i < n; ++i) {
    if (role[i] == 2) { ok &= R.pos[i] == -1; continue; }
    ok &= R.pos[i] >= 0 && R.pos[i] < R.n + R.n_keep && !seen[R.pos[i]];
    if (R.pos[i] >= 0) seen[R.pos[i]] = 1;
    ok &= (role[i] == 0) == (R.pos[i] < R.n);
  }
  for (int fr = 0; fr < nf; ++fr)
    for (int c : R.nodes[fr].child) ok &= c < fr;                    // children before parents
  long long fill = 0, maxnf = 0;
  double flops = 0;
  for (int fr = 0; fr < nf && ok; ++fr) {
    const int np = (int)R.nodes[fr].piv.size(), last = R.first[fr] + np;
    const std::vector<int>& u = R.upd[fr];
    ok &= std::is_sorted(u.begin(), u.end()) && std::adjacent_find(u.begin(), u.end()) == u.end();
    ok &= u.empty() || u.front() >= last;
    // every later neighbour of a pivot is a pivot of this front or one of its update positions
    for (int v : R.nodes[fr].piv)
      for (int p = R.sp[v]; p < R.sp[v + 1]; ++p) {
        const int w = R.pos[R.si[p]];
        if (w >= last) ok &= std::binary_search(u.begin(), u.end(), w);
      }
    // the update set is contained in the parent's pivots + update set
    const int par = R.parent[fr];
    if (par >= 0) {
      const int pf = R.first[par], pl = pf + (int)R.nodes[par].piv.size();
      for (int w : u) ok &= (w >= pf && w < pl) || std::binary_search(R.upd[par].begin(), R.upd[par].end(), w);
    } else {
      for (int w : u) ok &= w >= R.n;                                // roots only keep Schur boundary unknowns
    }
    const double a = np, b = (double)u.size();
    fill += (long long)(a * a + 2 * a * b);
    flops += (2.0 / 3.0) * a * a * a + 2 * a * a * b + 2 * a * b * b;
    maxnf = std::max<long long>(maxnf, np + (long long)u.size());
  }
  int nlevels = 0;
  for (int fr = 0; fr < nf; ++fr) nlevels = std::max(nlevels, R.level[fr] + 1);
  // order-independent digest of the permutation (to compare runs with different thread counts)
  unsigned long long digest = 1469598103934665603ull;
  for (int64_t i = 0; i < n; ++i) { digest ^= (unsigned long long)(R.pos[i] + 2); digest *= 1099511628211ull; }
  printf("{\"valid\": %s, \"n\": %lld, \"n_keep\": %lld, \"fronts\": %d, \"levels\": %d, \"max_front\": %lld, "
         "\"fill\": %lld, \"flops\": %.6e, \"pos_digest\": \"%016llx\", \"threads\": %u, "
         "\"ms\": {\"symmetrise\": %.1f, \"nested_dissection\": %.1f, \"update_sets\": %.1f, \"total\": %.1f}}\n",
         ok ? "true" : "false", (long long)R.n, (long long)R.n_keep, nf, nlevels, maxnf, fill, flops, digest,
         vlfs_sym::host_threads(), R.ms_sym, R.ms_nd, R.ms_upd, total);
  return ok ? 0 : 1;
}
