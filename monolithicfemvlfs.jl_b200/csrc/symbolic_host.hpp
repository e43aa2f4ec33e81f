// symbolic_host.hpp - host side of the sparse LU analysis (no CUDA in this file, so that it can be
// compiled and timed with plain g++: scripts/symbolic_bench.cpp, tests/test_symbolic_host.py).
//
//   1. pattern of A + A^T restricted to the eliminated unknowns (counting-sort transpose + merge of
//      two sorted rows; rows in parallel),
//   2. geometric nested dissection on the DOF coordinates: split along the axis with the most
//      distinct coordinate planes at the median plane; the coordinates are rank-transformed once
//      so that a split costs O(|S|) counting instead of three sorts,
//   3. update index sets of the fronts bottom-up, fronts of one tree level in parallel.
//
// Everything is deterministic: the threads only split index ranges, the tree and the numbering do
// not depend on their number.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace vlfs_sym {

struct TreeNode { std::vector<int> piv; int child[2] = {-1, -1}; };

struct Analysis {
  std::vector<int> sp, si;                 // symmetrised pattern (rows of eliminated unknowns)
  std::vector<TreeNode> nodes;             // fronts in postorder (children before parents)
  std::vector<int> pos;                    // original index -> permuted position (-1 ignored; kept last)
  std::vector<int> front_of_pos, first, parent, level;
  std::vector<std::vector<int>> upd;       // update positions of every front (ascending)
  long long n = 0, n_keep = 0;
  double ms_sym = 0, ms_nd = 0, ms_upd = 0;
};

inline unsigned host_threads() {
  if (const char* e = getenv("VLFS_HOST_THREADS")) { int t = atoi(e); if (t >= 1) return (unsigned)t; }
  unsigned t = std::thread::hardware_concurrency();
  return t < 1 ? 1 : (t > 32 ? 32 : t);
}

// f(begin, end) over [0, n) in contiguous chunks
template <typename F>
void parallel_ranges(long long n, long long grain, F&& f) {
  unsigned nt = host_threads();
  if (n < 2 * grain) nt = 1;
  nt = (unsigned)std::min<long long>(nt, (n + grain - 1) / grain);
  if (nt <= 1) { f(0LL, n); return; }
  std::vector<std::thread> th;
  const long long chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const long long b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([&f, b, e]() { f(b, e); });
  }
  for (auto& x : th) x.join();
}

inline double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// pattern of A + A^T without the diagonal: row i keeps j when role[i] == 0 and role[j] != 2
inline void symmetrise(long long n_total, const int* rp, const int* ci, const signed char* role,
                       std::vector<int>& sp, std::vector<int>& si) {
  // transpose by counting sort, row ranges in parallel with one histogram per range: the entries
  // of a transposed row come out in ascending order because range t precedes range t+1
  const long long nnz = rp[n_total];
  unsigned nt = host_threads();
  if (nnz < (1 << 20)) nt = 1;
  std::vector<long long> rb(nt + 1);
  for (unsigned t = 0; t <= nt; ++t) {            // ranges with about equal numbers of entries
    const long long target = nnz / nt * t;
    rb[t] = t == nt ? n_total : std::lower_bound(rp, rp + n_total + 1, (int)target) - rp;
  }
  rb[0] = 0;
  std::vector<std::vector<int>> hist(nt);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        hist[t].assign(n_total, 0);
        for (long long i = rb[t]; i < rb[t + 1]; ++i)
          for (int p = rp[i]; p < rp[i + 1]; ++p) ++hist[t][ci[p]];
      });
    for (auto& x : th) x.join();
  }
  std::vector<int> tp(n_total + 1, 0);
  {
    int acc = 0;
    for (long long c = 0; c < n_total; ++c) {
      tp[c] = acc;
      for (unsigned t = 0; t < nt; ++t) { const int h = hist[t][c]; hist[t][c] = acc; acc += h; }
    }
    tp[n_total] = acc;
  }
  std::vector<int> ti(nnz);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() {
        int* fill = hist[t].data();
        for (long long i = rb[t]; i < rb[t + 1]; ++i)
          for (int p = rp[i]; p < rp[i + 1]; ++p) ti[fill[ci[p]]++] = (int)i;
      });
    for (auto& x : th) x.join();
  }
  std::vector<std::vector<int>>().swap(hist);
  // merge of the two sorted rows: count, then fill
  auto merge_row = [&](long long i, int* out) -> int {
    if (role[i] != 0) return 0;
    int a = rp[i], ae = rp[i + 1], b = tp[i], be = tp[i + 1], k = 0;
    while (a < ae || b < be) {
      int j;
      if (b >= be || (a < ae && ci[a] < ti[b])) j = ci[a++];
      else if (a >= ae || ti[b] < ci[a]) j = ti[b++];
      else { j = ci[a]; ++a; ++b; }
      if (j == (int)i || role[j] == 2) continue;
      if (out) out[k] = j;
      ++k;
    }
    return k;
  };
  sp.assign(n_total + 1, 0);
  parallel_ranges(n_total, 1 << 14, [&](long long b, long long e) {
    for (long long i = b; i < e; ++i) sp[i + 1] = merge_row(i, nullptr);
  });
  for (long long i = 0; i < n_total; ++i) sp[i + 1] += sp[i];
  si.resize(sp[n_total]);
  parallel_ranges(n_total, 1 << 14, [&](long long b, long long e) {
    for (long long i = b; i < e; ++i) merge_row(i, si.data() + sp[i]);
  });
}

// coordinates -> plane ids per axis (equal coordinates share an id)
inline void rank_transform(long long n_total, const double* xy, int dim, std::vector<int>& plane, std::vector<int>& nplanes) {
  plane.assign((size_t)n_total * dim, 0);
  nplanes.assign(dim, 0);
  std::vector<std::thread> th;
  for (int a = 0; a < dim; ++a)
    th.emplace_back([&, a]() {
      std::vector<int> order(n_total);
      std::iota(order.begin(), order.end(), 0);
      std::sort(order.begin(), order.end(), [&](int u, int v) {
        const double cu = xy[(size_t)u * dim + a], cv = xy[(size_t)v * dim + a];
        return cu < cv || (cu == cv && u < v);
      });
      int id = -1;
      double prev = 0;
      for (long long k = 0; k < n_total; ++k) {
        const double c = xy[(size_t)order[k] * dim + a];
        if (k == 0 || c != prev) { ++id; prev = c; }
        plane[(size_t)order[k] * dim + a] = id;
      }
      nplanes[a] = id + 1;
    });
  for (auto& x : th) x.join();
}

// Coordinate-free fallback of the geometric dissection: graph distances from three far-apart
// vertices serve as "coordinates" (a level set of a breadth-first search is a separator, George's
// level-structure dissection).  Used when the host gives a matrix without DOF coordinates
// (`LUSolver()` drop-in on a matrix Gridap assembled: numerical_setup(ss, A)).
inline void pseudo_coords(long long n, const int* rp, const int* ci, std::vector<double>& xyz) {
  const int dim = 3;
  xyz.assign((size_t)n * dim, 0.0);
  // symmetric adjacency (pattern + transpose) so that distances are well defined
  std::vector<int> deg(n + 1, 0);
  for (long long i = 0; i < n; ++i)
    for (int p = rp[i]; p < rp[i + 1]; ++p) if (ci[p] != i) { ++deg[i + 1]; ++deg[ci[p] + 1]; }
  for (long long i = 0; i < n; ++i) deg[i + 1] += deg[i];
  std::vector<int> adj(deg[n]), fill(deg.begin(), deg.end() - 1);
  for (long long i = 0; i < n; ++i)
    for (int p = rp[i]; p < rp[i + 1]; ++p) if (ci[p] != i) { adj[fill[i]++] = ci[p]; adj[fill[ci[p]]++] = (int)i; }
  std::vector<int> dist(n), queue(n);
  auto bfs = [&](int src) -> int {           // returns the last vertex reached (farthest)
    std::fill(dist.begin(), dist.end(), -1);
    long long head = 0, tail = 0;
    int last = src, far = 0;
    auto run = [&](int s0) {
      dist[s0] = far; queue[tail++] = s0;
      while (head < tail) {
        const int v = queue[head++];
        last = v;
        for (int p = deg[v]; p < deg[v + 1]; ++p) { const int w = adj[p]; if (dist[w] < 0) { dist[w] = dist[v] + 1; queue[tail++] = w; } }
      }
      far = dist[last] + 1;
    };
    run(src);
    for (long long v = 0; v < n; ++v) if (dist[v] < 0) run((int)v);   // further components continue the numbering
    return last;
  };
  int a = bfs(0);
  a = bfs(a);                                 // pseudo-peripheral vertex
  int b = bfs(a);
  for (long long v = 0; v < n; ++v) xyz[(size_t)v * dim + 0] = dist[v];
  bfs(b);
  int c = 0; long long best = -1;
  for (long long v = 0; v < n; ++v) {
    xyz[(size_t)v * dim + 1] = dist[v];
    const long long m = std::min<long long>((long long)xyz[(size_t)v * dim], dist[v]);   // far from both a and b
    if (m > best) { best = m; c = (int)v; }
  }
  bfs(c);
  for (long long v = 0; v < n; ++v) xyz[(size_t)v * dim + 2] = dist[v];
}

struct NDBuilder {
  const int* ptr; const int* idx; const int* plane; int dim; int leaf;
  int* mark;                       // shared: stamp of the split that last put the vertex on its B side
  std::atomic<int>* stamp;         // shared source of unique stamps (their values never matter)
  int par_depth;                   // the two sides of a split are dissected by two threads above this depth
  std::vector<int> hist;           // scratch histogram over plane ids
  std::vector<TreeNode> nodes;     // subtree in postorder, child indices local to this builder

  // appends the nodes of a finished sub-builder (children before parents is preserved)
  int adopt(NDBuilder& sub, int root_local) {
    const int off = (int)nodes.size();
    for (TreeNode& t : sub.nodes) {
      for (int& c : t.child) if (c >= 0) c += off;
      nodes.push_back(std::move(t));
    }
    return off + root_local;
  }

  int make_leaf(std::vector<int>& S) {
    TreeNode t;
    t.piv.swap(S);
    nodes.push_back(std::move(t));
    return (int)nodes.size() - 1;
  }
  int rec(std::vector<int>& S, int depth = 0) {
    if ((int)S.size() <= leaf) return make_leaf(S);
    // split axis = the one with the most distinct coordinate planes (the best-resolved direction:
    // its separators are the smallest; the longest geometric extent is a poor guide on strongly
    // anisotropic meshes such as x-refined slabs); median plane by counting
    int ax = 0, med = 0;
    long long best = 0;
    for (int a = 0; a < dim; ++a) {
      int lo = plane[(size_t)S[0] * dim + a], hi = lo;
      for (int v : S) { const int p = plane[(size_t)v * dim + a]; lo = std::min(lo, p); hi = std::max(hi, p); }
      const int range = hi - lo + 1;
      if ((int)hist.size() < range) hist.resize(range);
      std::fill(hist.begin(), hist.begin() + range, 0);
      for (int v : S) ++hist[plane[(size_t)v * dim + a] - lo];
      long long distinct = 0;
      for (int k = 0; k < range; ++k) distinct += hist[k] > 0;
      if (distinct > best) {
        best = distinct; ax = a;
        // the sorted sequence's element number |S|/2
        long long acc = 0;
        const long long target = (long long)S.size() / 2;
        for (int k = 0; k < range; ++k) { acc += hist[k]; if (acc > target) { med = lo + k; break; } }
      }
    }
    std::vector<int> A, B;
    for (int v : S) (plane[(size_t)v * dim + ax] <= med ? A : B).push_back(v);
    if (A.empty() || B.empty()) {
      A.clear(); B.clear();
      for (int v : S) (plane[(size_t)v * dim + ax] < med ? A : B).push_back(v);
      if (A.empty() || B.empty()) return make_leaf(S);
    }
    std::vector<int>().swap(S);
    const int st = stamp->fetch_add(1) + 1;
    for (int v : B) mark[v] = st;
    std::vector<int> sep, A2;
    for (int v : A) {
      bool touch = false;
      for (int p = ptr[v]; p < ptr[v + 1] && !touch; ++p) touch = mark[idx[p]] == st;
      (touch ? sep : A2).push_back(v);
    }
    std::vector<int>().swap(A);
    int c0, c1;
    if (depth < par_depth && A2.size() > 4096 && B.size() > 4096) {
      // no edge joins A2 and B (that is what the separator is for), so the two dissections only
      // share `mark` entries that neither of them writes; the node order below is the one the
      // sequential recursion produces (left subtree, right subtree, parent)
      NDBuilder left{ptr, idx, plane, dim, leaf, mark, stamp, par_depth};
      NDBuilder right{ptr, idx, plane, dim, leaf, mark, stamp, par_depth};
      int l0 = -1, r0 = -1;
      std::thread th([&]() { l0 = left.rec(A2, depth + 1); });
      r0 = right.rec(B, depth + 1);
      th.join();
      c0 = adopt(left, l0);
      c1 = adopt(right, r0);
    } else {
      c0 = A2.empty() ? -1 : rec(A2, depth + 1);
      c1 = rec(B, depth + 1);
    }
    TreeNode t;
    t.piv.swap(sep);
    t.child[0] = c0 >= 0 ? c0 : c1;
    t.child[1] = c0 >= 0 ? c1 : -1;
    nodes.push_back(std::move(t));
    return (int)nodes.size() - 1;
  }
};

// role[i]: 0 eliminate, 1 keep as Schur boundary (numbered after the eliminated ones), 2 ignore
inline void analyse(long long n_total, const int* rp, const int* ci, const signed char* role,
                    const double* coords, int dim, int leaf, Analysis& R) {
  std::vector<int> elim, keep;
  for (long long i = 0; i < n_total; ++i) {
    if (role[i] == 0) elim.push_back((int)i);
    else if (role[i] == 1) keep.push_back((int)i);
  }
  const long long n = (long long)elim.size();
  R.n = n; R.n_keep = (long long)keep.size();
  if (n == 0) throw std::string("sparse LU: nothing to eliminate");
  auto t0 = std::chrono::steady_clock::now();
  symmetrise(n_total, rp, ci, role, R.sp, R.si);
  R.ms_sym = ms_since(t0);
  t0 = std::chrono::steady_clock::now();
  std::vector<int> plane, nplanes;
  rank_transform(n_total, coords, dim, plane, nplanes);
  std::vector<int> mark(n_total, 0);
  std::atomic<int> stamp{0};
  int par_depth = 0;
  for (unsigned t = host_threads(); t > 1; t >>= 1) ++par_depth;
  NDBuilder nd{R.sp.data(), R.si.data(), plane.data(), dim, leaf, mark.data(), &stamp, par_depth};
  nd.nodes.reserve(2 * n / leaf + 16);
  nd.rec(elim);
  R.nodes.swap(nd.nodes);
  R.ms_nd = ms_since(t0);
  t0 = std::chrono::steady_clock::now();
  const int nf = (int)R.nodes.size();
  // permutation (postorder = creation order)
  R.pos.assign(n_total, -1);
  R.front_of_pos.resize(n);
  R.first.resize(nf);
  R.parent.assign(nf, -1);
  {
    int k = 0;
    for (int f = 0; f < nf; ++f) {
      R.first[f] = k;
      for (int v : R.nodes[f].piv) { R.pos[v] = k; R.front_of_pos[k] = f; ++k; }
      for (int c : R.nodes[f].child) if (c >= 0) R.parent[c] = f;
    }
    if (k != n) throw std::string("nested dissection lost DOFs");
    for (size_t q = 0; q < keep.size(); ++q) R.pos[keep[q]] = (int)(n + q);
  }
  // update sets bottom-up: fronts of one level are independent
  R.level.assign(nf, 0);
  int nlevels = 0;
  for (int f = 0; f < nf; ++f) {
    for (int c : R.nodes[f].child) if (c >= 0) R.level[f] = std::max(R.level[f], R.level[c] + 1);
    nlevels = std::max(nlevels, R.level[f] + 1);
  }
  std::vector<std::vector<int>> by_level(nlevels);
  for (int f = 0; f < nf; ++f) by_level[R.level[f]].push_back(f);
  R.upd.assign(nf, {});
  for (int L = 0; L < nlevels; ++L) {
    const std::vector<int>& fl = by_level[L];
    parallel_ranges((long long)fl.size(), fl.size() > 64 ? 16 : 1, [&](long long b, long long e) {
      std::vector<int> tmp;
      for (long long q = b; q < e; ++q) {
        const int f = fl[q];
        const int last = R.first[f] + (int)R.nodes[f].piv.size();
        tmp.clear();
        for (int v : R.nodes[f].piv)
          for (int p = R.sp[v]; p < R.sp[v + 1]; ++p) { const int w = R.pos[R.si[p]]; if (w >= last) tmp.push_back(w); }
        for (int c : R.nodes[f].child)
          if (c >= 0) for (int w : R.upd[c]) if (w >= last) tmp.push_back(w);
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        R.upd[f] = tmp;
      }
    });
  }
  R.ms_upd = ms_since(t0);
}

}  // namespace vlfs_sym
