// ilu0_host.hpp - host analysis of the level-scheduled ILU(0) (plain C++17, no CUDA: compiled by nvcc into
// ilu0.cu and by g++ into tests/ilu0_host_harness.cpp for the CPU tests).
//
// Input: CSR pattern of the n owned rows (column indices strictly ascending inside a row; columns >= n are
// ghosts).  block_rows = 0: one block, the ILU(0) of the whole owned block; B > 0: diagonal blocks of B
// consecutive rows, couplings between blocks dropped.  Output: position of the diagonal of every row, the
// dependency level of every row for the factorisation / forward solve (row i after every row k < i of its
// pattern inside its block) and for the backward solve (row i after every row j > i), the rows ordered by
// level, and the launch plan.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

namespace vlfs_ilu {

struct Launch { int l0, l1, chain; };     // levels [l0,l1): one wide level (chain = 0) or a run of narrow ones

constexpr int kNarrow = 32;               // a level of at most this many rows is "narrow" (one warp per row in one CTA)

struct Schedule {
  std::vector<int> rows;                  // rows ordered by (block,) level, ascending row id inside a level
  std::vector<int> lptr;                  // level boundaries into rows (per block: its levels + one terminator)
  std::vector<int> blk;                   // blocks only: first lptr entry of every block, size nblocks + 1
  std::vector<Launch> plan;               // one block only: launches in order
  int nlevels = 0;                        // one block: number of levels; blocks: the largest count of a block
};

struct Analysis {
  long long n = 0;
  int block_rows = 0, nblocks = 1;
  std::vector<int> diag;
  std::vector<int> level_lower, level_upper;
  Schedule lower, upper;
};

// rows ordered by level (counting sort); wide levels get a launch of their own, runs of narrow levels share one
inline void plan_levels(const std::vector<int>& level, int nlev, Schedule& S) {
  const int n = (int)level.size();
  S.lptr.assign((size_t)nlev + 1, 0);
  for (int i = 0; i < n; ++i) ++S.lptr[(size_t)level[i] + 1];
  for (int l = 0; l < nlev; ++l) S.lptr[(size_t)l + 1] += S.lptr[l];
  S.rows.resize(n);
  std::vector<int> fill(S.lptr.begin(), S.lptr.end() - 1);
  for (int i = 0; i < n; ++i) S.rows[(size_t)fill[level[i]]++] = i;
  S.plan.clear();
  for (int l = 0; l < nlev;) {
    if (S.lptr[(size_t)l + 1] - S.lptr[l] > kNarrow) { S.plan.push_back({l, l + 1, 0}); ++l; continue; }
    int e = l;
    while (e < nlev && S.lptr[(size_t)e + 1] - S.lptr[e] <= kNarrow) ++e;
    S.plan.push_back({l, e, 1});
    l = e;
  }
  S.nlevels = nlev;
  S.blk.clear();
}

// rows ordered by (block, level); lptr holds the level boundaries of every block followed by one terminating
// entry per block, blk[b] the first lptr entry of block b
inline void plan_blocks(const std::vector<int>& level, long long n, int block_rows, Schedule& S) {
  const int nb = (int)((n + block_rows - 1) / block_rows);
  S.rows.resize((size_t)n);
  S.lptr.clear();
  S.blk.assign((size_t)nb + 1, 0);
  S.nlevels = 0;
  std::vector<int> cnt;
  for (int b = 0; b < nb; ++b) {
    const long long r0 = (long long)b * block_rows, r1 = std::min(n, r0 + block_rows);
    int nl = 0;
    for (long long i = r0; i < r1; ++i) nl = std::max(nl, level[(size_t)i] + 1);
    cnt.assign((size_t)nl + 1, 0);
    for (long long i = r0; i < r1; ++i) ++cnt[(size_t)level[(size_t)i] + 1];
    for (int l = 0; l < nl; ++l) cnt[(size_t)l + 1] += cnt[l];
    S.blk[b] = (int)S.lptr.size();
    for (int l = 0; l <= nl; ++l) S.lptr.push_back((int)r0 + cnt[l]);
    for (long long i = r0; i < r1; ++i) S.rows[(size_t)(r0 + cnt[level[(size_t)i]]++)] = (int)i;
    S.nlevels = std::max(S.nlevels, nl);
  }
  S.blk[nb] = (int)S.lptr.size();
  S.plan.assign(1, Launch{0, S.nlevels, 1});
}

// throws std::string on an unusable pattern
inline Analysis analyze(long long n, const int* rp, const int* ci, int block_rows) {
  Analysis A;
  A.n = n;
  if (block_rows < 0 || block_rows >= n) block_rows = 0;
  A.block_rows = block_rows;
  A.nblocks = block_rows ? (int)((n + block_rows - 1) / block_rows) : 1;
  const long long B = block_rows ? block_rows : (n > 0 ? n : 1);
  A.diag.assign((size_t)n, 0);
  A.level_lower.assign((size_t)n, 0);
  A.level_upper.assign((size_t)n, 0);
  std::vector<int>& ll = A.level_lower;
  std::vector<int>& ul = A.level_upper;
  int nl = 0, nu = 0;
  for (long long i = 0; i < n; ++i) {
    const long long r0 = i / B * B;
    int d = -1, lev = 0;
    for (int p = rp[i]; p < rp[i + 1]; ++p) {
      const int c = ci[p];
      if (p > rp[i] && c <= ci[p - 1]) throw std::string("ILU0: column indices of a row must be strictly ascending");
      if (c < i) { if (c >= r0) lev = std::max(lev, ll[(size_t)c] + 1); }      // couplings inside the diagonal block only
      else if (c == i) d = p;
    }
    if (d < 0) throw std::string("ILU0: structurally zero diagonal in row ") + std::to_string(i);
    A.diag[(size_t)i] = d; ll[(size_t)i] = lev; nl = std::max(nl, lev + 1);
  }
  for (long long i = n - 1; i >= 0; --i) {
    const long long r1 = std::min(n, (i / B + 1) * B);
    int lev = 0;
    for (int p = A.diag[(size_t)i] + 1; p < rp[i + 1]; ++p) {
      const int c = ci[p];
      if (c < r1) lev = std::max(lev, ul[(size_t)c] + 1);
    }
    ul[(size_t)i] = lev; nu = std::max(nu, lev + 1);
  }
  if (n == 0) nl = nu = 0;
  if (block_rows) { plan_blocks(ll, n, block_rows, A.lower); plan_blocks(ul, n, block_rows, A.upper); }
  else { plan_levels(ll, nl, A.lower); plan_levels(ul, nu, A.upper); }
  return A;
}

}  // namespace vlfs_ilu
