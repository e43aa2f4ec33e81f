// CPU harness of csrc/ilu0_host.hpp (test infrastructure): reads a CSR pattern, runs the level analysis and
// prints it as JSON.  usage: ilu0_host_harness <file> <block_rows>
// file = int64 n, int64 nnz, int32 rowptr[n+1], int32 colind[nnz]
#include "ilu0_host.hpp"
#include <cstdint>
#include <cstdio>
#include <cstdlib>

template <typename T>
static void print_vec(const char* name, const std::vector<T>& v, bool comma = true) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) printf(i ? ",%d" : "%d", (int)v[i]);
  printf("]%s", comma ? ", " : "");
}

static void print_schedule(const char* name, const vlfs_ilu::Schedule& S, bool comma) {
  printf("\"%s\": {", name);
  print_vec("rows", S.rows);
  print_vec("lptr", S.lptr);
  print_vec("blk", S.blk);
  printf("\"plan\": [");
  for (size_t i = 0; i < S.plan.size(); ++i) printf("%s[%d,%d,%d]", i ? "," : "", S.plan[i].l0, S.plan[i].l1, S.plan[i].chain);
  printf("], \"nlevels\": %d}%s", S.nlevels, comma ? ", " : "");
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t n = 0, nnz = 0;
  if (fread(&n, 8, 1, f) != 1 || fread(&nnz, 8, 1, f) != 1) return 2;
  std::vector<int> rp((size_t)n + 1), ci((size_t)nnz);
  if (fread(rp.data(), 4, rp.size(), f) != rp.size()) return 2;
  if (nnz && fread(ci.data(), 4, ci.size(), f) != ci.size()) return 2;
  fclose(f);
  try {
    vlfs_ilu::Analysis A = vlfs_ilu::analyze(n, rp.data(), ci.data(), atoi(argv[2]));
    printf("{\"n\": %lld, \"block_rows\": %d, \"nblocks\": %d, ", A.n, A.block_rows, A.nblocks);
    print_vec("diag", A.diag);
    print_vec("level_lower", A.level_lower);
    print_vec("level_upper", A.level_upper);
    print_schedule("lower", A.lower, true);
    print_schedule("upper", A.upper, false);
    printf("}\n");
  } catch (const std::string& e) {
    printf("{\"error\": \"%s\"}\n", e.c_str());
    return 1;
  }
  return 0;
}
