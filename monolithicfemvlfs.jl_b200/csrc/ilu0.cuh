// ilu0.cuh - zero-fill incomplete LU of the owned block (ilu0.cu): state shared with api.cu / krylov.cu
#pragma once
#include "common.cuh"
#include <memory>
#include "ilu0_host.hpp"

struct Ilu0 {
  long long n = 0, nnz = 0;                    // owned rows; entries of those rows
  int cplx = 0;
  const int* rowptr = nullptr;                 // the context's pattern (device)
  const int* colind = nullptr;
  cudaStream_t s = nullptr;
  int block_rows = 0, nblocks = 1;             // block_rows > 0: diagonal blocks of that many rows (block Jacobi)
  DevBuf<int> diag, lrows, lptr, urows, uptr, lblk, ublk, info;
  DevBuf<double> lu, scale;
  std::vector<int> lptr_host, uptr_host;
  std::vector<vlfs_ilu::Launch> lplan, uplan;
  int nlevels_lower = 0, nlevels_upper = 0;
  long long perturbed = 0, nonfinite = 0;      // pivot statistics of the last factorisation
  double symbolic_ms = 0, numeric_ms = 0;      // host analysis (wall), last factorisation (CUDA events)
};

std::unique_ptr<Ilu0> ilu0_symbolic(const int* rowptr_dev, const int* colind_dev, long long n, int is_complex, int block_rows,
                                    cudaStream_t s);
void ilu0_numeric(Ilu0& P, const double* vals, long long* launches);
void ilu0_apply(Ilu0& P, const double* r, double* z, long long* launches);
