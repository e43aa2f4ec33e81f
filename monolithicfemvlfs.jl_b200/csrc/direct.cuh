// direct.cuh - interface of the GPU multifrontal sparse LU (direct.cu)
#pragma once
#include "common.cuh"
#include <memory>

struct DirectSolver;
struct DirectDeleter { void operator()(DirectSolver*) const; };
using DirectPtr = std::unique_ptr<DirectSolver, DirectDeleter>;

// symbolic analysis (host) for the CSR pattern living on the device; coords [n][dim] host
DirectPtr direct_symbolic(const int* d_rowptr, const int* d_colind, long long n_total, long long n_owned,
                          long long nnz, const double* coords, int dim, int leaf, int is_complex, int nsm,
                          cudaStream_t s, long long* launches);
void direct_numeric(DirectSolver& D, const int* d_rowind, const int* d_colind, const double* d_vals);
void direct_solve(DirectSolver& D, const double* b_dev, double* x_dev);
// {nfronts, nlevels, maxfront, fill(entries), flops, pool bytes, symbolic ms, numeric ms}
void direct_info(const DirectSolver& D, double* out8);
