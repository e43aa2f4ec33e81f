// direct.cuh - interface of the GPU multifrontal sparse LU (direct.cu)
#pragma once
#include "common.cuh"
#include <memory>

struct DirectSolver;
struct DirectDeleter { void operator()(DirectSolver*) const; };
using DirectPtr = std::unique_ptr<DirectSolver, DirectDeleter>;

// symbolic analysis (host) for the CSR pattern living on the device; coords [n_total][dim] host;
// role[i] (host): 0 eliminate, 1 keep as Schur boundary, 2 ignore
DirectPtr direct_symbolic(const int* d_rowptr, const int* d_colind, long long n_total,
                          const signed char* role, long long nnz, const double* coords, int dim, int leaf,
                          int is_complex, int nsm, cudaStream_t s, long long* launches);
void direct_numeric(DirectSolver& D, const int* d_rowind, const int* d_colind, const double* d_vals);
// {pivots replaced by the static perturbation, non-finite pivots} of the last factorisation
void direct_pivot_status(const DirectSolver& D, long long* perturbed, long long* nonfinite);
void direct_solve(DirectSolver& D, const double* b_dev, double* x_dev);
// split solve for substructuring: forward accumulates -A_KE A_EE^-1 b_E into g_keep; backward takes
// the kept values and writes eliminated + kept entries of x (original numbering)
void direct_forward(DirectSolver& D, const double* b_dev, double* g_keep_dev);
void direct_backward(DirectSolver& D, const double* x_keep_dev, double* x_dev);
void direct_schur(DirectSolver& D, double* S_dev);     // S += -A_KE A_EE^-1 A_EK  (n_keep^2, row-major)
long long direct_nkeep(const DirectSolver& D);
// dense n x n LU with the same front kernels (interface system)
DirectPtr direct_dense(long long n, int is_complex, int nsm, cudaStream_t s, long long* launches);
void direct_dense_factor(DirectSolver& D, const double* S_dev);
// one partial front (np pivots, nf - np kept): link of the distributed interface chain (dist.cu)
DirectPtr direct_front(long long np, long long nf, int is_complex, int nsm, cudaStream_t s, long long* launches);
double* direct_pool(DirectSolver& D);                 // nf x nf row-major entries of that front
void direct_front_factor_async(DirectSolver& D);      // no host synchronisation
void direct_front_read_status(DirectSolver& D);       // pivot statistics (synchronises)
// block-tridiagonal n x n system (blocks of the given sizes, dense row-major S): chain of fronts
DirectPtr direct_chain(int nblocks, const long long* sizes, int is_complex, int nsm, cudaStream_t s,
                       long long* launches);
void direct_chain_factor(DirectSolver& D, const double* S_dev);
// {nfronts, nlevels, maxfront, fill(entries), flops, pool bytes, symbolic ms, numeric ms}
void direct_info(const DirectSolver& D, double* out8);
