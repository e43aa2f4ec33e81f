// solver.cuh - Krylov solver state shared between api.cu and krylov.cu
#pragma once
#include "common.cuh"
#include <memory>
#include "direct.cuh"
#include "ilu0.cuh"

// Rank-local view of a row-partitioned system (dist.cu): the Krylov loops call these instead of the
// plain local operations when the context belongs to a multi-GPU system.  Everything is ordered
// on the context's stream; vectors have n_local = owned + ghost entries.
struct DistHooks {
  long long n_local = 0;
  virtual ~DistHooks() {}
  // y = alpha A x + beta y over the owned rows; fills the ghost block of x first (halo exchange
  // overlapped with the rows that need no ghost value)
  virtual void spmv(const double* vals, double* x, double* y, const double* alpha, const double* beta) = 0;
  virtual void exchange(double* x) = 0;                             // ghost block of x <- owners (on the context's stream)
  virtual void allreduce_sum(double* dev, int ndoubles) = 0;        // in place, over all ranks
  virtual bool has_exact_apply() const = 0;                         // substructuring solver set up
  virtual void apply(const double* r, double* z) = 0;               // z = A^-1 r (all ranks together)
};

struct SolverState;
struct SolverDeleter { void operator()(SolverState*) const; };

using SolverPtr = std::unique_ptr<SolverState, SolverDeleter>;

SolverPtr solver_create(const int* rowptr, const int* colind, const double* vals,
                        long long n, long long nnz, int is_complex,
                                           const vlfs_solver_opts& opts, int nsm, cudaStream_t s,
                                           long long* launches, DirectSolver* direct, DistHooks* dist = nullptr,
                                           Ilu0* ilu = nullptr);
void solver_set_matrix(SolverState& S, const double* vals);   // after a refactorisation of another matrix
int solver_solve(SolverState& S, const double* b_dev, double* x_dev, vlfs_solve_stats* stats);
int newmark_run(SolverState& S, const int* rowptr, const int* colind, const double* Kvals, const double* Cvals,
                const double* Mvals, long long n, long long nnz, int nsteps, double dt, double gamma,
                double beta, int nb, double* const* bvecs, const double* tcoef, double* u0, double* v0,
                double* a0, long long out_lo, long long out_hi, int out_stride, double* out,
                vlfs_solve_stats* stats, int nsm, cudaStream_t s, long long* launches);

void solver_precond(SolverState& S, const double* r_dev, double* z_dev);
// device-pointer vector algebra (krylov.cu)
void dev_multidot(const double* V, size_t ldv, int nv, const double* w, long long n, int cplx,
                  double* partial, int nred, double* out_dev, cudaStream_t s);
void dev_gemv_sub(const double* V, size_t ldv, int nv, const double* h_dev, double* w, long long n, int cplx,
                  int nsm, cudaStream_t s);
void dev_axpby(double are, double aim, const double* x, double bre, double bim, double* y, long long n, int cplx,
               int nsm, cudaStream_t s);
void dev_gather(const int* idx, const double* src, double* dst, long long n, int cplx, int nsm, cudaStream_t s);
