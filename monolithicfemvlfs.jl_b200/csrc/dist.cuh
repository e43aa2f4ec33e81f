// dist.cuh - multi-GPU data plane of libvlfs_b200 (dist.cu): NCCL communicator owned by the
// library, halo exchange overlapped with the interior rows of the SpMV, all-reduce of the Krylov
// dot products, and the exact distributed solve by substructuring on a chain of slab cuts.
#pragma once
#include "solver.cuh"
#include <vector>

struct DistState;
struct DistDeleter { void operator()(DistState*) const; };
using DistPtr = std::unique_ptr<DistState, DistDeleter>;

// 128-byte NCCL unique id (rank 0 creates it, the host hands it to every rank)
void dist_unique_id(void* id128);
// one communicator rank on `device`; collective over the nranks callers (processes or threads)
DistPtr dist_init_rank(int rank, int nranks, const void* id128, int device);
// all ranks of one process at once (ncclCommInitAll): rank r lives on devs[r]
void dist_init_all(int n, const int* devs, std::vector<DistPtr>& out);
void dist_bind(DistState& D, cudaStream_t s, int is_complex, int nsm, long long* launches);
int dist_rank(const DistState& D);
int dist_nranks(const DistState& D);

// Halo plan of a row-partitioned local system with LOCAL ids [owned | ghosts grouped by owner]:
// peers[p] exchanges with this rank; send_idx[send_ptr[p] .. send_ptr[p+1]) are the owned local ids
// peer p needs (in the order of ITS ghost block), recv_ptr[p] .. recv_ptr[p+1] the range of this
// rank's ghost block that peer p owns.  rowptr/colind: device CSR of the local matrix (rows that
// touch no ghost column are multiplied while the halo travels).
void dist_set_halo(DistState& D, long long n_owned, long long n_local, int npeers, const int* peers,
                   const long long* send_ptr, const int* send_idx, const long long* recv_ptr,
                   const int* rowptr_dev, const int* colind_dev, long long nnz);
bool dist_has_halo(const DistState& D);
DistHooks* dist_hooks(DistState& D);
void dist_exchange(DistState& D, double* x_dev);          // ghost block of x <- owners (on the main stream)

// Substructuring.  roles for the local multifrontal LU derived from the halo plan: owned unknowns a
// lower rank holds as ghosts (the cut Γ_r) and ghosts owned by a higher rank are kept (1), ghosts of
// lower ranks ignored (2), the rest eliminated (0).  Throws when the cuts do not form a chain
// (every rank may only exchange cut unknowns with rank-1 and rank+1).
void dist_roles(DistState& D, std::vector<signed char>& role);
// after the local partial factorisation: assemble this rank's front of the interface chain (own cut
// = pivots, next cut = update set), add the corner received from rank-1, factorise, pass the
// trailing corner to rank+1.  Collective.
void dist_schur_factor(DistState& D, DirectSolver& local, const int* rowptr, const int* colind,
                       const double* vals);
void dist_schur_info(const DistState& D, double* out4);    // {n_left, n_right, chain factor ms, 0}
