/* vlfs_b200.h - C ABI of the B200-native assembly + solve path for the monolithic
 * potential-flow / beam / plate weak forms of MonolithicFEMVLFS.jl.
 *
 * This is the drop-in boundary behind the Gridap surface the reference calls:
 *   AffineFEOperator(a,l,X,Y)            src/Yago_freq_domain.jl:167, src/Multi_geo_freq_domain.jl:131,
 *                                        src/Khabakhpasheva_freq_domain.jl:240
 *   solve(op)                            src/Yago_freq_domain.jl:171, src/Khabakhpasheva_freq_domain.jl:241
 *   TransientConstant[Matrix]FEOperator  src/Khabakhpasheva_time_domain.jl:257, src/Periodic_Beam.jl:104
 *   Newmark(LUSolver(),dt,γ,β) + solve   src/Khabakhpasheva_time_domain.jl:258-266, src/Periodic_Beam.jl:107-116
 *
 * The host (Julia shim through `ccall`, or the Python mirror through ctypes) hands over
 * TABLES - cell coordinates, cell->DOF maps, face->cell glue selectors, reference
 * shape-function tables at the quadrature points, coefficient fields sampled at the
 * quadrature points - and a list of TERMS; the library integrates on the GPU, builds the
 * CSR pattern, scatters, multiplies and solves.  There is no CPU fallback: every compute
 * entry point fails with VLFS_ERR_CUDA when no sm_100 device is usable.
 *
 * Conventions
 *   - all indices 0-based int32 (DOF ids carry the multi-field offset already);
 *   - complex numbers are interleaved (re,im) doubles (Julia ComplexF64 / cuDoubleComplex);
 *   - the caller owns every host buffer; the library copies during the call and never
 *     keeps a host pointer after returning; all device memory lives behind vlfs_ctx;
 *   - every function returns a status: 0 ok, <0 argument error, >0 runtime class
 *     (VLFS_ERR_CUDA, VLFS_ERR_NCCL); VLFS_NOT_CONVERGED is a status, not a failure;
 *   - one in-flight call per context; no callbacks; no exceptions cross the boundary.
 */
#ifndef VLFS_B200_H
#define VLFS_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLFS_OK 0
#define VLFS_ERR_ARG (-1)
#define VLFS_ERR_STATE (-2)
#define VLFS_ERR_CUDA 1
#define VLFS_ERR_NCCL 2
#define VLFS_NOT_CONVERGED 3
#define VLFS_PIVOT_PERTURBED 4 /* status of a factorisation: finished and usable, but some pivots were
                                  tiny and replaced by a static perturbation (vlfs_direct_status);
                                  refinement in vlfs_solve recovers the accuracy or reports
                                  VLFS_NOT_CONVERGED */
#define VLFS_ERR_SINGULAR 5    /* a pivot was Inf/NaN: the factorisation is unusable (UMFPACK would
                                  raise SingularException, src/Periodic_Beam.jl:107 `LUSolver()`) */

#define VLFS_MAX_SLOTS 2

/* Operators applied to a basis function at a quadrature point (Gridap: v, ∇(v), ∇(ϕ)⋅e,
 * Δ(v), ∇∇(v), C⊙∇∇(η), ∇(v)⋅nΛ, (C⊙∇∇(η))⋅nΛ.⁺).  Number of components in brackets,
 * D = ambient dimension, S = D(D+1)/2. */
enum {
  VLFS_OP_VAL = 0,         /* [1] value                                             */
  VLFS_OP_GRAD = 1,        /* [D] physical gradient                                 */
  VLFS_OP_DDIR = 2,        /* [1] gradient . dir   (∇ₙ(ϕ) = ∇(ϕ)⋅VectorValue(0,0,1))  */
  VLFS_OP_LAPL = 3,        /* [1] Laplacian = tr(∇∇)                                */
  VLFS_OP_HESS = 4,        /* [S] symmetric Hessian components                       */
  VLFS_OP_CHESS = 5,       /* [S] C⊙∇∇ with the off-diagonal multiplicity folded in,   */
                           /*     so that HESS . CHESS = ∇∇(v)⊙(C⊙∇∇(η))             */
  VLFS_OP_GRAD_NOWN = 6,   /* [1] gradient . outward normal of the slot's own cell   */
  VLFS_OP_CHESS_NPLUS = 7, /* [D] (C⊙∇∇) . outward normal of slot 0's cell           */
  VLFS_OP_COUNT = 8
};

enum { VLFS_MEASURE_CELL = 0,   /* sqrt(det(Jt Jt^T)) of the measure slot's cell     */
       VLFS_MEASURE_FACET = 1   /* |Jt^T t_ref| of the local facet (skeletons; 1 for points) */ };

/* One field (or one side of a skeleton) seen from an integration domain.  The basis lives
 * in the reference space of `geometry cells` given by `coords`; `variant` selects the
 * table set per integration cell (local face of the glued volume cell, local edge of the
 * surface cell on a skeleton); NULL = variant 0. */
typedef struct {
  int32_t dref;              /* reference dimension of the geometry cells             */
  int32_t nv;                /* geometry nodes per cell                               */
  int32_t nvar;              /* number of table variants                              */
  int32_t ndofs;             /* local DOFs                                            */
  const double* coords;      /* [ncells][nv][D]                                       */
  const int32_t* variant;    /* [ncells] or NULL                                      */
  const double* geo_grad;    /* [nvar][nq][nv][dref]   reference gradients of the map */
  const double* val;         /* [nvar][nq][ndofs]                                     */
  const double* grad;        /* [nvar][nq][ndofs][dref]                               */
  const double* hess;        /* [nvar][nq][ndofs][dref][dref] or NULL                 */
  const int32_t* dofs;       /* [ncells][ndofs] global DOF ids                        */
  const double* ref_normal;  /* [nvar][dref] or NULL (outward reference facet normal) */
  const double* ref_tangent; /* [nvar][dref] or NULL (reference facet edge vector)    */
} vlfs_slot_desc;

/* One integration domain = triangulation + measure (Gridap `Measure(trian,degree)`). */
typedef struct {
  int32_t D;                 /* ambient dimension                                     */
  int32_t nq;                /* quadrature points per cell                            */
  int32_t nslots;            /* 1..VLFS_MAX_SLOTS                                     */
  int32_t measure_slot;      /* slot whose geometry defines the measure               */
  int32_t measure_kind;      /* VLFS_MEASURE_*                                        */
  int32_t nweights;          /* number of coefficient fields                          */
  int32_t affine;            /* 1: every cell has a constant Jacobian (enables the    */
                             /*    reference-matrix fast path); 0: general quadrature */
  int64_t ncells;
  const double* qweights;    /* [nq]                                                  */
  const double* weights;     /* [nweights][ncells][nq] coefficient fields at the points */
  const double* C;           /* [D][D][D][D] or NULL                                  */
  vlfs_slot_desc slot[VLFS_MAX_SLOTS];
} vlfs_domain_desc;

/* One bilinear term:  mat += coef * ∫ weight(x) (scale_t . OP_t(test)) . (scale_u . OP_u(trial)) dμ.
 * The local DOF vector of a domain is the concatenation of its slots; a slot with scale 0
 * is not touched by the factor.  Touched (test slot, trial slot) blocks are stored in full,
 * explicit zeros included (Gridap block semantics). */
typedef struct {
  int32_t mat;               /* target matrix                                         */
  int32_t weight;            /* -1 or coefficient-field index                         */
  double coef_re, coef_im;
  int32_t test_op, trial_op;
  double test_scale[VLFS_MAX_SLOTS];
  double trial_scale[VLFS_MAX_SLOTS];
  double test_dir[3], trial_dir[3];   /* for VLFS_OP_DDIR                             */
} vlfs_term_desc;

/* One linear term:  vec += coef * ∫ (w_re + i w_im)(x) (scale . OP(test)) dμ. */
typedef struct {
  int32_t vec;               /* target right-hand-side vector                         */
  int32_t weight_re, weight_im;  /* coefficient-field indices, -1 = absent (1 / 0)    */
  double coef_re, coef_im;
  int32_t op;
  double scale[VLFS_MAX_SLOTS];
  double dir[3];
} vlfs_rhs_desc;

typedef struct {
  int32_t method;            /* VLFS_SOLVER_*                                         */
  int32_t precond;           /* VLFS_PRECOND_*                                        */
  int32_t restart;           /* GMRES restart length                                  */
  int32_t maxit;
  double rtol;               /* on the true relative residual |b-Ax|/|b|               */
  int32_t block_size_hint;   /* block-Jacobi / ILU sweeps parameter                   */
  int32_t sweeps;
} vlfs_solver_opts;

typedef struct {
  int32_t iterations;
  int32_t converged;         /* 1: |b-Ax|/|b| <= rtol.  2: refinement stopped at the attainable accuracy of
                                double precision, normwise backward error <= 1e-14 (what LUSolver() itself
                                delivers).  0: neither (VLFS_NOT_CONVERGED)                               */
  double rel_residual;       /* true residual |b-Ax|_2 / |b|_2                                           */
  double setup_ms, solve_ms;
  double backward_error;     /* |b-Ax|_2 / (|A|_F |x|_2 + |b|_2); 0 when rtol was met without needing it */
} vlfs_solve_stats;

enum { VLFS_SOLVER_GMRES = 0, VLFS_SOLVER_BICGSTAB = 1, VLFS_SOLVER_CG = 2 };
enum { VLFS_PRECOND_NONE = 0, VLFS_PRECOND_JACOBI = 1,
       VLFS_PRECOND_BLOCK_ILU0 = 2, /* per-GPU block preconditioner: zero-fill incomplete LU (IKJ order on
                                       the CSR pattern) of the block owned rows x owned columns of this
                                       context, level-scheduled factorisation and triangular solves
                                       (ilu0.cu).  Converges on the Newmark matrices; stalls on the
                                       indefinite frequency-domain operators, where
                                       VLFS_PRECOND_SPARSE_LU is the solver */
       VLFS_PRECOND_SPARSE_LU = 3 /* GPU multifrontal LU of the matrix (the LUSolver() twin) */ };

typedef struct vlfs_ctx vlfs_ctx;

/* ---- lifetime ------------------------------------------------------------------ */
int vlfs_version(void);
int vlfs_device_count(int* n);
int vlfs_create(vlfs_ctx** ctx, int device);
int vlfs_destroy(vlfs_ctx* ctx);
const char* vlfs_last_error(const vlfs_ctx* ctx);

/* ---- problem definition (replaces MultiFieldFESpace + the weak-form closures) ------ */
/* ndofs: size of the monolithic system; is_complex: vector_type=Vector{ComplexF64};
 * nmat: matrices sharing one union pattern (1: A;  3: K,C,M of the transient operators);
 * nvec: right-hand-side vectors. */
int vlfs_system_init(vlfs_ctx* ctx, int64_t ndofs, int is_complex, int nmat, int nvec);
int vlfs_add_domain(vlfs_ctx* ctx, const vlfs_domain_desc* desc, int* domain_id);
int vlfs_add_term(vlfs_ctx* ctx, int domain_id, const vlfs_term_desc* term, int* term_id);
int vlfs_add_rhs_term(vlfs_ctx* ctx, int domain_id, const vlfs_rhs_desc* term, int* term_id);
/* frequency sweeps change scalars and coefficient fields only */
int vlfs_set_term_coef(vlfs_ctx* ctx, int domain_id, int term_id, double re, double im);
int vlfs_set_rhs_coef(vlfs_ctx* ctx, int domain_id, int term_id, double re, double im);
int vlfs_set_weights(vlfs_ctx* ctx, int domain_id, int weight, const double* values);

/* ---- SparseMatrixAssembler: symbolic + numeric ------------------------------------ */
int vlfs_build_pattern(vlfs_ctx* ctx, int64_t* nnz, int64_t* ncoo);
int vlfs_get_pattern(vlfs_ctx* ctx, int32_t* rowptr, int32_t* colind);           /* CSR   */
int vlfs_get_pattern_csc(vlfs_ctx* ctx, int32_t* colptr, int32_t* rowval, int32_t* csr_slot);
int vlfs_assemble(vlfs_ctx* ctx);                      /* all matrices and vectors        */
int vlfs_assemble_matrices(vlfs_ctx* ctx);
int vlfs_assemble_vectors(vlfs_ctx* ctx);
int vlfs_get_values(vlfs_ctx* ctx, int mat, void* values);   /* [nnz] real or complex    */
int vlfs_get_vector(vlfs_ctx* ctx, int vec, void* values);   /* [ndofs]                  */
/* A matrix the host already holds (Gridap's `numerical_setup(ss, A)` hands the LinearSolver an
 * assembled SparseMatrixCSC/CSR: src/Periodic_Beam.jl:107 `Newmark(LUSolver(),...)`): 0-based CSR
 * pattern shared by nmat value arrays; replaces vlfs_system_init + domains + vlfs_build_pattern.
 * vlfs_set_values / vlfs_set_vector overwrite the device copy of a value array / a rhs vector. */
int vlfs_import_csr(vlfs_ctx* ctx, int64_t ndofs, int is_complex, int nmat, const int32_t* rowptr,
                    const int32_t* colind);
int vlfs_set_values(vlfs_ctx* ctx, int mat, const void* values);
int vlfs_set_vector(vlfs_ctx* ctx, int vec, const void* values);
/* mat[target] = sum_k coef[k] * mat[src[k]], n <= 4  (Newmark: K + γ/(βΔt) C + 1/(βΔt²) M; ω sweeps by
 * basis matrices: A(ω) = B₀ + ω B₁ + ω² B₂ + k B₃, src/Yago_freq_domain.jl:69-94,162-165)          */
int vlfs_combine(vlfs_ctx* ctx, int target, int n, const int* src, const double* coef_re_im);
/* One more value array on the built pattern that no term targets (zero-filled): the target of
 * vlfs_combine when all nmat <= 4 assembled matrices are sources.  *mat receives its index (>= nmat),
 * valid wherever a matrix index is taken except vlfs_add_term.  Dropped by vlfs_system_init.      */
int vlfs_add_work_matrix(vlfs_ctx* ctx, int* mat);

/* ---- algebra -------------------------------------------------------------------- */
int vlfs_spmv(vlfs_ctx* ctx, int mat, const void* x, void* y);         /* host in/out   */
/* DOF node coordinates [ndofs][dim] (Gridap: get_cell_points(get_fe_dof_basis(V))); drives the
 * geometric nested dissection of VLFS_PRECOND_SPARSE_LU.  Optional: without them graph distances
 * from three far-apart unknowns are used instead (level-structure dissection). */
int vlfs_set_dof_coords(vlfs_ctx* ctx, int dim, const double* coords);
int vlfs_solver_setup(vlfs_ctx* ctx, int mat, const vlfs_solver_opts* opts);
/* refactorise the current values of `mat` keeping the symbolic analysis (ω sweeps): the elimination tree of
 * VLFS_PRECOND_SPARSE_LU, the dependency levels of VLFS_PRECOND_BLOCK_ILU0 */
int vlfs_solver_refactor(vlfs_ctx* ctx, int mat);
/* {nfronts, nlevels, maxfront, fill entries, flops, pool bytes, symbolic ms, numeric ms} */
int vlfs_direct_info(vlfs_ctx* ctx, double* out8);
/* pivot statistics of the last (re)factorisation (sparse LU or ILU(0), whichever the solver uses): pivots
 * replaced by the static perturbation (LU: after the in-block threshold partial pivoting), non-finite pivots */
int vlfs_direct_status(vlfs_ctx* ctx, int64_t* perturbed, int64_t* nonfinite);
/* z = M^-1 r with the preconditioner of the last vlfs_solver_setup (host buffers, ndofs entries in,
 * owned entries out); {lower levels, upper levels, lower launches, upper launches, level analysis ms (host),
 * factorisation ms, last vlfs_precond_apply ms (CUDA events), entries of the block rows} of the ILU(0) */
int vlfs_precond_apply(vlfs_ctx* ctx, const void* r, void* z);
int vlfs_precond_info(vlfs_ctx* ctx, double* out8);
int vlfs_solve(vlfs_ctx* ctx, const void* b, void* x, vlfs_solve_stats* stats);   /* host */
int vlfs_solve_vec(vlfs_ctx* ctx, int vec, void* x, vlfs_solve_stats* stats); /* assembled rhs */

/* ---- Newmark(ls,dt,γ,β) on a constant operator -------------------------------------
 * Solves (K + γ/(βΔt) C + 1/(βΔt²) M) u_{n+1} = Σ_m g_m(t_{n+1}) b_m - C v* - M a*  for
 * nsteps steps; mats = {K,C,M} indices, rhs vectors b_m are the assembled vectors 0..nb-1,
 * tcoef[nsteps][nb] their time factors.  out receives the DOFs [out_lo,out_hi) of every
 * `out_stride`-th step.  u0,v0,a0 are updated in place to the final state. */
int vlfs_newmark_run(vlfs_ctx* ctx, int matK, int matC, int matM, int nsteps, double dt,
                     double gamma, double beta, int nb, const double* tcoef,
                     double* u0, double* v0, double* a0,
                     int64_t out_lo, int64_t out_hi, int out_stride, double* out,
                     vlfs_solve_stats* stats);

/* ---- device-resident benchmarking hooks (no host copies) --------------------------- */
int vlfs_spmv_device(vlfs_ctx* ctx, int mat, int reps, float* ms_total);
int vlfs_assemble_timed(vlfs_ctx* ctx, int reps, float* ms_total);
/* per-kernel CUDA-event times of one assembly pass (averaged over reps); names ';'-joined */
int vlfs_assemble_profile(vlfs_ctx* ctx, int reps, char* names, int names_cap, float* ms, int ms_cap, int* nrec);
int vlfs_last_kernel_launches(vlfs_ctx* ctx, int64_t* n);
/* measured FP64 FMA rate of the device (TFLOP/s, dependent-free DFMA chains on every SM): the roofline
 * denominator of the general-quadrature kernel and of the LU's Schur updates */
int vlfs_measure_fp64_peak(vlfs_ctx* ctx, double* tflops);
int vlfs_device_ptrs(vlfs_ctx* ctx, int mat, void** rowptr, void** colind, void** values);

/* ---- functionals of FE functions (the integrals the reference evaluates after each step:
 * l2(x) = √(∑∫(x⋅x)dΩ) and the energies, src/Periodic_Beam.jl:119-120,141-148) ------------------
 * value = ∑_cells ∫ | Σ_s scale_s OP(u_h) − f |² dμ on a registered domain; u_h from the DOF
 * vector x (host, real, length ndofs); f = coefficient-field index of the domain (update it with
 * vlfs_set_weights, e.g. the analytic solution at the points at time t) or -1. */
typedef struct {
  int32_t op;                          /* VLFS_OP_*                                         */
  int32_t weight_f;                    /* -1 or coefficient-field index (scalar operators)  */
  double scale[VLFS_MAX_SLOTS];
  double dir[3];
} vlfs_functional_desc;
int vlfs_eval_functional(vlfs_ctx* ctx, int domain_id, const vlfs_functional_desc* f,
                         const void* x_host, double* value);

/* ---- distributed building blocks (row-partitioned systems, one context per GPU) ------------
 * A rank registers the cells that touch its owned DOFs with LOCAL DOF ids [owned | ghost]; rows
 * [0,n_owned) of the local CSR are then complete.  The entry points below take DEVICE pointers
 * (vectors of the local length, scalars interleaved as elsewhere) and run on the stream given to
 * vlfs_set_stream (e.g. the caller's torch / CUDA.jl stream); halo exchange and the all-reduce of
 * the dot products are the caller's with these low-level calls (see the multi-GPU section below
 * for the path in which the library does both itself).
 * Replaces nothing in the (serial) reference: SURVEY.md §8(e). */
int vlfs_set_stream(vlfs_ctx* ctx, void* cuda_stream);          /* before vlfs_solver_setup     */
int vlfs_set_owned_rows(vlfs_ctx* ctx, int64_t n_owned);
int vlfs_dev_spmv(vlfs_ctx* ctx, int mat, const void* x_dev, void* y_dev);   /* owned rows      */
int vlfs_dev_precond(vlfs_ctx* ctx, const void* r_dev, void* z_dev);         /* owned block     */
/* out_host[k] = <V_k, w> over n entries (conjugate-linear in V), V_k = V + k*ldv scalars        */
int vlfs_dev_multidot(vlfs_ctx* ctx, const void* V_dev, int64_t ldv, int nv, const void* w_dev,
                      int64_t n, void* out_host);
int vlfs_dev_gemv_sub(vlfs_ctx* ctx, const void* V_dev, int64_t ldv, int nv, const void* h_host,
                      void* w_dev, int64_t n);                                /* w -= sum h_k V_k */
int vlfs_dev_axpby(vlfs_ctx* ctx, double a_re, double a_im, const void* x_dev, double b_re,
                   double b_im, void* y_dev, int64_t n);                      /* y = a x + b y   */
int vlfs_dev_gather(vlfs_ctx* ctx, int64_t n, const int32_t* idx_dev, const void* src_dev,
                    void* dst_dev);                                           /* halo pack       */
int vlfs_get_vector_device(vlfs_ctx* ctx, int vec, void** ptr);

/* ---- multi-GPU systems: the library owns the communicator and the whole data plane ------------
 * Every rank (one context per GPU) holds the rows of the DOFs it owns, with LOCAL ids [owned | ghosts
 * grouped by owner]; it registers the cells that touch an owned DOF, so that assembly needs no
 * communication.  Ranks are threads of ONE process (vlfs_multi_create: ncclCommInitAll; what the Julia
 * shim uses) or one process each (vlfs_dist_unique_id on rank 0, the host hands the 128 bytes to every
 * rank, vlfs_dist_init).  After vlfs_dist_set_halo the ordinary entry points become collective:
 *   vlfs_spmv / vlfs_dev_spmv / vlfs_spmv_device : pack + grouped ncclSend/ncclRecv of the ghost values
 *       on a side stream while the rows without ghost columns are multiplied, then the others;
 *   vlfs_solver_setup(VLFS_PRECOND_SPARSE_LU) / vlfs_solver_refactor : local multifrontal LU with the
 *       rank's own cut and the next rank's cut kept, then the rank's link of the block-tridiagonal
 *       interface chain (corner relayed from rank-1, own corner passed to rank+1) - requires that the
 *       cuts form a chain (slab-like partitions; VLFS_ERR_STATE otherwise);
 *   vlfs_solve / vlfs_solve_vec : GMRES whose dot products are ONE ncclAllReduce per iteration, with
 *       the substructuring solve as exact preconditioner (x: n_local entries, owned part meaningful).
 * Ranks that share one device (tests on a 1-GPU box) use an in-process transport instead of NCCL.
 * The vlfs_multi_* calls run the per-rank call on one host thread per rank (collectives block on each
 * other); arrays of per-rank results are indexed by rank.  Replaces nothing in the serial reference:
 * SURVEY.md §8(e). */
int vlfs_dist_unique_id(void* id128);
int vlfs_dist_init(vlfs_ctx* ctx, int rank, int nranks, const void* id128);
int vlfs_multi_create(vlfs_ctx** ctxs, int ngpu, const int* devices);
/* peers[p]: rank exchanged with; send_idx[send_ptr[p]..send_ptr[p+1]) owned local ids peer p needs, in
 * the order of its ghost block; recv_ptr[p]..recv_ptr[p+1]: range of this rank's ghost block owned by
 * peer p (the ranges cover the ghost block).  After vlfs_build_pattern. */
int vlfs_dist_set_halo(vlfs_ctx* ctx, int64_t n_owned, int npeers, const int32_t* peers,
                       const int64_t* send_ptr, const int32_t* send_idx, const int64_t* recv_ptr);
/* {own cut size, next cut size, interface-chain factorisation ms (incl. waiting for rank-1), rows with ghost columns} */
int vlfs_dist_info(vlfs_ctx* ctx, double* out4);
int vlfs_multi_build_pattern(vlfs_ctx** ctxs, int n, int64_t* nnz, int64_t* ncoo);
int vlfs_multi_assemble(vlfs_ctx** ctxs, int n);
int vlfs_multi_solver_setup(vlfs_ctx** ctxs, int n, int mat, const vlfs_solver_opts* opts);
int vlfs_multi_solver_refactor(vlfs_ctx** ctxs, int n, int mat);
int vlfs_multi_solve_vec(vlfs_ctx** ctxs, int n, int vec, void** x_hosts, vlfs_solve_stats* stats);
int vlfs_multi_spmv(vlfs_ctx** ctxs, int n, int mat, const void* const* x_hosts, void** y_hosts);
int vlfs_multi_spmv_device(vlfs_ctx** ctxs, int n, int mat, int reps, float* ms_total);

/* Substructuring (exact distributed solve).  role[i]: 0 eliminate, 1 keep (interface unknown,
 * never a pivot), 2 ignore.  After vlfs_solver_setup(VLFS_PRECOND_SPARSE_LU) the factorisation is
 * partial: vlfs_schur_get accumulates S += -A_KE A_EE^-1 A_EK into an n_keep x n_keep row-major
 * device matrix (kept unknowns in ascending local id), vlfs_schur_forward accumulates
 * g += -A_KE A_EE^-1 b_E, vlfs_schur_backward solves the eliminated unknowns for given kept values.
 * vlfs_dense_factor / vlfs_chain_factor + vlfs_dense_solve: LU of the assembled interface system
 * (dense, or block tridiagonal by cuts) on one GPU. */
int vlfs_set_roles(vlfs_ctx* ctx, const int8_t* role);
int vlfs_schur_size(vlfs_ctx* ctx, int64_t* n_keep);
int vlfs_schur_get(vlfs_ctx* ctx, void* S_dev);
int vlfs_schur_forward(vlfs_ctx* ctx, const void* b_dev, void* g_keep_dev);
int vlfs_schur_backward(vlfs_ctx* ctx, const void* x_keep_dev, void* x_dev);
int vlfs_dense_factor(vlfs_ctx* ctx, int64_t n, const void* S_dev);
/* block-tridiagonal S (slab cuts in rank order, blocks of `sizes`): chain of fronts, O(sum n_r^3) */
int vlfs_chain_factor(vlfs_ctx* ctx, int nblocks, const int64_t* sizes, const void* S_dev);
int vlfs_dense_solve(vlfs_ctx* ctx, const void* b_dev, void* x_dev);

#ifdef __cplusplus
}
#endif
#endif
