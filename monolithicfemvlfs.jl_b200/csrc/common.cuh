// common.cuh - shared declarations of the vlfs_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/vlfs_b200.h"

#define VLFS_MAX_OPINST 12
#define VLFS_MAX_TERMS 24
#define VLFS_MAX_BLOCKS 4
#define VLFS_MAX_WEIGHTS 8
// dest encoding: v >= 0 CSR slot; VLFS_DEST_PAD = padding entry of a 4x4 tile.  (Measured on
// B200: turning single-contribution slots into plain 8-byte stores is SLOWER than reductions -
// scattered per-lane transactions cost the LSU the same either way and the branch diverges.)
#define VLFS_DEST_PAD (-1)

struct CudaError { cudaError_t e; const char* file; int line; };
struct NcclError { std::string msg; };

#define CUDA_TRY(x)                                              \
  do {                                                           \
    cudaError_t _e = (x);                                        \
    if (_e != cudaSuccess) throw CudaError{_e, __FILE__, __LINE__}; \
  } while (0)

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count) {
    release();
    if (count) CUDA_TRY(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    alloc(count);
    if (count) CUDA_TRY(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void zero(cudaStream_t s) { if (n) CUDA_TRY(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

// ---- device-side descriptors ------------------------------------------------------
struct SlotDev {
  int dref, nv, nvar, ndofs, has_hess;
  const double* coords;
  const int* variant;
  const double* geo_grad;
  const double* val;
  const double* grad;
  const double* hess;
  const int* dofs;
  const double* ref_normal;
  const double* ref_tangent;
};

// an operator instance evaluated once per cell into shared memory
struct OpInst {
  int slot, op, ncomp;
  double dir[3];
  int smem_off;     // offset (doubles) of the [nq][ncomp][ndp] array (multiple of 4)
  int ndp;          // row stride: local DOF count rounded up to a multiple of 4 (zero padded)
  int buf_stride;   // 0: cell-invariant table evaluated once per CTA; else distance between
                    //    the two pipeline buffers of this per-cell table
};

// One contraction feeding a (block, matrix, re|im) target of the tiled bilinear stage:
// target += coef * sum_q meas[q] W[q] sum_c A[q][c][i] B[q][c][j]
struct TileProd {
  int offA, offB;       // shared-memory offsets of the operator arrays (pipeline buffer 0)
  int bufA, bufB;       // their pipeline-buffer strides (0 = cell-invariant)
  int ndpA, ndpB;       // their row strides
  int ncomp, weight;    // components per point, coefficient-field index or -1
  int tref;             // >= 0: both operators are cell-invariant, no coefficient field and the
                        // domain is affine: sum_q qw[q] A B was contracted once per CTA into the
                        // shared-memory matrix at this offset ([nr4][nc4], row stride ldt)
  int ldt, nr4;         // padded column / row counts of the reference matrix
  int tref_owner;       // index of the product that computes the shared reference matrix
  double coef;          // coefficient part (re or im) times the slot scales
};
struct TileTarget {
  int mat, part;        // part: 0 = real, 1 = imaginary
  int first, count;     // range in the TileProd array
};

struct TermDev {
  int mat, weight;
  double cre, cim;
  int opi_test[VLFS_MAX_SLOTS];   // op-instance index per slot or -1
  int opi_trial[VLFS_MAX_SLOTS];
  double st[VLFS_MAX_SLOTS], su[VLFS_MAX_SLOTS];
  int ncomp;
};

struct RhsDev {
  int vec, wre, wim;
  double cre, cim;
  int opi[VLFS_MAX_SLOTS];
  double sc[VLFS_MAX_SLOTS];
  int ncomp;
};

struct BlockDev {     // touched (test slot, trial slot) block
  int ts, us, nr, nc, off;   // off: entry offset inside a cell's COO run
};

struct DomainDev {
  int D, nq, nslots, measure_slot, measure_kind, nweights, affine;
  long long ncells;
  const double* qweights;
  const double* weights;
  const double* C;
  SlotDev slot[VLFS_MAX_SLOTS];
  int nopi;
  OpInst opi[VLFS_MAX_OPINST];
  int nblocks;
  BlockDev blk[VLFS_MAX_BLOCKS];
  int entries_per_cell;
  int tile_base[VLFS_MAX_BLOCKS + 1];  // prefix of 4x4 tiles per block
  int tiles_j[VLFS_MAX_BLOCKS];        // tiles along the trial index
  int target_first[VLFS_MAX_BLOCKS + 1];
  const TileTarget* targets;
  const TileProd* prods;
  // shared-memory layout (doubles).  Geometry records have 4 buffers (computed two cells
  // ahead), per-cell operator tables and coefficient fields 3 (one cell ahead); the extra
  // buffer of each lets warps skew by one phase under the split-phase pipeline barrier.
  // Cell-invariant operator tables and reference matrices are single.
  int nqg;                             // geometry points per slot: 1 when affine, else nq
  int smem_geo_stride;                 // one geometry buffer: [nslots][nqg][12] + meas[nq]
  int smem_meas_off;                   // offset of meas inside a geometry buffer
  int smem_inv_off, smem_inv_len;      // cell-invariant operator tables
  int smem_tref_off, smem_tref_len;    // reference matrices of the invariant contractions
  int smem_buf_off, smem_buf_stride;   // per-cell buffers: weights [nweights][nq] + tables
  int nbuf;                            // number of per-cell table buffers (3, or 2 if large)
  int smem_total;                      // doubles
};

// ---- reference-matrix contributions, gathered slot by slot (integrate.cu: gather_reference) ----
// value(slot) = sum over the COO entries of the slot of  sum_products coef * sum_m geo[cell][geo_idx+m]
//               * tabs[tab_off + m*nr*nc + i*nc + j]
struct RefProd {
  int mat, ntab, geo_idx;
  long long tab_off;
  double cre, cim;
};
#define VLFS_CELLGEO 7         // per cell: measure, then the D(D+1)/2 factors meas * (J^-T J^-1) sym
struct GatherDom {
  long long coo_lo, coo_hi;
  int epc, nblocks;
  BlockDev blk[VLFS_MAX_BLOCKS];
  int rp_first[VLFS_MAX_BLOCKS + 1];
  const RefProd* rp;
  const double* geo;      // [ncells][VLFS_CELLGEO]
  const double* tabs;
  int tab_len;            // doubles in tabs
  // geometry classes (meshes with few distinct cell shapes, e.g. Cartesian): the complete
  // reference-type element block of every class is tabulated once per assembly
  int ncls;               // 0: no classes, per-cell factors are used
  const int* cls;         // [ncells] class of every cell
  const double* cls_geo;  // [ncls][VLFS_CELLGEO]
  double* ctab_m[4];      // per matrix: [block][class][local entry] scalars
  long long ctab_base;    // first entry of this domain in the context-wide contiguous tables
  long long ctab_off[VLFS_MAX_BLOCKS];   // entry offset of every block inside a class table
};

// ---- primitives -------------------------------------------------------------------
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_dev,
                        cudaStream_t s);
// stable LSD radix sort of (key, value) pairs on the bit range [bit_lo, bit_hi) of key
void radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                      size_t n, int nranges, const int* bit_lo, const int* bit_hi,
                      cudaStream_t s, long long* launches);
