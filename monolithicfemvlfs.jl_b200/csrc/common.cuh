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

struct CudaError { cudaError_t e; const char* file; int line; };

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
  int smem_off;     // offset (doubles) of the [nq][ncomp][ndofs] array
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
  int smem_geom_off[VLFS_MAX_SLOTS];   // per slot: pinv [nq][D][dref], normal [nq][D]
  int smem_meas_off;                   // [nq] measure*qweight
  int smem_w_off;                      // [nweights][nq]
  int smem_total;                      // doubles
};

// ---- primitives -------------------------------------------------------------------
void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_dev,
                        cudaStream_t s);
// stable LSD radix sort of (key, value) pairs on the bit range [bit_lo, bit_hi) of key
void radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                      size_t n, int nranges, const int* bit_lo, const int* bit_hi,
                      cudaStream_t s, long long* launches);
