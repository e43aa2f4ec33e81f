// pattern.cu - symbolic phase of SparseMatrixAssembler on the GPU (sm_100a).
//
// Replaces Gridap's symbolic loop + `sparse(I,J,V)` sort/merge behind
// AffineFEOperator (reference call sites src/Yago_freq_domain.jl:167,
// src/Khabakhpasheva_freq_domain.jl:240): every touched block entry of every cell
// becomes a 64-bit (row,col) key; a stable radix sort groups duplicates; the run
// index of each key is its CSR slot.  The COO->slot permutation `dest` is kept so the
// numeric phase is a pure scatter-add into a fixed pattern.  No atomics anywhere here.
#include "common.cuh"

namespace {

__global__ void gen_keys(DomainDev dom, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx,
                         unsigned long long coo_base) {
  const long long epc = dom.entries_per_cell;
  const long long total = dom.ncells * epc;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    long long c = g / epc;
    int e = (int)(g - c * epc);
    int b = 0;
    while (b + 1 < dom.nblocks && e >= dom.blk[b + 1].off) ++b;
    const BlockDev& B = dom.blk[b];
    int le = e - B.off;
    int i = le / B.nc, j = le - i * B.nc;
    uint32_t row = (uint32_t)dom.slot[B.ts].dofs[c * dom.slot[B.ts].ndofs + i];
    uint32_t col = (uint32_t)dom.slot[B.us].dofs[c * dom.slot[B.us].ndofs + j];
    keys[coo_base + g] = ((uint64_t)row << 32) | col;
    idx[coo_base + g] = (uint32_t)(coo_base + g);
  }
}

__global__ void flag_heads(const uint64_t* __restrict__ keys, uint32_t* __restrict__ flag, size_t n) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n;
       g += (size_t)gridDim.x * blockDim.x)
    flag[g] = (g == 0 || keys[g] != keys[g - 1]) ? 1u : 0u;
}

// slot[g] = (#heads at positions <= g) - 1 ; given exclusive scan ex and flag
__global__ void emit_pattern(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                             const uint32_t* __restrict__ flag, const uint32_t* __restrict__ ex,
                             int* __restrict__ colind, int* __restrict__ rowind,
                             int* __restrict__ dest, uint32_t* __restrict__ run_start, size_t n) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n;
       g += (size_t)gridDim.x * blockDim.x) {
    uint32_t f = flag[g];
    int slot = (int)(ex[g] + f) - 1;
    dest[perm[g]] = slot;
    if (f) {
      run_start[slot] = (uint32_t)g;
      uint64_t k = keys[g];
      colind[slot] = (int)(uint32_t)(k & 0xffffffffull);
      rowind[slot] = (int)(uint32_t)(k >> 32);
    }
  }
}

// rowptr[r] = first slot whose row >= r   (rowind ascending)
__global__ void build_rowptr(const int* __restrict__ rowind, long long nnz, int* __restrict__ rowptr,
                             long long nrows) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r <= nrows;
       r += (long long)gridDim.x * blockDim.x) {
    long long lo = 0, hi = nnz;
    while (lo < hi) {
      long long mid = (lo + hi) >> 1;
      if (rowind[mid] < r) lo = mid + 1; else hi = mid;
    }
    rowptr[r] = (int)lo;
  }
}

__global__ void gen_csc_keys(const int* __restrict__ rowind, const int* __restrict__ colind,
                             uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, long long nnz) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < nnz;
       g += (long long)gridDim.x * blockDim.x) {
    keys[g] = ((uint64_t)(uint32_t)colind[g] << 32) | (uint32_t)rowind[g];
    idx[g] = (uint32_t)g;
  }
}

__global__ void split_csc(const uint64_t* __restrict__ keys, int* __restrict__ rowval,
                          int* __restrict__ colsorted, long long nnz) {
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < nnz;
       g += (long long)gridDim.x * blockDim.x) {
    rowval[g] = (int)(uint32_t)(keys[g] & 0xffffffffull);
    colsorted[g] = (int)(uint32_t)(keys[g] >> 32);
  }
}

// dest in the order the tiled integration kernel consumes it: [cell][tile][4][4], -1 = padding
__global__ void retile_dest(DomainDev dom, const int* __restrict__ dest, int* __restrict__ tiled) {
  const int ntiles = dom.tile_base[dom.nblocks];
  const long long total = dom.ncells * (long long)ntiles * 16;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < total;
       g += (long long)gridDim.x * blockDim.x) {
    const long long cell = g / (ntiles * 16);
    const int r = (int)(g - cell * (ntiles * 16));
    const int tile = r >> 4, a = (r >> 2) & 3, e = r & 3;
    int b = 0;
    while (tile >= dom.tile_base[b + 1]) ++b;
    const BlockDev& B = dom.blk[b];
    const int lt = tile - dom.tile_base[b];
    const int ti = lt / dom.tiles_j[b], tj = lt - ti * dom.tiles_j[b];
    const int i = ti * 4 + a, j = tj * 4 + e;
    tiled[g] = (i < B.nr && j < B.nc) ? dest[cell * dom.entries_per_cell + B.off + i * B.nc + j]
                                      : VLFS_DEST_PAD;
  }
}

int bits_for(long long n) {
  int b = 1;
  while ((1ll << b) < n) ++b;
  return b;
}

inline unsigned grid_for(size_t n, int threads = 256) {
  size_t g = (n + threads - 1) / threads;
  size_t cap = 148ull * 32;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

// Builds rowptr/colind/rowind and the per-domain dest arrays.  dest_out[d] receives a
// pointer into one big allocation `dest_all` in COO order (domain, cell, block, i, j).
void build_pattern_device(const std::vector<DomainDev>& doms, long long ndofs,
                          DevBuf<int>& rowptr, DevBuf<int>& colind, DevBuf<int>& rowind,
                          DevBuf<int>& dest_all, std::vector<long long>& dest_off,
                          DevBuf<uint32_t>& perm_out, DevBuf<uint32_t>& run_start,
                          long long* nnz_out, long long* ncoo_out, cudaStream_t s,
                          long long* launches) {
  long long ncoo = 0;
  dest_off.clear();
  for (auto& d : doms) {
    dest_off.push_back(ncoo);
    ncoo += d.ncells * d.entries_per_cell;
  }
  if (ncoo >= (1ll << 32)) throw std::string("COO count exceeds 2^32 on one GPU");
  DevBuf<uint64_t> keys, keys_tmp;
  DevBuf<uint32_t> idx, idx_tmp, flag, ex, total;
  keys.alloc(ncoo); keys_tmp.alloc(ncoo); idx.alloc(ncoo); idx_tmp.alloc(ncoo);
  for (size_t d = 0; d < doms.size(); ++d) {
    long long n = doms[d].ncells * doms[d].entries_per_cell;
    if (n == 0) continue;
    gen_keys<<<grid_for(n), 256, 0, s>>>(doms[d], keys.p, idx.p, (unsigned long long)dest_off[d]);
    if (launches) ++*launches;
  }
  CUDA_TRY(cudaGetLastError());
  int nb = bits_for(ndofs);
  int lo[2] = {0, 32}, hi[2] = {nb, 32 + nb};
  radix_sort_pairs(keys.p, idx.p, keys_tmp.p, idx_tmp.p, (size_t)ncoo, 2, lo, hi, s, launches);
  keys_tmp.release(); idx_tmp.release();
  flag.alloc(ncoo); ex.alloc(ncoo); total.alloc(1);
  flag_heads<<<grid_for(ncoo), 256, 0, s>>>(keys.p, flag.p, (size_t)ncoo);
  exclusive_scan_u32(flag.p, ex.p, (size_t)ncoo, total.p, s);
  uint32_t nnz32 = 0;
  CUDA_TRY(cudaMemcpyAsync(&nnz32, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  long long nnz = nnz32;
  if (nnz >= (1ll << 31)) throw std::string("nnz exceeds int32 range on one GPU");
  colind.alloc(nnz); rowind.alloc(nnz); rowptr.alloc(ndofs + 1); dest_all.alloc(ncoo);
  run_start.alloc(nnz + 1);
  const uint32_t ncoo32 = (uint32_t)ncoo;
  CUDA_TRY(cudaMemcpyAsync(run_start.p + nnz, &ncoo32, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
  emit_pattern<<<grid_for(ncoo), 256, 0, s>>>(keys.p, idx.p, flag.p, ex.p, colind.p, rowind.p,
                                              dest_all.p, run_start.p, (size_t)ncoo);
  build_rowptr<<<grid_for(ndofs + 1), 256, 0, s>>>(rowind.p, nnz, rowptr.p, ndofs);
  if (launches) *launches += 4;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s));
  perm_out = std::move(idx);          // sorted position -> COO index (slot-centric gather)
  *nnz_out = nnz;
  *ncoo_out = ncoo;
}

void build_tiled_dest(const DomainDev& dom, const int* dest, DevBuf<int>& tiled, cudaStream_t s,
                      long long* launches) {
  const long long total = dom.ncells * (long long)dom.tile_base[dom.nblocks] * 16;
  tiled.alloc((size_t)total);
  if (total == 0) return;
  retile_dest<<<grid_for((size_t)total), 256, 0, s>>>(dom, dest, tiled.p);
  if (launches) ++*launches;
  CUDA_TRY(cudaGetLastError());
}

// CSC view of the CSR pattern: colptr, rowval and the CSR slot of every CSC entry
void build_csc_device(const int* rowind, const int* colind, long long nnz, long long ndofs,
                      DevBuf<int>& colptr, DevBuf<int>& rowval, DevBuf<int>& csr_slot,
                      cudaStream_t s, long long* launches) {
  DevBuf<uint64_t> keys, keys_tmp;
  DevBuf<uint32_t> idx, idx_tmp;
  DevBuf<int> colsorted;
  keys.alloc(nnz); keys_tmp.alloc(nnz); idx.alloc(nnz); idx_tmp.alloc(nnz);
  gen_csc_keys<<<grid_for(nnz), 256, 0, s>>>(rowind, colind, keys.p, idx.p, nnz);
  int nb = bits_for(ndofs);
  int lo[1] = {32}, hi[1] = {32 + nb};        // stable: rows stay ascending inside a column
  radix_sort_pairs(keys.p, idx.p, keys_tmp.p, idx_tmp.p, (size_t)nnz, 1, lo, hi, s, launches);
  rowval.alloc(nnz); colsorted.alloc(nnz); colptr.alloc(ndofs + 1); csr_slot.alloc(nnz);
  split_csc<<<grid_for(nnz), 256, 0, s>>>(keys.p, rowval.p, colsorted.p, nnz);
  build_rowptr<<<grid_for(ndofs + 1), 256, 0, s>>>(colsorted.p, nnz, colptr.p, ndofs);
  CUDA_TRY(cudaMemcpyAsync(csr_slot.p, idx.p, nnz * sizeof(int), cudaMemcpyDeviceToDevice, s));
  if (launches) *launches += 3;
  CUDA_TRY(cudaStreamSynchronize(s));
}
