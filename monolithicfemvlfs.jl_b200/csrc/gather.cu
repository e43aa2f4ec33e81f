// gather.cu - numeric pass over a fixed pattern when every integration domain is class-tabulated
// (the element blocks of all cell classes sit in one contiguous table per matrix, see integrate.cu):
//
//   gather_chunks   slot-centric sum of the class-table entries of every CSR slot, record stream in
//                   slot order, one warp per chunk of the stream, no run pointers, no atomics;
//   expand_rows     CSR rows whose record streams are identical have identical values: the gather runs
//                   on ONE representative row per class of rows (the row templates) and this kernel
//                   copies the template of every row into the CSR value array (one coalesced write of
//                   every value, the templates are read through L2).
//
// What these kernels replace in the reference: the numeric phase of Gridap's SparseMatrixAssembler
// behind AffineFEOperator(a,l,X,Y) (src/Yago_freq_domain.jl:167, src/Khabakhpasheva_freq_domain.jl:240),
// i.e. the loop `for cell: for (i,j): V[n] = elmat[i,j]` followed by sparse(I,J,V), which sums
// duplicates in COO order.
#include "common.cuh"
#include <type_traits>
#include <cstdlib>

namespace {

struct MatPtrs { double* mat[4]; };

// ---- chunked slot-centric pass ----------------------------------------------------------------------
// The records (class-table index of every COO entry) are stored in CSR-slot order, so the records of
// a range of consecutive slots are one contiguous piece of the stream.  A CHUNK is the set of slots
// whose run starts in records [qR, (q+1)R): at most VLFS_CHUNK_CAP records, a few hundred slots.
// One warp owns a chunk:
//   1. record-centric: the lanes stream the chunk's records with coalesced loads (all CAP/32 loads
//      of a lane issued back to back), fetch the class-table values (CAP/32 independent L2 gathers
//      per lane in flight, no idle lanes, no run pointers) and park them in shared memory;
//      bit 31 of a record marks the first record of a slot, so a ballot + popc gives every head its
//      slot and the run boundaries go to a small shared array;
//   2. slot-centric: lane k sums the run of slot k from shared memory in COO order (the order in
//      which Julia's sparse(I,J,V) adds duplicates: deterministic, no atomics) and the warp writes
//      the chunk's values with coalesced 16-byte stores - every value is written exactly once
//      (this is also the zero fill), every record is read exactly once.
// HBM traffic = nslots*sT + 4*nrec + 16 bytes per chunk.
#define VLFS_CHUNK_CAP 256
#define VLFS_REC_NONE 0x7fffffffu
struct ChunkMeta { uint32_t slot0, rec0, nslots, nrec; };      // 16 bytes, stored in processing order

inline unsigned grid_for(long long n, int threads = 256) {
  long long g = (n + threads - 1) / threads;
  return (unsigned)(g < 148 * 32 ? (g ? g : 1) : 148 * 32);
}

// longest run of COO entries of one slot
__global__ void chunk_max_run(const uint32_t* __restrict__ run_start, long long nnz, unsigned* __restrict__ maxrun) {
  unsigned mx = 0;
  for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < nnz;
       slot += (long long)gridDim.x * blockDim.x)
    mx = max(mx, run_start[slot + 1] - run_start[slot]);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(maxrun, mx);      // (a maximum: order independent; set-up only)
}

// records -> chunk format: bits 0..30 class-table index (VLFS_REC_NONE: the block has no table),
// bit 31 set on the first record of every slot
__global__ void chunk_mark_heads(const uint32_t* __restrict__ run_start, long long nnz, uint32_t* __restrict__ rec) {
  for (long long slot = blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < nnz;
       slot += (long long)gridDim.x * blockDim.x) {
    const uint32_t p0 = run_start[slot], p1 = run_start[slot + 1];
    for (uint32_t p = p0; p < p1; ++p) {
      uint32_t v = rec[p];
      if (v == 0xffffffffu) v = VLFS_REC_NONE;
      rec[p] = v | (p == p0 ? 0x80000000u : 0u);
    }
  }
}

// first slot of every chunk of the slots [slot_lo, slot_hi) (every window of R records holds a run
// start because no run is longer than R)
__global__ void chunk_first_slots(const uint32_t* __restrict__ run_start, long long slot_lo, long long slot_hi,
                                  uint32_t rbase, uint32_t R, long long nchunks, uint32_t* __restrict__ cslot) {
  for (long long slot = slot_lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; slot < slot_hi;
       slot += (long long)gridDim.x * blockDim.x) {
    const uint32_t q = (run_start[slot] - rbase) / R;
    if (slot == slot_lo || (run_start[slot - 1] - rbase) / R != q) cslot[q] = (uint32_t)slot;
    if (slot == slot_hi - 1) cslot[nchunks] = (uint32_t)slot_hi;
  }
}

// early[q] = 1 when no record of the chunk points into a class table that comes out of the
// quadrature kernel (those tables are ready later than the directly tabulated ones); warp per chunk
__global__ void chunk_flags(const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ cslot, long long nchunks,
                            const uint32_t* __restrict__ rec, const long long* __restrict__ late_lo,
                            const long long* __restrict__ late_hi, int nlate, uint32_t* __restrict__ early) {
  const int lane = threadIdx.x & 31;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < nchunks; q += nw) {
    const uint32_t p0 = run_start[cslot[q]], p1 = run_start[cslot[q + 1]];
    bool late = false;
    for (uint32_t p = p0 + lane; p < p1; p += 32) {
      const long long r = (long long)(rec[p] & VLFS_REC_NONE);
      for (int k = 0; k < nlate; ++k) late |= r >= late_lo[k] && r < late_hi[k];
    }
    late = __any_sync(0xffffffffu, late);
    if (lane == 0) early[q] = late ? 0u : 1u;
  }
}

// chunk metadata in processing order: the early chunks (in slot order), then the late ones
__global__ void chunk_emit(const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ cslot, long long nchunks,
                           const uint32_t* __restrict__ early, const uint32_t* __restrict__ epos,
                           const uint32_t* __restrict__ n_early, ChunkMeta* __restrict__ meta) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nchunks; q += (long long)gridDim.x * blockDim.x) {
    const uint32_t s0 = cslot[q], s1 = cslot[q + 1], p0 = run_start[s0], p1 = run_start[s1];
    const uint32_t pos = early[q] ? epos[q] : *n_early + ((uint32_t)q - epos[q]);
    meta[pos] = ChunkMeta{s0, p0, s1 - s0, p1 - p0};
  }
}

template <int NMAT, bool CPLX, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
gather_chunks(const uint4* __restrict__ meta, long long nchunks, const uint32_t* __restrict__ rec, MatPtrs ctab_all,
              MatPtrs ptrs) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  constexpr int J = VLFS_CHUNK_CAP / 32;
  __shared__ __align__(16) T sval[WARPS][VLFS_CHUNK_CAP];
  __shared__ unsigned short sstart[WARPS][VLFS_CHUNK_CAP + 2];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long c = (long long)blockIdx.x * WARPS + w;
  if (c >= nchunks) return;
  const uint4 m = __ldg(meta + c);                 // slot0, rec0, nslots, nrec
  uint32_t r[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const uint32_t i = j * 32 + lane;
    r[j] = i < m.w ? __ldcs(rec + m.y + i) : VLFS_REC_NONE;
  }
  unsigned short* st = sstart[w];
  {
    uint32_t base = 0;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const uint32_t hm = __ballot_sync(0xffffffffu, r[j] >> 31);
      if (r[j] >> 31) st[base + __popc(hm & lt)] = (unsigned short)(j * 32 + lane);
      base += __popc(hm);
    }
    if (lane == 0) st[m.z] = (unsigned short)m.w;
  }
  T* sv = sval[w];
#pragma unroll
  for (int mm = 0; mm < NMAT; ++mm) {
    const T* __restrict__ tab = reinterpret_cast<const T*>(ctab_all.mat[mm]);
    T v[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const uint32_t idx = r[j] & VLFS_REC_NONE;
      v[j] = idx != VLFS_REC_NONE ? __ldg(tab + idx) : T{};
    }
#pragma unroll
    for (int j = 0; j < J; ++j) sv[j * 32 + lane] = v[j];
    __syncwarp();
    T* __restrict__ out = reinterpret_cast<T*>(ptrs.mat[mm]) + m.x;
    for (uint32_t k = lane; k < m.z; k += 32) {
      const uint32_t a = st[k], b = st[k + 1];
      T acc = sv[a];
      for (uint32_t i = a + 1; i < b; ++i) {
        if constexpr (CPLX) { const T t = sv[i]; acc.x += t.x; acc.y += t.y; }
        else acc += sv[i];
      }
      out[k] = acc;
    }
    __syncwarp();
  }
}

// ---- row templates ------------------------------------------------------------------------------------
// Signature of a CSR row = hash of its piece of the record stream (head bits included) and its
// length.  Rows with equal signatures are checked word by word against the first row of their class.

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

// warp per row
__global__ void row_signatures(const int* __restrict__ rowptr, const uint32_t* __restrict__ run_start, long long nrows,
                               const uint32_t* __restrict__ rec, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += nw) {
    const int s0 = rowptr[row], s1 = rowptr[row + 1];
    const uint32_t p0 = run_start[s0], p1 = run_start[s1];
    uint64_t h = 0;
    for (uint32_t p = p0 + lane; p < p1; p += 32)
      h += mix64(((uint64_t)(p - p0) << 32 | rec[p]) + 0x9e3779b97f4a7c15ull);
#pragma unroll
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    h = mix64(h ^ ((uint64_t)(uint32_t)(s1 - s0) << 32 | (p1 - p0)));
    if (lane == 0) { key[row] = h; idx[row] = (uint32_t)row; }
  }
}

__global__ void flag_key_heads(const uint64_t* __restrict__ key, long long n, uint32_t* __restrict__ head) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
    head[q] = (q == 0 || key[q] != key[q - 1]) ? 1u : 0u;
}

// sorted position -> class of the row and representative of the class (first row of the class in
// the stable sorted order = smallest row index)
__global__ void assign_row_classes(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ head,
                                   const uint32_t* __restrict__ hpos, long long n, uint32_t* __restrict__ row_class,
                                   uint32_t* __restrict__ rep) {
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
    const uint32_t c = hpos[q] + head[q] - 1u;        // inclusive count of heads - 1
    row_class[idx[q]] = c;
    if (head[q]) rep[c] = idx[q];
  }
}

// word-by-word check of every row against its representative (warp per row); also the class's
// late flag (a record of the representative points into a table that is ready late)
__global__ void verify_row_classes(const int* __restrict__ rowptr, const uint32_t* __restrict__ run_start, long long nrows,
                                   const uint32_t* __restrict__ rec, const uint32_t* __restrict__ row_class,
                                   const uint32_t* __restrict__ rep, const long long* __restrict__ late_lo,
                                   const long long* __restrict__ late_hi, int nlate, uint32_t* __restrict__ class_early,
                                   int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < nrows; row += nw) {
    const uint32_t c = row_class[row];
    const long long r = rep[c];
    const int s0 = rowptr[row], s1 = rowptr[row + 1], t0 = rowptr[r], t1 = rowptr[r + 1];
    const uint32_t p0 = run_start[s0], p1 = run_start[s1], q0 = run_start[t0], q1 = run_start[t1];
    bool differ = (s1 - s0 != t1 - t0) || (p1 - p0 != q1 - q0);
    if (!differ && r != row)
      for (uint32_t p = lane; p < p1 - p0; p += 32) differ |= rec[p0 + p] != rec[q0 + p];
    if (__any_sync(0xffffffffu, differ)) { if (lane == 0) atomicExch(bad, 1); }
    if (r == row) {
      bool late = false;
      for (uint32_t p = p0 + lane; p < p1; p += 32) {
        const long long v = (long long)(rec[p] & VLFS_REC_NONE);
        for (int k = 0; k < nlate; ++k) late |= v >= late_lo[k] && v < late_hi[k];
      }
      late = __any_sync(0xffffffffu, late);
      if (lane == 0) class_early[c] = late ? 0u : 1u;
    }
  }
}

// classes renumbered: the early ones first (stable); sizes of the templates in the new order
__global__ void order_classes(const uint32_t* __restrict__ class_early, const uint32_t* __restrict__ epos,
                              const uint32_t* __restrict__ n_early, long long nclasses, const uint32_t* __restrict__ rep,
                              const int* __restrict__ rowptr, const uint32_t* __restrict__ run_start,
                              uint32_t* __restrict__ newid, uint32_t* __restrict__ rep_new, uint32_t* __restrict__ tslots,
                              uint32_t* __restrict__ trecs) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < nclasses; c += (long long)gridDim.x * blockDim.x) {
    const uint32_t n = class_early[c] ? epos[c] : *n_early + ((uint32_t)c - epos[c]);
    const uint32_t r = rep[c];
    newid[c] = n;
    rep_new[n] = r;
    tslots[n] = (uint32_t)(rowptr[r + 1] - rowptr[r]);
    trecs[n] = run_start[rowptr[r + 1]] - run_start[rowptr[r]];
  }
}

// template stream: the records of the representative rows, class after class (warp per class)
__global__ void build_template_stream(const uint32_t* __restrict__ rep_new, const uint32_t* __restrict__ toff,
                                      const uint32_t* __restrict__ troff, long long nclasses, const int* __restrict__ rowptr,
                                      const uint32_t* __restrict__ run_start, const uint32_t* __restrict__ rec,
                                      uint32_t* __restrict__ trec, uint32_t* __restrict__ trun, uint32_t T, uint32_t TR) {
  const int lane = threadIdx.x & 31;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < nclasses; c += nw) {
    const long long r = rep_new[c];
    const int s0 = rowptr[r], s1 = rowptr[r + 1];
    const uint32_t p0 = run_start[s0], p1 = run_start[s1], o = troff[c], so = toff[c];
    for (uint32_t p = lane; p < p1 - p0; p += 32) trec[o + p] = rec[p0 + p];
    for (int k = lane; k < s1 - s0; k += 32) trun[so + k] = o + (run_start[s0 + k] - p0);
    if (c == nclasses - 1 && lane == 0) trun[T] = TR;
  }
}

// rows in processing order (rows of early classes first, stable) with the expand metadata
// {first CSR slot, first template slot, length, 0}
__global__ void build_row_list(const uint32_t* __restrict__ row_class, const uint32_t* __restrict__ newid,
                               const uint32_t* __restrict__ toff, const uint32_t* __restrict__ row_early,
                               const uint32_t* __restrict__ rpos, const uint32_t* __restrict__ n_early, long long nrows,
                               const int* __restrict__ rowptr, uint4* __restrict__ list) {
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < nrows; row += (long long)gridDim.x * blockDim.x) {
    const uint32_t pos = row_early[row] ? rpos[row] : *n_early + ((uint32_t)row - rpos[row]);
    list[pos] = make_uint4((uint32_t)rowptr[row], toff[newid[row_class[row]]], (uint32_t)(rowptr[row + 1] - rowptr[row]), 0u);
  }
}

__global__ void row_early_flags(const uint32_t* __restrict__ row_class, const uint32_t* __restrict__ class_early,
                                long long nrows, uint32_t* __restrict__ row_early) {
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < nrows; row += (long long)gridDim.x * blockDim.x)
    row_early[row] = class_early[row_class[row]];
}

// A group of 32 list entries (consecutive CSR rows inside the early / late part of the list) is shared
// by SPLIT warps: every lane loads one entry (one coalesced 512-byte read), warp `sub` copies the rows
// sub, sub + SPLIT, ... of the group.  A row is one contiguous piece of the template array and of the
// value array: the lanes copy it 16 bytes each, up to U independent loads per lane in flight.
// Measured alternatives (Y1, late rows: 853 MB written, 777 MB of template reads through L2, 183 us =
// 72 % of the L2 slice throughput): a flat item -> row search over the group's prefix sums (8 shuffles
// per item, L1 data pipe 72 % busy, same time); the list ordered class after class with the template
// row kept in registers (L2 reads gone, but the scattered destination rows cost more DRAM write
// efficiency than the reads saved: 200 us).
template <int NMAT, bool CPLX, int U, int SPLIT>
__global__ void __launch_bounds__(256)
expand_rows(const uint4* __restrict__ list, long long nlist, MatPtrs tmpl, MatPtrs ptrs) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long first = (w / SPLIT) * 32;
  const int sub = (int)(w % SPLIT);
  if (first >= nlist) return;
  const int cnt = (int)min(32ll, nlist - first);
  const uint4 e = lane < cnt ? __ldcs(list + first + lane) : make_uint4(0u, 0u, 0u, 0u);
  // two rows per step: the loads of both rows are issued before the first store (twice the bytes in flight)
  for (int r = sub; r < cnt; r += 2 * SPLIT) {
    const int r2 = r + SPLIT < cnt ? r + SPLIT : r;
    const uint32_t dA = __shfl_sync(0xffffffffu, e.x, r), sA = __shfl_sync(0xffffffffu, e.y, r);
    const uint32_t lenA = __shfl_sync(0xffffffffu, e.z, r);
    const uint32_t dB = __shfl_sync(0xffffffffu, e.x, r2), sB = __shfl_sync(0xffffffffu, e.y, r2);
    const uint32_t lenB = r2 != r ? __shfl_sync(0xffffffffu, e.z, r2) : (__shfl_sync(0xffffffffu, e.z, r2), 0u);
#pragma unroll
    for (int mm = 0; mm < NMAT; ++mm) {
      const T* __restrict__ inA = reinterpret_cast<const T*>(tmpl.mat[mm]) + sA;
      const T* __restrict__ inB = reinterpret_cast<const T*>(tmpl.mat[mm]) + sB;
      T* __restrict__ outA = reinterpret_cast<T*>(ptrs.mat[mm]) + dA;
      T* __restrict__ outB = reinterpret_cast<T*>(ptrs.mat[mm]) + dB;
      T vA[U], vB[U];
#pragma unroll
      for (int u = 0; u < U; ++u) if (lane + 32 * u < lenA) vA[u] = __ldg(inA + lane + 32 * u);
#pragma unroll
      for (int u = 0; u < U; ++u) if (lane + 32 * u < lenB) vB[u] = __ldg(inB + lane + 32 * u);
#pragma unroll
      for (int u = 0; u < U; ++u) if (lane + 32 * u < lenA) __stcs(outA + lane + 32 * u, vA[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) if (lane + 32 * u < lenB) __stcs(outB + lane + 32 * u, vB[u]);
      for (uint32_t k = 32 * U + lane; k < lenA; k += 32) __stcs(outA + k, __ldg(inA + k));      // rows longer than 32 U
      for (uint32_t k = 32 * U + lane; k < lenB; k += 32) __stcs(outB + k, __ldg(inB + k));
    }
  }
}

}  // namespace

// longest run of one slot; 0 when the stream cannot take the chunk format
unsigned chunk_longest_run(const uint32_t* run_start, long long nnz, cudaStream_t s) {
  DevBuf<unsigned> mx;
  mx.alloc(1); mx.zero(s);
  chunk_max_run<<<grid_for(nnz), 256, 0, s>>>(run_start, nnz, mx.p);
  unsigned maxrun = 0;
  CUDA_TRY(cudaMemcpyAsync(&maxrun, mx.p, sizeof maxrun, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return (maxrun == 0 || maxrun > VLFS_CHUNK_CAP / 2) ? 0u : maxrun;
}

void launch_chunk_mark_heads(const uint32_t* run_start, long long nnz, uint32_t* rec, cudaStream_t s) {
  chunk_mark_heads<<<grid_for(nnz), 256, 0, s>>>(run_start, nnz, rec);
  CUDA_TRY(cudaGetLastError());
}

// chunk tables of the slots [slot_lo, slot_hi) of a record stream in chunk format
void build_chunks(const uint32_t* run_start, long long slot_lo, long long slot_hi, unsigned maxrun, const uint32_t* rec,
                  const long long* late_lo, const long long* late_hi, int nlate, DevBuf<unsigned char>& meta_out,
                  long long* nchunks_out, long long* nearly_out, cudaStream_t s, long long* launches) {
  *nchunks_out = *nearly_out = 0;
  if (slot_hi <= slot_lo) { meta_out.release(); return; }
  uint32_t rb[2];
  CUDA_TRY(cudaMemcpyAsync(&rb[0], run_start + slot_lo, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&rb[1], run_start + slot_hi, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  const uint32_t R = VLFS_CHUNK_CAP - maxrun + 1;
  const long long nchunks = ((long long)(rb[1] - rb[0]) + R - 1) / R;
  DevBuf<uint32_t> cslot, early, epos, total;
  cslot.alloc((size_t)nchunks + 1); early.alloc((size_t)nchunks); epos.alloc((size_t)nchunks); total.alloc(1);
  chunk_first_slots<<<grid_for(slot_hi - slot_lo), 256, 0, s>>>(run_start, slot_lo, slot_hi, rb[0], R, nchunks, cslot.p);
  chunk_flags<<<grid_for(nchunks * 32), 256, 0, s>>>(run_start, cslot.p, nchunks, rec, late_lo, late_hi, nlate, early.p);
  exclusive_scan_u32(early.p, epos.p, (size_t)nchunks, total.p, s);
  meta_out.alloc((size_t)nchunks * sizeof(ChunkMeta));
  chunk_emit<<<grid_for(nchunks), 256, 0, s>>>(run_start, cslot.p, nchunks, early.p, epos.p, total.p,
                                                reinterpret_cast<ChunkMeta*>(meta_out.p));
  CUDA_TRY(cudaGetLastError());
  unsigned ne = 0;
  CUDA_TRY(cudaMemcpyAsync(&ne, total.p, sizeof ne, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (launches) *launches += 5;
  *nchunks_out = nchunks;
  *nearly_out = (long long)ne;
}

void launch_gather_chunks(const void* meta, long long first, long long nchunks, const uint32_t* rec,
                          double* const* ctab_all, double* const* mats, int nmat, int is_complex, cudaStream_t s) {
  if (nchunks == 0) return;
  MatPtrs P{}, CT{};
  for (int m = 0; m < nmat && m < 4; ++m) { P.mat[m] = mats[m]; CT.mat[m] = ctab_all[m]; }
  static const int warps_env = getenv("VLFS_CHUNK_WARPS") ? atoi(getenv("VLFS_CHUNK_WARPS")) : 8;
  const uint4* mp = reinterpret_cast<const uint4*>(meta) + first;
  auto go = [&](auto k4, auto k8) {
    if (warps_env == 8) k8<<<(unsigned)((nchunks + 7) / 8), 256, 0, s>>>(mp, nchunks, rec, CT, P);
    else k4<<<(unsigned)((nchunks + 3) / 4), 128, 0, s>>>(mp, nchunks, rec, CT, P);
  };
#define VLFS_CH(NM, CX) go(gather_chunks<NM, CX, 4>, gather_chunks<NM, CX, 8>)
  switch (nmat * 2 + (is_complex ? 1 : 0)) {
    case 2: VLFS_CH(1, false); break;
    case 3: VLFS_CH(1, true); break;
    case 4: VLFS_CH(2, false); break;
    case 5: VLFS_CH(2, true); break;
    case 6: VLFS_CH(3, false); break;
    case 7: VLFS_CH(3, true); break;
    case 8: VLFS_CH(4, false); break;
    default: VLFS_CH(4, true); break;
  }
#undef VLFS_CH
  CUDA_TRY(cudaGetLastError());
}

// Row templates of a record stream in chunk format.  On success (true): `trec`/`trun` hold the
// template stream (T slots, early classes first: slots [0, T_early)), `list` the rows in processing
// order with their expand metadata (n_rows_early first).  false: rows are (nearly) all different
// or a signature collided - the caller gathers the whole stream.
bool build_row_templates(const int* rowptr, const uint32_t* run_start, long long nrows, long long nnz, const uint32_t* rec,
                         const long long* late_lo, const long long* late_hi, int nlate, DevBuf<uint32_t>& trec,
                         DevBuf<uint32_t>& trun, long long* T_out, long long* T_early_out, DevBuf<unsigned char>& list,
                         long long* rows_early_out, long long* nclasses_out, cudaStream_t s, long long* launches) {
  if (nrows == 0 || nrows >= (1ll << 32)) return false;
  DevBuf<uint64_t> key, key_tmp;
  DevBuf<uint32_t> idx, idx_tmp, head, hpos, total;
  key.alloc((size_t)nrows); key_tmp.alloc((size_t)nrows); idx.alloc((size_t)nrows); idx_tmp.alloc((size_t)nrows);
  head.alloc((size_t)nrows); hpos.alloc((size_t)nrows); total.alloc(1);
  row_signatures<<<grid_for(nrows * 32), 256, 0, s>>>(rowptr, run_start, nrows, rec, key.p, idx.p);
  int lo[1] = {0}, hi[1] = {64};
  radix_sort_pairs(key.p, idx.p, key_tmp.p, idx_tmp.p, (size_t)nrows, 1, lo, hi, s, launches);
  key_tmp.release(); idx_tmp.release();
  flag_key_heads<<<grid_for(nrows), 256, 0, s>>>(key.p, nrows, head.p);
  exclusive_scan_u32(head.p, hpos.p, (size_t)nrows, total.p, s);
  unsigned ncls = 0;
  CUDA_TRY(cudaMemcpyAsync(&ncls, total.p, sizeof ncls, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (launches) *launches += 3;
  if (ncls == 0 || (long long)ncls * 3 > nrows) return false;          // fewer than 3 rows per class: not worth it
  DevBuf<uint32_t> row_class, rep, class_early, epos;
  DevBuf<int> bad;
  row_class.alloc((size_t)nrows); rep.alloc(ncls); class_early.alloc(ncls); epos.alloc(ncls);
  bad.alloc(1); bad.zero(s);
  assign_row_classes<<<grid_for(nrows), 256, 0, s>>>(idx.p, head.p, hpos.p, nrows, row_class.p, rep.p);
  verify_row_classes<<<grid_for(nrows * 32), 256, 0, s>>>(rowptr, run_start, nrows, rec, row_class.p, rep.p, late_lo, late_hi,
                                                           nlate, class_early.p, bad.p);
  int hbad = 0;
  CUDA_TRY(cudaMemcpyAsync(&hbad, bad.p, sizeof hbad, cudaMemcpyDeviceToHost, s));
  exclusive_scan_u32(class_early.p, epos.p, ncls, total.p, s);
  unsigned ncls_early = 0;
  CUDA_TRY(cudaMemcpyAsync(&ncls_early, total.p, sizeof ncls_early, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (launches) *launches += 3;
  if (hbad) return false;                                              // signature collision
  key.release(); idx.release(); head.release(); hpos.release();
  DevBuf<uint32_t> n_early_dev, newid, rep_new, tslots, trecs, toff, troff;
  n_early_dev.alloc(1);
  CUDA_TRY(cudaMemcpyAsync(n_early_dev.p, total.p, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  newid.alloc(ncls); rep_new.alloc(ncls); tslots.alloc((size_t)ncls + 1); trecs.alloc((size_t)ncls + 1);
  toff.alloc((size_t)ncls + 1); troff.alloc((size_t)ncls + 1);
  CUDA_TRY(cudaMemsetAsync(tslots.p + ncls, 0, sizeof(uint32_t), s));
  CUDA_TRY(cudaMemsetAsync(trecs.p + ncls, 0, sizeof(uint32_t), s));
  order_classes<<<grid_for(ncls), 256, 0, s>>>(class_early.p, epos.p, n_early_dev.p, ncls, rep.p, rowptr, run_start, newid.p,
                                               rep_new.p, tslots.p, trecs.p);
  // (one extra zero item so that toff[ncls] / troff[ncls] are the totals and toff[ncls_early] is the early size)
  exclusive_scan_u32(tslots.p, toff.p, (size_t)ncls + 1, nullptr, s);
  exclusive_scan_u32(trecs.p, troff.p, (size_t)ncls + 1, nullptr, s);
  uint32_t T = 0, TR = 0, Te = 0;
  CUDA_TRY(cudaMemcpyAsync(&T, toff.p + ncls, sizeof T, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&TR, troff.p + ncls, sizeof TR, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&Te, toff.p + ncls_early, sizeof Te, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (launches) *launches += 3;
  if ((long long)T * 3 > nnz) return false;
  trec.alloc(TR); trun.alloc((size_t)T + 1);
  build_template_stream<<<grid_for((long long)ncls * 32), 256, 0, s>>>(rep_new.p, toff.p, troff.p, ncls, rowptr, run_start, rec,
                                                                       trec.p, trun.p, T, TR);
  DevBuf<uint32_t> row_early, rpos;
  row_early.alloc((size_t)nrows); rpos.alloc((size_t)nrows);
  row_early_flags<<<grid_for(nrows), 256, 0, s>>>(row_class.p, class_early.p, nrows, row_early.p);
  exclusive_scan_u32(row_early.p, rpos.p, (size_t)nrows, total.p, s);
  unsigned nre = 0;
  CUDA_TRY(cudaMemcpyAsync(&nre, total.p, sizeof nre, cudaMemcpyDeviceToHost, s));
  list.alloc((size_t)nrows * sizeof(uint4));
  build_row_list<<<grid_for(nrows), 256, 0, s>>>(row_class.p, newid.p, toff.p, row_early.p, rpos.p, total.p, nrows, rowptr,
                                                  reinterpret_cast<uint4*>(list.p));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s));
  if (launches) *launches += 4;
  *T_out = T; *T_early_out = Te; *rows_early_out = nre; *nclasses_out = ncls;
  return true;
}

void launch_expand_rows(const void* list, long long first, long long nlist, double* const* tmpl, double* const* mats, int nmat,
                        int is_complex, cudaStream_t s) {
  if (nlist == 0) return;
  MatPtrs P{}, TP{};
  for (int m = 0; m < nmat && m < 4; ++m) { P.mat[m] = mats[m]; TP.mat[m] = tmpl[m]; }
  const uint4* lp = reinterpret_cast<const uint4*>(list) + first;
  constexpr int SPLIT = 4;                                          // 8 warps = 2 groups of 32 rows per CTA
  const unsigned grid = (unsigned)((nlist + 63) / 64);
  auto go = [&](auto kern) { kern<<<grid, 256, 0, s>>>(lp, nlist, TP, P); };
  switch (nmat * 2 + (is_complex ? 1 : 0)) {
    case 2: go(expand_rows<1, false, 4, SPLIT>); break;
    case 3: go(expand_rows<1, true, 4, SPLIT>); break;
    case 4: go(expand_rows<2, false, 4, SPLIT>); break;
    case 5: go(expand_rows<2, true, 4, SPLIT>); break;
    case 6: go(expand_rows<3, false, 4, SPLIT>); break;
    case 7: go(expand_rows<3, true, 4, SPLIT>); break;
    case 8: go(expand_rows<4, false, 4, SPLIT>); break;
    default: go(expand_rows<4, true, 4, SPLIT>); break;
  }
  CUDA_TRY(cudaGetLastError());
}
