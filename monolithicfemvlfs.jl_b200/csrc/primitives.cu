// primitives.cu - device-wide exclusive scan and stable LSD radix sort (sm_100a).
//
// Both are atomic-free: the radix ranking uses warp-level __match_any_sync and
// warp-private counters, so the symbolic phase (pattern build) is deterministic.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan of one tile held as SCAN_ITEMS consecutive items per thread;
// returns the tile total in every thread
__device__ uint32_t block_excl_scan(uint32_t (&item)[SCAN_ITEMS], uint32_t* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) tsum += item[k];
  uint32_t incl = warp_incl_scan(tsum, lane);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? smem[lane] : 0;
    uint32_t wi = warp_incl_scan(w, lane);
    if (lane < SCAN_THREADS / 32) smem[lane] = wi - w;
    if (lane == SCAN_THREADS / 32 - 1) smem[32] = wi;
  }
  __syncthreads();
  uint32_t run = smem[warp] + incl - tsum;
  uint32_t total = smem[32];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    uint32_t v = item[k];
    item[k] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums(const uint32_t* __restrict__ in, uint32_t* __restrict__ sums, size_t n) {
  __shared__ uint32_t smem[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t item[SCAN_ITEMS];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) item[k] = base + k < n ? in[base + k] : 0u;
  uint32_t total = block_excl_scan(item, smem);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_final(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                const uint32_t* __restrict__ offs, size_t n) {
  __shared__ uint32_t smem[33];
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t item[SCAN_ITEMS];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) item[k] = base + k < n ? in[base + k] : 0u;
  block_excl_scan(item, smem);
  uint32_t off = offs ? offs[blockIdx.x] : 0u;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) out[base + k] = item[k] + off;
}

__global__ void scan_single_total(const uint32_t* in, const uint32_t* ex, size_t n, uint32_t* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total = n ? ex[n - 1] + in[n - 1] : 0u;
}

void scan_rec(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t s) {
  if (n == 0) return;
  size_t nt = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nt == 1) {
    scan_tile_final<<<1, SCAN_THREADS, 0, s>>>(in, out, nullptr, n);
    return;
  }
  DevBuf<uint32_t> sums, offs;
  sums.alloc(nt);
  offs.alloc(nt);
  scan_tile_sums<<<(unsigned)nt, SCAN_THREADS, 0, s>>>(in, sums.p, n);
  scan_rec(sums.p, offs.p, nt, s);
  scan_tile_final<<<(unsigned)nt, SCAN_THREADS, 0, s>>>(in, out, offs.p, n);
  CUDA_TRY(cudaStreamSynchronize(s));   // temporaries die here
}

// ---- radix sort ---------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int RS_BINS = 256;

// per-warp digit histogram of the warp's chunk; keys kept in registers
__device__ __forceinline__ void warp_histogram(const uint64_t* __restrict__ keys, size_t n,
                                               size_t wbase, int shift, int nbits, int lane,
                                               uint32_t* whist, uint64_t (&k)[RS_ITEMS],
                                               uint32_t (&dg)[RS_ITEMS]) {
  const uint32_t mask = (1u << nbits) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    size_t idx = wbase + (size_t)r * 32 + lane;
    bool valid = idx < n;
    k[r] = valid ? keys[idx] : 0ull;
    uint32_t d = valid ? (uint32_t)((k[r] >> shift) & mask) : (uint32_t)(RS_BINS + lane);
    dg[r] = d;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (valid && lane == (__ffs(peers) - 1)) whist[d] += __popc(peers);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_histogram(const uint64_t* __restrict__ keys, size_t n, int shift, int nbits,
             uint32_t* __restrict__ block_hist, unsigned nblocks) {
  __shared__ uint32_t whist[RS_WARPS][RS_BINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&whist[0][0])[i] = 0;
  __syncthreads();
  size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * RS_ITEMS * 32;
  uint64_t k[RS_ITEMS];
  uint32_t dg[RS_ITEMS];
  warp_histogram(keys, n, wbase, shift, nbits, lane, whist[warp], k, dg);
  __syncthreads();
  for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) t += whist[w][b];
    block_hist[(size_t)b * nblocks + blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
           uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
           int nbits, const uint32_t* __restrict__ block_off, unsigned nblocks) {
  __shared__ uint32_t whist[RS_WARPS][RS_BINS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&whist[0][0])[i] = 0;
  __syncthreads();
  size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)warp * RS_ITEMS * 32;
  uint64_t k[RS_ITEMS];
  uint32_t dg[RS_ITEMS];
  warp_histogram(keys, n, wbase, shift, nbits, lane, whist[warp], k, dg);
  __syncthreads();
  // turn the per-warp counts into per-warp starting positions
  for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) {
    uint32_t run = block_off[(size_t)b * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      uint32_t c = whist[w][b];
      whist[w][b] = run;
      run += c;
    }
  }
  __syncthreads();
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    size_t idx = wbase + (size_t)r * 32 + lane;
    bool valid = idx < n;
    uint32_t d = dg[r];
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t pos = 0;
    if (valid) pos = whist[warp][d] + __popc(peers & lt);
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[pos] = k[r];
      vals_out[pos] = vals[idx];
    }
  }
}

}  // namespace

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_dev,
                        cudaStream_t s) {
  // `in` and `out` may alias only if total_dev is null
  DevBuf<uint32_t> tmp;
  const uint32_t* src = in;
  if (in == out && total_dev) {
    tmp.alloc(n);
    CUDA_TRY(cudaMemcpyAsync(tmp.p, in, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    src = tmp.p;
  }
  scan_rec(src, out, n, s);
  if (total_dev) scan_single_total<<<1, 32, 0, s>>>(src, out, n, total_dev);
  CUDA_TRY(cudaStreamSynchronize(s));
}

void radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                      size_t n, int nranges, const int* bit_lo, const int* bit_hi,
                      cudaStream_t s, long long* launches) {
  if (n == 0) return;
  unsigned nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE);
  DevBuf<uint32_t> hist, off;
  hist.alloc((size_t)RS_BINS * nblocks);
  off.alloc((size_t)RS_BINS * nblocks);
  uint64_t* kin = keys; uint32_t* vin = vals;
  uint64_t* kout = keys_tmp; uint32_t* vout = vals_tmp;
  for (int r = 0; r < nranges; ++r) {
    for (int lo = bit_lo[r]; lo < bit_hi[r]; lo += 8) {
      int nbits = bit_hi[r] - lo < 8 ? bit_hi[r] - lo : 8;
      rs_histogram<<<nblocks, RS_THREADS, 0, s>>>(kin, n, lo, nbits, hist.p, nblocks);
      scan_rec(hist.p, off.p, (size_t)RS_BINS * nblocks, s);
      rs_scatter<<<nblocks, RS_THREADS, 0, s>>>(kin, vin, kout, vout, n, lo, nbits, off.p, nblocks);
      if (launches) *launches += 4;
      uint64_t* tk = kin; kin = kout; kout = tk;
      uint32_t* tv = vin; vin = vout; vout = tv;
    }
  }
  if (kin != keys) {
    CUDA_TRY(cudaMemcpyAsync(keys, kin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(vals, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
}
