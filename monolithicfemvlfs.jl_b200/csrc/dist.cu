// dist.cu - multi-GPU data plane of libvlfs_b200 (sm_100a, NCCL over NVLink 5 / NVSwitch).
//
// The reference is a serial Julia program; this file is what lets the same `AffineFEOperator` +
// `solve(op)` path (src/Yago_freq_domain.jl:167-171, src/Multi_geo_freq_domain.jl:131-135) run
// row-partitioned on 2/4/8 B200 behind the C ABI, with no host framework in the data path:
//   * the library owns the NCCL communicator (one rank per context; ranks are threads of one
//     process - ncclCommInitAll, what a Julia `ccall` host uses - or one process each);
//   * y = A x: pack kernel + grouped ncclSend/ncclRecv of the ghost values on a side stream,
//     overlapped with the rows that touch no ghost column; the rows that do follow;
//   * Krylov: every batch of dot products is ONE ncclAllReduce (krylov.cu: gmres);
//   * exact solve by substructuring on the chain of slab cuts: every rank eliminates its interior
//     with the multifrontal LU keeping its own cut and the next rank's cut as Schur boundary, then
//     factorises ITS link of the block-tridiagonal interface chain (own cut = pivots, next cut =
//     update set) after adding the corner rank-1 passes on, and passes its own corner to rank+1.
//     No rank ever holds the whole interface; an application is one forward relay (rank 0 -> P-1)
//     and one backward relay of cut-sized vectors between the local triangular sweeps.
// NCCL is bound at run time (dlopen) so that the single-GPU library has no dependency on it.
#include "dist.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <memory>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>

void launch_spmv_rows(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                      const int* rows, long long nrows, long long nnz, int is_complex, const double* alpha,
                      const double* beta, int nsm, cudaStream_t s);
void launch_spmv(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                 long long nrows, long long nnz, int is_complex, const double* alpha,
                 const double* beta, int nsm, cudaStream_t s);

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) return;
    auto sym = [&](const char* n) { return dlsym(api.handle, n); };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.Send || !api.Recv || !api.AllReduce)
    throw NcclError{"libnccl.so.2 not found or incomplete (multi-GPU entry points need NCCL)"};
  return api;
}

#define NCCL_TRY(x)                                                                   \
  do {                                                                                \
    ncclResult_t _r = (x);                                                            \
    if (_r != ncclSuccess) {                                                          \
      NcclApi& _a = nccl();                                                           \
      throw NcclError{std::string("NCCL: ") + (_a.GetErrorString ? _a.GetErrorString(_r) : "error") + \
                      " at " __FILE__ ":" + std::to_string(__LINE__)};               \
    }                                                                                 \
  } while (0)

// In-process transport for ranks that share ONE device (threads of one process, e.g. the 1-GPU test
// box, where NCCL refuses two ranks on a device): messages are staged device buffers handed over
// through a host mailbox and ordered by CUDA events; reductions add the ranks' parts in rank order.
// No kernel ever waits for another kernel; only host threads wait for each other.
struct Loopback {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  struct Msg { double* buf; size_t bytes; cudaEvent_t ev; };
  std::vector<std::deque<Msg>> q;                 // [src * n + dst]
  std::vector<double*> stage;                     // all-reduce staging, one per rank
  size_t stage_cap = 1 << 14;                     // doubles
  int arrived = 0, generation = 0;
  explicit Loopback(int n_) : n(n_), q((size_t)n_ * n_), stage(n_, nullptr) {}
  ~Loopback() { for (double* p : stage) if (p) cudaFree(p); }
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const int gen = generation;
    if (++arrived == n) { arrived = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&]() { return generation != gen; });
  }
};

__global__ void loopback_sum(double* __restrict__ dst, const double* const* __restrict__ parts, int nparts, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < nparts; ++r) s += parts[r][i];
    dst[i] = s;
  }
}

template <bool C> struct Val { using T = double; };
template <> struct Val<true> { using T = double2; };
__device__ __forceinline__ double vadd(double a, double b) { return a + b; }
__device__ __forceinline__ double2 vadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }

// dst[k] = src[idx[k]]
template <bool C>
__global__ void pack_entries(const int* __restrict__ idx, const typename Val<C>::T* __restrict__ src,
                             typename Val<C>::T* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
// dst[k] += src[idx[k]]   (idx == nullptr: dst[k] += src[k])
template <bool C>
__global__ void add_entries(const int* __restrict__ idx, const typename Val<C>::T* __restrict__ src,
                            typename Val<C>::T* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = vadd(dst[i], src[idx ? idx[i] : i]);
}
// flag[r] = 1 when row r (< n_owned) has a ghost column
__global__ void flag_boundary_rows(const int* __restrict__ rowptr, const int* __restrict__ colind, long long n_owned,
                                   uint32_t* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = w; r < n_owned; r += nw) {
    bool any = false;
    for (int p = rowptr[r] + lane; p < rowptr[r + 1]; p += 32) any |= colind[p] >= n_owned;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flag[r] = any ? 1u : 0u;
  }
}
// rows split by the flag: interior rows to lists[0..], boundary rows to lists[n_int..]; ex = exclusive
// scan of flag
__global__ void split_rows(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ ex, long long n_owned,
                           long long n_int, int* __restrict__ lists, const int* __restrict__ rowptr,
                           unsigned long long* __restrict__ nnz_bnd) {
  unsigned long long mine = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_owned; r += (long long)gridDim.x * blockDim.x) {
    if (flag[r]) { lists[n_int + ex[r]] = (int)r; mine += (unsigned long long)(rowptr[r + 1] - rowptr[r]); }
    else lists[r - ex[r]] = (int)r;
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(nnz_bnd, mine);     // (integer count: order-independent)
}
// F[ki][kj] += A[i][j] for kept rows i and kept columns j, except the (next cut, next cut) corner, which
// belongs to the next rank's link.  One warp per kept row.
template <bool C>
__global__ void add_kept_entries(const int* __restrict__ keep, long long nk, long long nL, const int* __restrict__ keepidx,
                                 const int* __restrict__ rowptr, const int* __restrict__ colind,
                                 const typename Val<C>::T* __restrict__ vals, typename Val<C>::T* __restrict__ F) {
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long ki = w; ki < nk; ki += nw) {
    const int i = keep[ki];
    for (int p = rowptr[i] + lane; p < rowptr[i + 1]; p += 32) {
      const int kj = keepidx[colind[p]];
      if (kj < 0 || (ki >= nL && kj >= nL)) continue;
      typename Val<C>::T* f = F + ki * nk + kj;
      *f = vadd(*f, vals[p]);
    }
  }
}
// strided nb x nb block copy / add:  mode 0: dst(contiguous nb x nb) = src block at (o,o) of an nf x nf
// matrix; mode 1: dst block at (0,0) of nf x nf += src (contiguous)
template <bool C>
__global__ void corner_op(typename Val<C>::T* __restrict__ M, long long nf, long long o, long long nb,
                          typename Val<C>::T* __restrict__ buf, int mode) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nb * nb; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / nb, j = t - i * nb;
    typename Val<C>::T* m = M + (o + i) * nf + o + j;
    if (mode == 0) buf[t] = *m; else *m = vadd(*m, buf[t]);
  }
}

inline unsigned vgrid(long long n, int nsm) {
  long long g = (n + 255) / 256, cap = (long long)nsm * 8;
  return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

struct DistState : DistHooks {
  int rank = 0, nranks = 1, device = 0;
  ncclComm_t comm = nullptr;
  std::shared_ptr<Loopback> loop;                   // in-process transport instead of NCCL (ranks on one device)
  DevBuf<const double*> loop_parts;
  cudaStream_t s = nullptr, cs = nullptr;           // main stream (the context's), halo stream
  cudaEvent_t ev_x = nullptr, ev_halo = nullptr;
  int cplx = 0, nsm = 148;
  size_t sc = 1;
  long long* launches = nullptr;
  long long n_owned = 0, nnz = 0;
  // halo plan
  std::vector<int> peers;
  std::vector<long long> send_ptr, recv_ptr;
  std::vector<int> send_idx_host;
  DevBuf<int> send_idx;
  DevBuf<double> send_buf;
  // SpMV split
  const int* rowptr = nullptr; const int* colind = nullptr;
  DevBuf<int> row_lists;
  long long n_int = 0, n_bnd = 0, nnz_bnd = 0;
  // substructuring
  bool schur_ready = false;
  DirectSolver* local = nullptr;
  DirectPtr front;
  long long nL = 0, nR = 0;
  int lower = -1, upper = -1;
  DevBuf<int> keep, keepidx;
  DevBuf<double> corner, g, xk, tbuf;
  double chain_ms = 0;

  ~DistState() override {
    if (cs) cudaStreamDestroy(cs);
    if (ev_x) cudaEventDestroy(ev_x);
    if (ev_halo) cudaEventDestroy(ev_halo);
    if (comm) { try { nccl().CommDestroy(comm); } catch (...) {} }
  }

  void send(const double* p, long long count, int peer, cudaStream_t st) {
    if (loop) {
      Loopback::Msg m{nullptr, (size_t)count * sc * sizeof(double), nullptr};
      CUDA_TRY(cudaMalloc(&m.buf, std::max<size_t>(m.bytes, 8)));
      CUDA_TRY(cudaMemcpyAsync(m.buf, p, m.bytes, cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaEventCreateWithFlags(&m.ev, cudaEventDisableTiming));
      CUDA_TRY(cudaEventRecord(m.ev, st));
      { std::lock_guard<std::mutex> lk(loop->mu); loop->q[(size_t)rank * loop->n + peer].push_back(m); }
      loop->cv.notify_all();
      return;
    }
    NCCL_TRY(nccl().Send(p, (size_t)count * sc, ncclDouble, peer, comm, st));
  }
  void recv(double* p, long long count, int peer, cudaStream_t st) {
    if (loop) {
      Loopback::Msg m;
      {
        std::unique_lock<std::mutex> lk(loop->mu);
        auto& qq = loop->q[(size_t)peer * loop->n + rank];
        loop->cv.wait(lk, [&]() { return !qq.empty(); });
        m = qq.front(); qq.pop_front();
      }
      if (m.bytes != (size_t)count * sc * sizeof(double)) throw std::string("loopback transport: message size mismatch");
      CUDA_TRY(cudaStreamWaitEvent(st, m.ev, 0));
      CUDA_TRY(cudaMemcpyAsync(p, m.buf, m.bytes, cudaMemcpyDeviceToDevice, st));
      CUDA_TRY(cudaStreamSynchronize(st));           // (test transport: simple lifetime of the staged buffer)
      cudaEventDestroy(m.ev);
      cudaFree(m.buf);
      return;
    }
    NCCL_TRY(nccl().Recv(p, (size_t)count * sc, ncclDouble, peer, comm, st));
  }
  void group_start() { if (!loop) NCCL_TRY(nccl().GroupStart()); }
  void group_end() { if (!loop) NCCL_TRY(nccl().GroupEnd()); }

  // ghost block of x <- owners, on stream st
  void exchange_on(double* x, cudaStream_t st) {
    if (peers.empty()) return;
    const long long nsend = send_ptr.back();
    if (nsend) {
      if (cplx) pack_entries<true><<<vgrid(nsend, nsm), 256, 0, st>>>(send_idx.p, (const double2*)x, (double2*)send_buf.p, nsend);
      else pack_entries<false><<<vgrid(nsend, nsm), 256, 0, st>>>(send_idx.p, x, send_buf.p, nsend);
      ++*launches;
    }
    group_start();                                   // (loopback: every send is posted before a receive waits)
    for (size_t p = 0; p < peers.size(); ++p) {
      const long long ns = send_ptr[p + 1] - send_ptr[p];
      if (ns) send(send_buf.p + (size_t)send_ptr[p] * sc, ns, peers[p], st);
    }
    for (size_t p = 0; p < peers.size(); ++p) {
      const long long nr = recv_ptr[p + 1] - recv_ptr[p];
      if (nr) recv(x + (size_t)(n_owned + recv_ptr[p]) * sc, nr, peers[p], st);
    }
    group_end();
  }

  void spmv(const double* vals, double* x, double* y, const double* alpha, const double* beta) override {
    if (peers.empty()) {
      launch_spmv(rowptr, colind, vals, x, y, n_owned, nnz, cplx, alpha, beta, nsm, s);
      return;
    }
    // halo on the side stream while the rows without ghost columns are multiplied
    CUDA_TRY(cudaEventRecord(ev_x, s));
    CUDA_TRY(cudaStreamWaitEvent(cs, ev_x, 0));
    exchange_on(x, cs);
    CUDA_TRY(cudaEventRecord(ev_halo, cs));
    launch_spmv_rows(rowptr, colind, vals, x, y, row_lists.p, n_int, nnz - nnz_bnd, cplx, alpha, beta, nsm, s);
    CUDA_TRY(cudaStreamWaitEvent(s, ev_halo, 0));
    launch_spmv_rows(rowptr, colind, vals, x, y, row_lists.p + n_int, n_bnd, nnz_bnd, cplx, alpha, beta, nsm, s);
  }

  void exchange(double* x) override { exchange_on(x, s); }

  void allreduce_sum(double* dev, int ndoubles) override {
    if (nranks == 1) return;
    if (loop) {
      if ((size_t)ndoubles > loop->stage_cap) throw std::string("loopback transport: reduction too long");
      CUDA_TRY(cudaMemcpyAsync(loop->stage[rank], dev, (size_t)ndoubles * sizeof(double), cudaMemcpyDeviceToDevice, s));
      CUDA_TRY(cudaStreamSynchronize(s));
      loop->barrier();
      loopback_sum<<<1, 256, 0, s>>>(dev, loop_parts.p, nranks, ndoubles);
      CUDA_TRY(cudaStreamSynchronize(s));
      loop->barrier();
      return;
    }
    NCCL_TRY(nccl().AllReduce(dev, dev, (size_t)ndoubles, ncclDouble, ncclSum, comm, s));
  }

  bool has_exact_apply() const override { return schur_ready; }

  // z = A^-1 r across all ranks
  void apply(const double* r, double* z) override {
    const long long nk = nL + nR;
    if (nk) CUDA_TRY(cudaMemsetAsync(g.p, 0, (size_t)nk * sc * sizeof(double), s));
    direct_forward(*local, r, nk ? g.p : nullptr);                 // g += -A_KE A_EE^-1 r_E
    if (nL) {                                                      // own cut: + r on the cut (+ relay of rank-1)
      if (cplx) add_entries<true><<<vgrid(nL, nsm), 256, 0, s>>>(keep.p, (const double2*)r, (double2*)g.p, nL);
      else add_entries<false><<<vgrid(nL, nsm), 256, 0, s>>>(keep.p, r, g.p, nL);
      ++*launches;
      if (lower >= 0) {
        recv(tbuf.p, nL, lower, s);
        if (cplx) add_entries<true><<<vgrid(nL, nsm), 256, 0, s>>>(nullptr, (const double2*)tbuf.p, (double2*)g.p, nL);
        else add_entries<false><<<vgrid(nL, nsm), 256, 0, s>>>(nullptr, tbuf.p, g.p, nL);
        ++*launches;
      }
      direct_forward(*front, g.p, nR ? g.p + (size_t)nL * sc : nullptr);
    }
    if (nR && upper >= 0) send(g.p + (size_t)nL * sc, nR, upper, s);
    // backward relay: the next cut's values come from rank+1
    if (nR && upper >= 0) recv(xk.p + (size_t)nL * sc, nR, upper, s);
    if (nL) {
      direct_backward(*front, nR ? xk.p + (size_t)nL * sc : nullptr, xk.p);
      if (lower >= 0) send(xk.p, nL, lower, s);
    }
    direct_backward(*local, nk ? xk.p : nullptr, z);
  }
};

void DistDeleter::operator()(DistState* p) const { delete p; }

void dist_unique_id(void* id128) {
  ncclUniqueId id;
  NCCL_TRY(nccl().GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  memcpy(id128, &id, sizeof id);
}

DistPtr dist_init_rank(int rank, int nranks, const void* id128, int device) {
  DistPtr D(new DistState());
  D->rank = rank; D->nranks = nranks; D->device = device;
  CUDA_TRY(cudaSetDevice(device));
  if (nranks > 1) {
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCCL_TRY(nccl().CommInitRank(&D->comm, nranks, id, rank));
  }
  return D;
}

void dist_init_all(int n, const int* devs, std::vector<DistPtr>& out) {
  out.clear();
  bool same_device = n > 1;
  for (int r = 1; r < n; ++r) same_device &= devs[r] == devs[0];
  if (same_device) {                                 // ranks share one GPU: in-process transport
    auto loop = std::make_shared<Loopback>(n);
    CUDA_TRY(cudaSetDevice(devs[0]));
    for (int r = 0; r < n; ++r) CUDA_TRY(cudaMalloc(&loop->stage[r], loop->stage_cap * sizeof(double)));
    for (int r = 0; r < n; ++r) {
      DistPtr D(new DistState());
      D->rank = r; D->nranks = n; D->device = devs[r]; D->loop = loop;
      std::vector<const double*> parts(loop->stage.begin(), loop->stage.end());
      D->loop_parts.alloc(n);
      CUDA_TRY(cudaMemcpy(D->loop_parts.p, parts.data(), n * sizeof(double*), cudaMemcpyHostToDevice));
      out.push_back(std::move(D));
    }
    return;
  }
  std::vector<ncclComm_t> comms(n, nullptr);
  if (n > 1) {
    if (!nccl().CommInitAll) throw NcclError{"ncclCommInitAll missing"};
    NCCL_TRY(nccl().CommInitAll(comms.data(), n, devs));
  }
  for (int r = 0; r < n; ++r) {
    DistPtr D(new DistState());
    D->rank = r; D->nranks = n; D->device = devs[r]; D->comm = comms[r];
    out.push_back(std::move(D));
  }
}

void dist_bind(DistState& D, cudaStream_t s, int is_complex, int nsm, long long* launches) {
  D.s = s; D.cplx = is_complex; D.sc = is_complex ? 2 : 1; D.nsm = nsm; D.launches = launches;
  if (!D.cs) {
    {
      // the halo stream outranks the product: its pack kernel and the NCCL send/recv kernels get their
      // CTAs as soon as one retires instead of queueing behind the interior rows
      int lo = 0, hi = 0;
      CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_TRY(cudaStreamCreateWithPriority(&D.cs, cudaStreamNonBlocking, hi));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&D.ev_x, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&D.ev_halo, cudaEventDisableTiming));
  }
}

int dist_rank(const DistState& D) { return D.rank; }
int dist_nranks(const DistState& D) { return D.nranks; }
bool dist_has_halo(const DistState& D) { return D.rowptr != nullptr; }
DistHooks* dist_hooks(DistState& D) { return &D; }
void dist_exchange(DistState& D, double* x_dev) { D.exchange_on(x_dev, D.s); }

void dist_set_halo(DistState& D, long long n_owned, long long n_local, int npeers, const int* peers,
                   const long long* send_ptr, const int* send_idx, const long long* recv_ptr,
                   const int* rowptr_dev, const int* colind_dev, long long nnz) {
  D.n_owned = n_owned; D.n_local = n_local; D.nnz = nnz;
  D.rowptr = rowptr_dev; D.colind = colind_dev;
  D.peers.assign(peers, peers + npeers);
  D.send_ptr.assign(send_ptr, send_ptr + npeers + 1);
  D.recv_ptr.assign(recv_ptr, recv_ptr + npeers + 1);
  if (D.send_ptr[0] != 0 || D.recv_ptr[0] != 0 || D.recv_ptr[npeers] != n_local - n_owned)
    throw std::string("halo plan: recv ranges must cover the ghost block exactly");
  const long long nsend = D.send_ptr[npeers];
  D.send_idx_host.assign(send_idx, send_idx + nsend);
  for (int v : D.send_idx_host) if (v < 0 || v >= n_owned) throw std::string("halo plan: send index is not an owned row");
  for (int p = 0; p < npeers; ++p)
    if (peers[p] < 0 || peers[p] >= D.nranks || peers[p] == D.rank) throw std::string("halo plan: bad peer rank");
  D.send_idx.upload(send_idx, (size_t)std::max<long long>(nsend, 1), D.s);
  D.send_buf.alloc((size_t)std::max<long long>(nsend, 1) * D.sc);
  D.schur_ready = false;
  // rows without / with ghost columns
  DevBuf<uint32_t> flag, ex, total;
  DevBuf<unsigned long long> nb;
  flag.alloc((size_t)n_owned); ex.alloc((size_t)n_owned); total.alloc(1); nb.alloc(1); nb.zero(D.s);
  flag_boundary_rows<<<vgrid(n_owned * 32, D.nsm), 256, 0, D.s>>>(rowptr_dev, colind_dev, n_owned, flag.p);
  exclusive_scan_u32(flag.p, ex.p, (size_t)n_owned, total.p, D.s);
  uint32_t nbnd = 0;
  CUDA_TRY(cudaMemcpyAsync(&nbnd, total.p, sizeof nbnd, cudaMemcpyDeviceToHost, D.s));
  CUDA_TRY(cudaStreamSynchronize(D.s));
  D.n_bnd = nbnd; D.n_int = n_owned - nbnd;
  D.row_lists.alloc((size_t)n_owned);
  split_rows<<<vgrid(n_owned, D.nsm), 256, 0, D.s>>>(flag.p, ex.p, n_owned, D.n_int, D.row_lists.p, rowptr_dev, nb.p);
  unsigned long long h = 0;
  CUDA_TRY(cudaMemcpyAsync(&h, nb.p, sizeof h, cudaMemcpyDeviceToHost, D.s));
  CUDA_TRY(cudaStreamSynchronize(D.s));
  D.nnz_bnd = (long long)h;
  *D.launches += 4;
  CUDA_TRY(cudaGetLastError());
}

void dist_roles(DistState& D, std::vector<signed char>& role) {
  const long long n_ghost = D.n_local - D.n_owned;
  role.assign((size_t)D.n_local, 0);
  D.lower = D.upper = -1;
  for (size_t p = 0; p < D.peers.size(); ++p) {
    const int peer = D.peers[p];
    const long long ns = D.send_ptr[p + 1] - D.send_ptr[p], nr = D.recv_ptr[p + 1] - D.recv_ptr[p];
    if (peer < D.rank) {
      if (ns) {                                                    // my cut: owned unknowns a lower rank sees
        if (peer != D.rank - 1) throw std::string("substructuring: the cuts do not form a chain (rank " + std::to_string(D.rank) +
                                                  " is seen by rank " + std::to_string(peer) + ")");
        D.lower = peer;
        for (long long k = D.send_ptr[p]; k < D.send_ptr[p + 1]; ++k) role[(size_t)D.send_idx_host[k]] = 1;
      }
      for (long long k = D.recv_ptr[p]; k < D.recv_ptr[p + 1]; ++k) role[(size_t)(D.n_owned + k)] = 2;
    } else {
      if (nr) {
        if (peer != D.rank + 1) throw std::string("substructuring: the cuts do not form a chain (rank " + std::to_string(D.rank) +
                                                  " sees rank " + std::to_string(peer) + ")");
        D.upper = peer;
        for (long long k = D.recv_ptr[p]; k < D.recv_ptr[p + 1]; ++k) role[(size_t)(D.n_owned + k)] = 1;
      }
    }
  }
  (void)n_ghost;
  std::vector<int> keep, keepidx((size_t)D.n_local, -1);
  for (long long i = 0; i < D.n_local; ++i) if (role[(size_t)i] == 1) { keepidx[(size_t)i] = (int)keep.size(); keep.push_back((int)i); }
  D.nL = 0;
  for (int i : keep) D.nL += i < D.n_owned;
  D.nR = (long long)keep.size() - D.nL;
  if (keep.empty()) keep.push_back(0);
  D.keep.upload(keep.data(), keep.size(), D.s);
  D.keepidx.upload(keepidx.data(), keepidx.size(), D.s);
  CUDA_TRY(cudaStreamSynchronize(D.s));
  D.schur_ready = false;
  D.front.reset();
}

void dist_schur_factor(DistState& D, DirectSolver& local, const int* rowptr, const int* colind, const double* vals) {
  const long long nk = D.nL + D.nR;
  if (direct_nkeep(local) != nk) throw std::string("substructuring: kept set of the local LU does not match the halo plan");
  D.local = &local;
  cudaStream_t s = D.s;
  if (!D.front) {
    D.front = direct_front(D.nL, nk, D.cplx, D.nsm, s, D.launches);
    D.g.alloc((size_t)std::max<long long>(nk, 1) * D.sc);
    D.xk.alloc((size_t)std::max<long long>(nk, 1) * D.sc);
    D.tbuf.alloc((size_t)std::max<long long>(D.nL, 1) * D.sc);
    D.corner.alloc((size_t)std::max<long long>(std::max(D.nL * D.nL, D.nR * D.nR), 1) * D.sc);
  }
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
  CUDA_TRY(cudaEventRecord(e0, s));
  double* F = direct_pool(*D.front);
  if (nk) {
    CUDA_TRY(cudaMemsetAsync(F, 0, (size_t)nk * nk * D.sc * sizeof(double), s));
    direct_schur(local, F);                                          // -A_KE A_EE^-1 A_EK
    const unsigned grid = vgrid(nk * 32, D.nsm);
    if (D.cplx) add_kept_entries<true><<<grid, 256, 0, s>>>(D.keep.p, nk, D.nL, D.keepidx.p, rowptr, colind, (const double2*)vals, (double2*)F);
    else add_kept_entries<false><<<grid, 256, 0, s>>>(D.keep.p, nk, D.nL, D.keepidx.p, rowptr, colind, vals, F);
    ++*D.launches;
  }
  if (D.nL && D.lower >= 0) {                                        // corner of rank-1's link lands on my cut
    D.recv(D.corner.p, D.nL * D.nL, D.lower, s);
    const unsigned grid = vgrid(D.nL * D.nL, D.nsm);
    if (D.cplx) corner_op<true><<<grid, 256, 0, s>>>((double2*)F, nk, 0, D.nL, (double2*)D.corner.p, 1);
    else corner_op<false><<<grid, 256, 0, s>>>(F, nk, 0, D.nL, D.corner.p, 1);
    ++*D.launches;
  }
  direct_front_factor_async(*D.front);
  if (D.nR && D.upper >= 0) {
    const unsigned grid = vgrid(D.nR * D.nR, D.nsm);
    if (D.cplx) corner_op<true><<<grid, 256, 0, s>>>((double2*)F, nk, D.nL, D.nR, (double2*)D.corner.p, 0);
    else corner_op<false><<<grid, 256, 0, s>>>(F, nk, D.nL, D.nR, D.corner.p, 0);
    ++*D.launches;
    D.send(D.corner.p, D.nR * D.nR, D.upper, s);
  }
  CUDA_TRY(cudaEventRecord(e1, s));
  CUDA_TRY(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  D.chain_ms = ms;
  if (D.nL) direct_front_read_status(*D.front);
  CUDA_TRY(cudaGetLastError());
  D.schur_ready = true;
}

void dist_schur_info(const DistState& D, double* out4) {
  out4[0] = (double)D.nL; out4[1] = (double)D.nR; out4[2] = D.chain_ms; out4[3] = (double)D.n_bnd;
}
