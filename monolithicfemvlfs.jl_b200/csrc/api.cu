// api.cu - C ABI of libvlfs_b200 (see include/vlfs_b200.h for the reference interfaces
// each entry point replaces).  Host-side bookkeeping only; every byte of arithmetic runs in
// the kernels of integrate.cu / pattern.cu / spmv.cu / krylov.cu.
#include "common.cuh"
#include "solver.cuh"
#include "symbolic_host.hpp"
#include "dist.cuh"
#include <thread>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <array>
#include <map>
#include <cmath>
#include <algorithm>

void build_pattern_device(const std::vector<DomainDev>& doms, long long ndofs,
                          DevBuf<int>& rowptr, DevBuf<int>& colind, DevBuf<int>& rowind,
                          DevBuf<int>& dest_all, std::vector<long long>& dest_off,
                          DevBuf<uint32_t>& perm_out, DevBuf<uint32_t>& run_start,
                          long long* nnz_out, long long* ncoo_out, cudaStream_t s,
                          long long* launches);
void launch_cell_geometry(const DomainDev& dom, int want_gradgrad, double* out, cudaStream_t s);
void launch_mass_reference(const double* valA, int ndA, const double* valB, int ndB, const double* qw, int nq,
                           double* T, cudaStream_t s);
void launch_cell_jacobians(const DomainDev& dom, double* out, cudaStream_t s);
void launch_rhs_affine(const DomainDev& dom, const RhsDev* rhs_dev, int nrhs, const double* cellgeo,
                       double* const* vecs, int nvec, int is_complex, int nsm, cudaStream_t s);
void launch_class_dest(const DomainDev& dom, int nreps, const GatherDom& G, int* tiled, cudaStream_t s);
void launch_build_class_tables(const GatherDom* doms, int d, long long work, int nmat, int is_complex, cudaStream_t s);
void launch_pack_gather_records(const uint32_t* perm, long long ncoo, const GatherDom* doms, int ndom,
                                unsigned long long* rec, cudaStream_t s);
void launch_compact_gather_records(const unsigned long long* rec, long long ncoo, const GatherDom* doms,
                                   uint32_t* out, cudaStream_t s);
void launch_gather_reference(const uint32_t* run_start, const unsigned long long* rec, const uint32_t* rec32,
                             double* const* ctab_all, long long nnz, const GatherDom* doms, int ndom, int nrp_total, long long tab_total,
                             double* const* mats, int nmat, int is_complex, int nsm, cudaStream_t s);
double measure_fp64_fma_tflops(int nsm, int reps, cudaStream_t s);
unsigned chunk_longest_run(const uint32_t* run_start, long long nnz, cudaStream_t s);
void launch_chunk_mark_heads(const uint32_t* run_start, long long nnz, uint32_t* rec, cudaStream_t s);
void build_chunks(const uint32_t* run_start, long long slot_lo, long long slot_hi, unsigned maxrun, const uint32_t* rec,
                  const long long* late_lo, const long long* late_hi, int nlate, DevBuf<unsigned char>& meta_out,
                  long long* nchunks_out, long long* nearly_out, cudaStream_t s, long long* launches);
void launch_gather_chunks(const void* meta, long long first, long long nchunks, const uint32_t* rec,
                          double* const* ctab_all, double* const* mats, int nmat, int is_complex, cudaStream_t s);
bool build_row_templates(const int* rowptr, const uint32_t* run_start, long long nrows, long long nnz, const uint32_t* rec,
                         const long long* late_lo, const long long* late_hi, int nlate, DevBuf<uint32_t>& trec,
                         DevBuf<uint32_t>& trun, long long* T_out, long long* T_early_out, DevBuf<unsigned char>& list,
                         long long* rows_early_out, long long* nclasses_out, cudaStream_t s, long long* launches);
void launch_expand_rows(const void* list, long long first, long long nlist, double* const* tmpl, double* const* mats, int nmat,
                        int is_complex, cudaStream_t s);
void build_csc_device(const int* rowind, const int* colind, long long nnz, long long ndofs,
                      DevBuf<int>& colptr, DevBuf<int>& rowval, DevBuf<int>& csr_slot,
                      cudaStream_t s, long long* launches);
void build_tiled_dest(const DomainDev& dom, const int* dest, DevBuf<int>& tiled, cudaStream_t s,
                      long long* launches);
void launch_integrate(const DomainDev& dom, const RhsDev* rhs_dev, int nrhs, const int* dest,
                      const int* cell_list, long long ncells_run, double* const* mats,
                      int nmat, double* const* vecs, int nvec, int is_complex, int do_mat,
                      int do_vec, int nsm, cudaStream_t s, int dest_by_run = 0, unsigned stage_mask = ~0u);
void launch_gradgrad_reference(const DomainDev& dom, double* Sref, cudaStream_t s);
void launch_functional(const DomainDev& dom, const OpInst& O0, const OpInst& O1, double sc0, double sc1,
                       int wf, const double* x, double* result, int nsm, cudaStream_t s);
void launch_spmv(const int* rowptr, const int* colind, const void* vals, const void* x, void* y,
                 long long nrows, long long nnz, int is_complex, const double* alpha,
                 const double* beta, int nsm, cudaStream_t s);
void launch_axpby_vals(double* target, size_t n, int nsrc, double* const* src,
                       const double* coef_re_im, int is_complex, cudaStream_t s);

namespace {

int op_ncomp(int op, int D) {
  switch (op) {
    case VLFS_OP_VAL: case VLFS_OP_DDIR: case VLFS_OP_LAPL: case VLFS_OP_GRAD_NOWN: return 1;
    case VLFS_OP_GRAD: case VLFS_OP_CHESS_NPLUS: return D;
    case VLFS_OP_HESS: case VLFS_OP_CHESS: return D * (D + 1) / 2;
  }
  return -1;
}
bool op_needs_hess(int op) {
  return op == VLFS_OP_LAPL || op == VLFS_OP_HESS || op == VLFS_OP_CHESS || op == VLFS_OP_CHESS_NPLUS;
}

struct SlotStore {
  DevBuf<double> coords, geo_grad, val, grad, hess, ref_normal, ref_tangent;
  DevBuf<int> variant, dofs;
  std::vector<int> variant_host;      // kept when the slot has a per-cell variant (classification)
};

struct Domain {
  vlfs_domain_desc desc;      // scalar fields only are meaningful after add
  DomainDev dev{};
  SlotStore st[VLFS_MAX_SLOTS];
  DevBuf<double> qweights, weights, C;
  std::vector<vlfs_term_desc> terms;
  std::vector<vlfs_rhs_desc> rhs;
  std::vector<TermDev> terms_host;
  std::vector<RhsDev> rhs_host;
  std::vector<char> term_fast;      // 0 general (cell-centric kernel), 1 grad.grad reference, 2 mass reference
  std::vector<RefProd> refprods_host;   // grouped by block
  std::vector<int> refprod_term;        // source term of every RefProd (coefficient refresh)
  std::vector<double> refprod_scale;    // its slot scales
  int rp_first[VLFS_MAX_BLOCKS + 1] = {0, 0, 0, 0, 0};
  DevBuf<RefProd> refprods_dev;
  DevBuf<double> reftabs, cellgeo;
  // element-block classes: cells with the same Jacobians, table variants and coefficient-field
  // values have the same element block; it is tabulated once per class and assembly
  int ncls = 0;                         // 0: no classes (per-cell paths)
  bool cls_full = false;                // the general (quadrature) terms are tabulated per class too
  bool rhs_fast = false;                // linear forms through the barrier-free affine kernel
  bool cls_dirty = true;
  DevBuf<int> cls, reps, dest_cls;
  DevBuf<double> cls_geo, jt_dev;
  double* ctab_ptr[4] = {nullptr, nullptr, nullptr, nullptr};   // this domain's slice of vlfs_ctx::ctab_all
  long long ctab_base = 0;                                      // its first entry there
  std::vector<double> jt_host, cellgeo_host, weights_host;
  std::vector<int> cls_host, reps_host;      // class of every cell / representative cell of every class
  long long ctab_off[VLFS_MAX_BLOCKS] = {-1, -1, -1, -1};
  long long ctab_entries = 0, ctab_work = 0;
  DevBuf<RhsDev> rhs_dev;
  std::vector<TileTarget> targets_host;
  std::vector<TileProd> prods_host;
  DevBuf<TileTarget> targets_dev;
  DevBuf<TileProd> prods_dev;
  DevBuf<double> Sref;
  DevBuf<int> dest_tiled;           // [cell][tile][4][4] CSR slots for the tiled kernel
  std::vector<std::array<int, 3>> tref_pairs;   // (test op instance, trial op instance, smem offset)
  bool finalized = false;
  bool coef_dirty = true;
};

}  // namespace

#define VLFS_NAUX 5
struct vlfs_ctx {
  int device = 0;
  int nsm = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t aux[VLFS_NAUX] = {};   // class tables of different domains in parallel
  cudaEvent_t ev_fork = nullptr, ev_join[VLFS_NAUX] = {};
  cudaStream_t aux_vec = nullptr;                      // linear forms next to the matrix passes
  cudaEvent_t ev_vec_fork = nullptr, ev_vec_join = nullptr;
  std::string err;
  long long ndofs = 0;
  int is_complex = 0, nmat = 0, nvec = 0;   // nmat: matrices the terms assemble; mats may hold work matrices after them
  int nmat_all() const { return (int)mats.size() > nmat ? (int)mats.size() : nmat; }
  std::vector<std::unique_ptr<Domain>> doms;
  bool pattern_built = false;
  bool imported = false;                // pattern given by vlfs_import_csr: no domains, nothing to assemble
  long long nnz = 0, ncoo = 0;
  DevBuf<int> rowptr, colind, rowind, dest_all;
  DevBuf<uint32_t> perm, run_start;     // sorted COO order and first sorted entry of every slot
  DevBuf<unsigned long long> gather_rec;   // packed (domain, block, local entry, cell) per sorted entry
  DevBuf<uint32_t> gather_rec32;           // index of the entry's class-table value (all domains tabulated)
  DevBuf<double> ctab_all[4];              // per matrix: the class tables of all domains, contiguous
  bool rec32_valid = false;
  // chunked slot-centric pass (gather.cu: gather_chunks) in two segments: chunks whose class tables are
  // ready early, then the others.  Without row templates the records live in gather_rec32 and the
  // values go to the matrices; with row templates (tmpl_valid) the pass runs on the template stream
  // (trec / trun) into tmpl_vals and expand_rows copies the template of every row into the matrices.
  DevBuf<unsigned char> chunk_meta[2];
  const unsigned char* seg_meta[2] = {nullptr, nullptr};
  long long seg_first[2] = {0, 0}, seg_count[2] = {0, 0};
  bool chunks_valid = false, chunks_two_phase = false;
  bool tmpl_valid = false;
  DevBuf<uint32_t> trec, trun;
  DevBuf<double> tmpl_vals[4];
  DevBuf<unsigned char> row_list;            // uint4 per row in processing order (early rows first)
  long long tmpl_slots = 0, tmpl_classes = 0, rows_early = 0;
  int nrp_total = 0;
  long long tab_total = 0;
  DevBuf<GatherDom> gather_doms;
  std::vector<long long> dest_off;
  std::vector<DevBuf<double>> mats, vecs;
  long long launches = 0;
  bool profile = false;
  std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> prof_events;
  SolverPtr solver;
  DirectPtr direct;
  std::unique_ptr<Ilu0> ilu;            // VLFS_PRECOND_BLOCK_ILU0 (ilu0.cu)
  int precond_kind = VLFS_PRECOND_NONE; // of ctx->solver
  double precond_apply_ms = 0;          // last vlfs_precond_apply (CUDA events)
  std::vector<double> dof_coords;
  int coord_dim = 0;
  int solver_mat = -1;
  long long nowned = 0;             // rows [0,nowned) are owned, the rest are ghosts (distributed)
  std::vector<signed char> roles;   // sparse-LU role of every unknown (empty: owned = eliminate)
  DirectPtr dense;                  // dense LU of an interface system (substructuring)
  bool own_stream = true;
  DevBuf<double> dev_partial, dev_h;  // scratch of the vlfs_dev_* reductions
  DistPtr dist;                       // rank of a multi-GPU system (dist.cu); null: single GPU
  bool dist_active() const { return dist && dist_has_halo(*dist); }
  size_t scalar() const { return is_complex ? 2 : 1; }
};

namespace {

int fail(vlfs_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

template <typename F>
int guarded(vlfs_ctx* ctx, F&& f) {
  if (!ctx) return VLFS_ERR_ARG;
  try {
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) throw CudaError{e, __FILE__, __LINE__};
    return f();
  } catch (const CudaError& ce) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)ce.e, cudaGetErrorString(ce.e),
             ce.file, ce.line);
    cudaGetLastError();
    return fail(ctx, VLFS_ERR_CUDA, buf);
  } catch (const NcclError& ne) {
    return fail(ctx, VLFS_ERR_NCCL, ne.msg);
  } catch (const std::string& s) {
    return fail(ctx, VLFS_ERR_STATE, s);
  } catch (const std::bad_alloc&) {
    return fail(ctx, VLFS_ERR_STATE, "host allocation failed");
  } catch (...) {
    return fail(ctx, VLFS_ERR_STATE, "unknown failure");
  }
}

int find_or_add_opi(DomainDev& dev, int slot, int op, const double* dir, int D) {
  for (int o = 0; o < dev.nopi; ++o) {
    const OpInst& O = dev.opi[o];
    if (O.slot == slot && O.op == op &&
        (op != VLFS_OP_DDIR || (O.dir[0] == dir[0] && O.dir[1] == dir[1] && O.dir[2] == dir[2])))
      return o;
  }
  if (dev.nopi >= VLFS_MAX_OPINST) throw std::string("too many distinct operators in one domain");
  OpInst& O = dev.opi[dev.nopi];
  O.slot = slot; O.op = op; O.ncomp = op_ncomp(op, D);
  for (int k = 0; k < 3; ++k) O.dir[k] = dir ? dir[k] : 0.0;
  O.smem_off = 0;
  return dev.nopi++;
}

// (block, matrix, re|im) targets of the tiled bilinear stage and their contractions; rebuilt
// whenever a coefficient changes (a zero coefficient part contributes nothing)
void build_tile_tables(vlfs_ctx* ctx, Domain& d) {
  DomainDev& dev = d.dev;
  d.targets_host.clear();
  d.prods_host.clear();
  for (int b = 0; b < dev.nblocks; ++b) {
    dev.target_first[b] = (int)d.targets_host.size();
    const BlockDev& B = dev.blk[b];
    for (int m = 0; m < ctx->nmat; ++m)
      for (int part = 0; part < (ctx->is_complex ? 2 : 1); ++part) {
        TileTarget tg{m, part, (int)d.prods_host.size(), 0};
        for (const TermDev& T : d.terms_host) {
          if (T.mat != m || T.opi_test[B.ts] < 0 || T.opi_trial[B.us] < 0) continue;
          const double c = (part == 0 ? T.cre : T.cim) * T.st[B.ts] * T.su[B.us];
          if (c == 0.0) continue;
          const OpInst& OA = dev.opi[T.opi_test[B.ts]];
          const OpInst& OB = dev.opi[T.opi_trial[B.us]];
          TileProd P{};
          P.offA = OA.smem_off; P.offB = OB.smem_off; P.bufA = OA.buf_stride; P.bufB = OB.buf_stride;
          P.ndpA = OA.ndp; P.ndpB = OB.ndp; P.ncomp = T.ncomp; P.weight = T.weight; P.coef = c;
          P.tref = -1; P.ldt = OB.ndp; P.nr4 = OA.ndp; P.tref_owner = -1;
          if (T.weight < 0)
            for (auto& pr : d.tref_pairs)
              if (pr[0] == T.opi_test[B.ts] && pr[1] == T.opi_trial[B.us]) P.tref = pr[2];
          d.prods_host.push_back(P);
          ++tg.count;
        }
        if (tg.count) d.targets_host.push_back(tg);
      }
  }
  dev.target_first[dev.nblocks] = (int)d.targets_host.size();
  for (size_t k = 0; k < d.prods_host.size(); ++k) {
    if (d.prods_host[k].tref < 0) continue;
    size_t owner = k;
    for (size_t l = 0; l < k; ++l)
      if (d.prods_host[l].tref == d.prods_host[k].tref) { owner = l; break; }
    d.prods_host[k].tref_owner = (int)owner;
  }
  d.targets_dev.upload(d.targets_host.data(), d.targets_host.size(), ctx->stream);
  d.prods_dev.upload(d.prods_host.data(), d.prods_host.size(), ctx->stream);
  dev.targets = d.targets_dev.p;
  dev.prods = d.prods_dev.p;
}

// reference products per block, their tables and the per-cell geometric factors
void build_reference_products(vlfs_ctx* ctx, Domain& d) {
  DomainDev& dev = d.dev;
  const int D = dev.D, NS = D * (D + 1) / 2;
  d.refprods_host.clear(); d.refprod_term.clear(); d.refprod_scale.clear();
  std::vector<std::array<long long, 4>> mass_tabs;     // (slot s, slot u, offset) of built mass tables
  long long tab_len = 0, gg_off = -1;
  bool want_gg = false;
  for (int b = 0; b < dev.nblocks; ++b) {
    d.rp_first[b] = (int)d.refprods_host.size();
    const BlockDev& B = dev.blk[b];
    for (size_t t = 0; t < d.terms.size(); ++t) {
      const vlfs_term_desc& T = d.terms[t];
      if (!d.term_fast[t] || T.test_scale[B.ts] == 0.0 || T.trial_scale[B.us] == 0.0) continue;
      RefProd P{};
      P.mat = T.mat;
      const double sc = T.test_scale[B.ts] * T.trial_scale[B.us];
      P.cre = T.coef_re * sc; P.cim = T.coef_im * sc;
      if (d.term_fast[t] == 1) {
        if (gg_off < 0) { gg_off = tab_len; tab_len += (long long)NS * B.nr * B.nc; }
        P.ntab = NS; P.geo_idx = 1; P.tab_off = gg_off; want_gg = true;
      } else {
        long long off = -1;
        for (auto& m : mass_tabs) if (m[0] == B.ts && m[1] == B.us) off = m[2];
        if (off < 0) { off = tab_len; tab_len += (long long)B.nr * B.nc; mass_tabs.push_back({B.ts, B.us, off, 0}); }
        P.ntab = 1; P.geo_idx = 0; P.tab_off = off;
      }
      d.refprods_host.push_back(P);
      d.refprod_term.push_back((int)t);
      d.refprod_scale.push_back(sc);
    }
  }
  d.rp_first[dev.nblocks] = (int)d.refprods_host.size();
  for (int b = dev.nblocks + 1; b <= VLFS_MAX_BLOCKS; ++b) d.rp_first[b] = d.rp_first[dev.nblocks];
  // linear forms: value tables without per-cell variant on constant-Jacobian cells
  d.rhs_fast = dev.affine && !d.rhs.empty();
  for (const vlfs_rhs_desc& R : d.rhs)
    for (int sl = 0; sl < dev.nslots; ++sl)
      if (R.scale[sl] != 0.0 && (R.op != VLFS_OP_VAL || dev.slot[sl].variant)) d.rhs_fast = false;
  if (d.refprods_host.empty() && d.rhs_fast) {
    d.cellgeo.alloc((size_t)dev.ncells * VLFS_CELLGEO);
    launch_cell_geometry(dev, 0, d.cellgeo.p, ctx->stream);
    ++ctx->launches;
  }
  if (d.refprods_host.empty()) return;
  d.reftabs.alloc((size_t)tab_len);
  if (gg_off >= 0) { launch_gradgrad_reference(dev, d.reftabs.p + gg_off, ctx->stream); ++ctx->launches; }
  for (auto& m : mass_tabs) {
    const SlotDev& A = dev.slot[m[0]];
    const SlotDev& Bs = dev.slot[m[1]];
    launch_mass_reference(A.val, A.ndofs, Bs.val, Bs.ndofs, dev.qweights, dev.nq, d.reftabs.p + m[2], ctx->stream);
    ++ctx->launches;
  }
  d.cellgeo.alloc((size_t)dev.ncells * VLFS_CELLGEO);
  launch_cell_geometry(dev, want_gg ? 1 : 0, d.cellgeo.p, ctx->stream);
  ++ctx->launches;
  d.refprods_dev.upload(d.refprods_host.data(), d.refprods_host.size(), ctx->stream);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  d.cellgeo_host.resize((size_t)dev.ncells * VLFS_CELLGEO);
  CUDA_TRY(cudaMemcpy(d.cellgeo_host.data(), d.cellgeo.p, d.cellgeo_host.size() * sizeof(double), cudaMemcpyDeviceToHost));
}

// Groups the cells of a constant-Jacobian domain into element-block classes and lays out the
// per-class tables.  Key = Jacobians of every slot rounded to 2^-46 of their largest entry (cells
// of a regular mesh differ only by round-off, ~1e-16, far below the 1e-12 entry tolerance; every
// member uses its representative's block), table variants, and the exact bits of the coefficient
// fields the bilinear terms use.  Worth it when there are few classes (regular meshes); otherwise
// the per-cell paths stay.
void classify_domain(vlfs_ctx* ctx, Domain& d) {
  DomainDev& dev = d.dev;
  d.cls_dirty = false;
  d.ncls = 0; d.cls_full = false; d.cls_host.clear(); d.reps_host.clear();
  for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) d.ctab_off[b] = -1;
  const bool any_general = !d.terms_host.empty();
  if (!dev.affine || dev.ncells == 0 || (d.refprods_host.empty() && !any_general)) return;
  if (d.jt_host.empty()) {
    d.jt_dev.alloc((size_t)dev.ncells * dev.nslots * 9);
    launch_cell_jacobians(dev, d.jt_dev.p, ctx->stream);
    ++ctx->launches;
    d.jt_host.resize(d.jt_dev.n);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(cudaMemcpy(d.jt_host.data(), d.jt_dev.p, d.jt_host.size() * sizeof(double), cudaMemcpyDeviceToHost));
    d.jt_dev.release();
  }
  std::vector<int> wused;
  for (const vlfs_term_desc& T : d.terms)
    if (T.weight >= 0 && std::find(wused.begin(), wused.end(), T.weight) == wused.end()) wused.push_back(T.weight);
  const int njt = dev.nslots * 9;
  std::vector<double> jmax(dev.nslots, 0.0);
  for (long long e = 0; e < dev.ncells; ++e)
    for (int k = 0; k < njt; ++k) jmax[k / 9] = std::max(jmax[k / 9], std::fabs(d.jt_host[(size_t)e * njt + k]));
  const size_t klen = (size_t)njt + dev.nslots + wused.size() * dev.nq;
  std::map<std::vector<long long>, int> classes;
  std::vector<int> cls((size_t)dev.ncells), reps;
  std::vector<long long> key(klen);
  const int max_cls = 4096;
  const size_t wstride = (size_t)dev.ncells * dev.nq;
  for (long long e = 0; e < dev.ncells; ++e) {
    size_t p = 0;
    for (int k = 0; k < njt; ++k)
      key[p++] = jmax[k / 9] > 0 ? std::llround(d.jt_host[(size_t)e * njt + k] / jmax[k / 9] * 70368744177664.0) : 0;
    for (int sl = 0; sl < dev.nslots; ++sl) key[p++] = d.st[sl].variant_host.empty() ? 0 : d.st[sl].variant_host[e];
    for (int w : wused)
      for (int q = 0; q < dev.nq; ++q) {
        long long bits;
        memcpy(&bits, &d.weights_host[(size_t)w * wstride + (size_t)e * dev.nq + q], sizeof bits);
        key[p++] = bits;
      }
    auto it = classes.find(key);
    if (it == classes.end()) {
      if ((int)classes.size() >= max_cls) { classes.clear(); break; }
      it = classes.emplace(key, (int)classes.size()).first;
      reps.push_back((int)e);
    }
    cls[e] = it->second;
  }
  const int ncls = (int)classes.size();
  if (getenv("VLFS_DEBUG")) fprintf(stderr, "[vlfs] domain: %lld cells, %d reference products, %d general terms, %d classes\n",
                                    (long long)dev.ncells, (int)d.refprods_host.size(), (int)d.terms_host.size(), ncls);
  if (ncls == 0) return;
  d.ncls = ncls;
  d.cls_full = any_general && (long long)ncls * 4 <= dev.ncells;
  d.cls.upload(cls.data(), cls.size(), ctx->stream);
  d.reps.upload(reps.data(), reps.size(), ctx->stream);
  d.cls_host = cls; d.reps_host = reps;
  if (!d.cellgeo_host.empty()) {
    std::vector<double> cg((size_t)ncls * VLFS_CELLGEO);
    for (int c = 0; c < ncls; ++c)
      memcpy(&cg[(size_t)c * VLFS_CELLGEO], &d.cellgeo_host[(size_t)reps[c] * VLFS_CELLGEO], VLFS_CELLGEO * sizeof(double));
    d.cls_geo.upload(cg.data(), cg.size(), ctx->stream);
  }
  long long off = 0;
  d.ctab_work = 0;
  for (int b = 0; b < dev.nblocks; ++b) {
    if (d.cls_full || d.rp_first[b + 1] > d.rp_first[b]) {
      d.ctab_off[b] = off;
      const long long ne = (long long)dev.blk[b].nr * dev.blk[b].nc;
      off += ncls * ne;
      d.ctab_work = std::max(d.ctab_work, ncls * ne);
    }
  }
  d.ctab_entries = off;
  if (d.cls_full) {
    GatherDom G{};
    for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) G.ctab_off[b] = d.ctab_off[b];
    d.dest_cls.alloc((size_t)ncls * dev.tile_base[dev.nblocks] * 16);
    launch_class_dest(dev, ncls, G, d.dest_cls.p, ctx->stream);
    ++ctx->launches;
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
}

// derive operator instances, touched blocks, shared-memory layout, device term tables
void finalize_domain(vlfs_ctx* ctx, Domain& d) {
  DomainDev& dev = d.dev;
  const int D = dev.D;
  dev.nopi = 0;
  d.terms_host.clear();
  d.rhs_host.clear();
  d.term_fast.assign(d.terms.size(), 0);
  bool touched[VLFS_MAX_SLOTS][VLFS_MAX_SLOTS] = {};
  for (size_t t = 0; t < d.terms.size(); ++t) {
    const vlfs_term_desc& T = d.terms[t];
    int nslots_t = 0, nslots_u = 0;
    for (int s = 0; s < dev.nslots; ++s) { nslots_t += T.test_scale[s] != 0.0; nslots_u += T.trial_scale[s] != 0.0; }
    // reference-type terms (constant-Jacobian cells, no coefficient field): the entry is a
    // per-cell geometric factor times a cell-independent reference matrix; they are evaluated by
    // the slot-centric gather pass instead of the cell-centric kernel
    int fast = 0;
    if (dev.affine && T.weight < 0) {
      if (T.test_op == VLFS_OP_GRAD && T.trial_op == VLFS_OP_GRAD && dev.nslots == 1 &&
          dev.slot[0].dref == D && dev.slot[0].nvar == 1 && dev.measure_kind == VLFS_MEASURE_CELL)
        fast = 1;
      else if (T.test_op == VLFS_OP_VAL && T.trial_op == VLFS_OP_VAL) {
        fast = 2;
        for (int s = 0; s < dev.nslots; ++s)
          if ((T.test_scale[s] != 0.0 || T.trial_scale[s] != 0.0) && dev.slot[s].variant) fast = 0;
      }
    }
    d.term_fast[t] = (char)fast;
    TermDev td{};
    td.mat = T.mat; td.weight = T.weight; td.cre = T.coef_re; td.cim = T.coef_im;
    td.ncomp = op_ncomp(T.test_op, D);
    for (int s = 0; s < VLFS_MAX_SLOTS; ++s) {
      td.opi_test[s] = td.opi_trial[s] = -1;
      td.st[s] = td.su[s] = 0.0;
    }
    for (int s = 0; s < dev.nslots; ++s) {
      if (T.test_scale[s] != 0.0) {
        td.st[s] = T.test_scale[s];
        if (!fast) td.opi_test[s] = find_or_add_opi(dev, s, T.test_op, T.test_dir, D);
      }
      if (T.trial_scale[s] != 0.0) {
        td.su[s] = T.trial_scale[s];
        if (!fast) td.opi_trial[s] = find_or_add_opi(dev, s, T.trial_op, T.trial_dir, D);
      }
      for (int u = 0; u < dev.nslots; ++u)
        if (T.test_scale[s] != 0.0 && T.trial_scale[u] != 0.0) touched[s][u] = true;
    }
    if (!fast) d.terms_host.push_back(td);
  }
  for (const vlfs_rhs_desc& R : d.rhs) {
    RhsDev rd{};
    rd.vec = R.vec; rd.wre = R.weight_re; rd.wim = R.weight_im; rd.cre = R.coef_re; rd.cim = R.coef_im;
    rd.ncomp = 1;
    for (int s = 0; s < VLFS_MAX_SLOTS; ++s) { rd.opi[s] = -1; rd.sc[s] = 0.0; }
    for (int s = 0; s < dev.nslots; ++s)
      if (R.scale[s] != 0.0) {
        rd.sc[s] = R.scale[s];
        rd.opi[s] = find_or_add_opi(dev, s, R.op, R.dir, D);
      }
    d.rhs_host.push_back(rd);
  }
  // blocks in (test slot, trial slot) order
  dev.nblocks = 0;
  int off = 0;
  for (int s = 0; s < dev.nslots; ++s)
    for (int u = 0; u < dev.nslots; ++u)
      if (touched[s][u]) {
        BlockDev& B = dev.blk[dev.nblocks++];
        B.ts = s; B.us = u; B.nr = dev.slot[s].ndofs; B.nc = dev.slot[u].ndofs; B.off = off;
        off += B.nr * B.nc;
      }
  dev.entries_per_cell = off;
  // 4x4 tiles of the touched blocks
  dev.tile_base[0] = 0;
  for (int b = 0; b < dev.nblocks; ++b) {
    dev.tiles_j[b] = (dev.blk[b].nc + 3) / 4;
    dev.tile_base[b + 1] = dev.tile_base[b] + ((dev.blk[b].nr + 3) / 4) * dev.tiles_j[b];
  }
  // shared memory layout (doubles); operator arrays start on 32-byte boundaries
  for (int o = 0; o < dev.nopi; ++o) {
    OpInst& O = dev.opi[o];
    if (op_needs_hess(O.op) && !dev.slot[O.slot].hess)
      throw std::string("a second-derivative operator needs the slot's hess table");
    if ((O.op == VLFS_OP_CHESS || O.op == VLFS_OP_CHESS_NPLUS) && !dev.C)
      throw std::string("C tensor missing for a C:Hessian operator");
    if ((O.op == VLFS_OP_GRAD_NOWN) && !dev.slot[O.slot].ref_normal)
      throw std::string("ref_normal missing for a normal-derivative operator");
    if ((O.op == VLFS_OP_CHESS_NPLUS) && !dev.slot[0].ref_normal)
      throw std::string("ref_normal of slot 0 missing for (C:Hessian).n+");
    O.ndp = (dev.slot[O.slot].ndofs + 3) & ~3;
  }
  dev.nqg = dev.affine ? 1 : dev.nq;
  dev.smem_meas_off = dev.nslots * dev.nqg * 12;
  dev.smem_geo_stride = (dev.smem_meas_off + dev.nq + 1 + 3) & ~3;
  int p = 4 * dev.smem_geo_stride;
  dev.smem_inv_off = p;
  for (int o = 0; o < dev.nopi; ++o) {         // values of a slot with one table variant do
    OpInst& O = dev.opi[o];                    // not depend on the cell
    if (O.op == VLFS_OP_VAL && !dev.slot[O.slot].variant) {
      O.smem_off = p; O.buf_stride = 0;
      p += dev.nq * O.ncomp * O.ndp;
    } else {
      // per-cell table (marked before the reference-matrix pairs below are chosen: a stale 0 here
      // made second-derivative products of the first cell stand in for every cell of a CTA)
      O.buf_stride = -1;
    }
  }
  dev.smem_inv_len = p - dev.smem_inv_off;
  // reference matrices of contractions between two cell-invariant tables without a coefficient
  // field on an affine domain: one [nr4][nc4] matrix per distinct (test op, trial op) pair
  dev.smem_tref_off = p;
  d.tref_pairs.clear();
  if (dev.affine)
    for (const TermDev& T : d.terms_host)
      for (int b = 0; b < dev.nblocks; ++b) {
        const int oa = T.opi_test[dev.blk[b].ts], ob = T.opi_trial[dev.blk[b].us];
        if (oa < 0 || ob < 0 || T.weight >= 0) continue;
        if (dev.opi[oa].buf_stride != 0 || dev.opi[ob].buf_stride != 0) continue;
        bool seen = false;
        for (auto& pr : d.tref_pairs) seen |= pr[0] == oa && pr[1] == ob;
        if (seen) continue;
        d.tref_pairs.push_back({oa, ob, p});
        p += dev.opi[oa].ndp * dev.opi[ob].ndp;
      }
  dev.smem_tref_len = p - dev.smem_tref_off;
  dev.smem_buf_off = p;
  int q = (dev.nweights * dev.nq + 3) & ~3;
  for (int o = 0; o < dev.nopi; ++o) {
    OpInst& O = dev.opi[o];
    if (O.op == VLFS_OP_VAL && !dev.slot[O.slot].variant) continue;
    O.smem_off = p + q; O.buf_stride = -1;
    q += dev.nq * O.ncomp * O.ndp;
  }
  dev.smem_buf_stride = q;
  for (int o = 0; o < dev.nopi; ++o)
    if (dev.opi[o].buf_stride < 0) dev.opi[o].buf_stride = q;
  // three table buffers allow the split-phase pipeline barrier; fall back to two when the
  // per-cell tables are large (second-derivative operators)
  // (large tables: two buffers when that lets one more CTA share the SM)
  {
    auto ctas = [&](int nb) { return (int)((220 * 1024) / ((size_t)(p + nb * q) * sizeof(double) + 1024)); };
    dev.nbuf = ((size_t)(p + 3 * q) * sizeof(double) <= 200 * 1024 && ctas(3) >= std::min(ctas(2), 3)) ? 3 : 2;
  }
  p += dev.nbuf * q;
  dev.smem_total = p;
  if ((size_t)p * sizeof(double) > 227 * 1024)
    throw std::string("domain needs more than 227 KB of shared memory per cell");
  build_tile_tables(ctx, d);
  d.rhs_dev.upload(d.rhs_host.data(), d.rhs_host.size(), ctx->stream);
  build_reference_products(ctx, d);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  d.finalized = true;
  d.coef_dirty = false;
}

void refresh_coefs(vlfs_ctx* ctx, Domain& d) {
  if (!d.coef_dirty) return;
  size_t k = 0;
  for (size_t t = 0; t < d.terms.size(); ++t) {
    if (d.term_fast[t]) continue;
    d.terms_host[k].cre = d.terms[t].coef_re;
    d.terms_host[k].cim = d.terms[t].coef_im;
    ++k;
  }
  for (size_t t = 0; t < d.rhs.size(); ++t) {
    d.rhs_host[t].cre = d.rhs[t].coef_re;
    d.rhs_host[t].cim = d.rhs[t].coef_im;
  }
  build_tile_tables(ctx, d);
  d.rhs_dev.upload(d.rhs_host.data(), d.rhs_host.size(), ctx->stream);
  if (!d.refprods_host.empty()) {                 // in place: gather_doms holds this pointer
    for (size_t k = 0; k < d.refprods_host.size(); ++k) {
      const vlfs_term_desc& T = d.terms[d.refprod_term[k]];
      d.refprods_host[k].cre = T.coef_re * d.refprod_scale[k];
      d.refprods_host[k].cim = T.coef_im * d.refprod_scale[k];
    }
    CUDA_TRY(cudaMemcpyAsync(d.refprods_dev.p, d.refprods_host.data(), d.refprods_host.size() * sizeof(RefProd),
                             cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  d.coef_dirty = false;
}

struct ProfScope {
  vlfs_ctx* ctx; cudaEvent_t e0{}, e1{}; bool on;
  ProfScope(vlfs_ctx* c, const std::string& name) : ctx(c), on(c->profile) {
    if (!on) return;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, ctx->stream);
    ctx->prof_events.push_back({name, {e0, e1}});
  }
  ~ProfScope() { if (on) cudaEventRecord(e1, ctx->stream); }
};

// one contiguous class-table array per matrix: domain slices in domain order
void layout_class_tables(vlfs_ctx* ctx) {
  long long total = 0;
  for (auto& d : ctx->doms) {
    d->ctab_base = total;
    if (d->ncls > 0) total += d->ctab_entries;
  }
  for (int m = 0; m < ctx->nmat; ++m) {
    if (ctx->ctab_all[m].n != (size_t)total * ctx->scalar()) ctx->ctab_all[m].alloc((size_t)total * ctx->scalar());
    for (auto& d : ctx->doms) d->ctab_ptr[m] = d->ncls > 0 ? ctx->ctab_all[m].p + (size_t)d->ctab_base * ctx->scalar() : nullptr;
  }
}

void upload_gather_doms(vlfs_ctx* ctx) {
  layout_class_tables(ctx);
  std::vector<GatherDom> gd(ctx->doms.size());
  for (size_t di = 0; di < ctx->doms.size(); ++di) {
    Domain& d = *ctx->doms[di];
    GatherDom& G = gd[di];
    G.coo_lo = ctx->dest_off[di];
    G.coo_hi = G.coo_lo + d.dev.ncells * (long long)d.dev.entries_per_cell;
    G.epc = d.dev.entries_per_cell > 0 ? d.dev.entries_per_cell : 1;
    G.nblocks = d.dev.nblocks;
    for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) G.blk[b] = d.dev.blk[b];
    for (int b = 0; b < d.dev.nblocks; ++b)
      if (d.dev.blk[b].nr * d.dev.blk[b].nc > 16384)
        throw std::string("a touched block has more than 16384 entries per cell (gather record field)");
    for (int b = 0; b <= VLFS_MAX_BLOCKS; ++b) G.rp_first[b] = d.rp_first[b];
    G.rp = d.refprods_dev.p; G.geo = d.cellgeo.p; G.tabs = d.reftabs.p; G.tab_len = (int)d.reftabs.n;
    G.ncls = d.ncls; G.cls = d.cls.p; G.cls_geo = d.cls_geo.p;
    for (int m = 0; m < 4; ++m) G.ctab_m[m] = d.ctab_ptr[m];
    G.ctab_base = d.ctab_base;
    for (int b = 0; b < VLFS_MAX_BLOCKS; ++b) G.ctab_off[b] = d.ctab_off[b];
  }
  if (ctx->gather_doms.n != gd.size()) ctx->gather_doms.alloc(gd.size());
  CUDA_TRY(cudaMemcpyAsync(ctx->gather_doms.p, gd.data(), gd.size() * sizeof(GatherDom), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
}

// 32-bit class records (the index of the COO entry's value in the contiguous class tables):
// possible when every domain that owns COO entries is class-tabulated
void rebuild_compact_records(vlfs_ctx* ctx) {
  ctx->rec32_valid = false;
  if (ctx->gather_rec.n == 0) return;
  long long total = 0;
  for (auto& d : ctx->doms) {
    if (d->dev.ncells * (long long)d->dev.entries_per_cell == 0) continue;
    if (d->ncls == 0) return;
    total += d->ctab_entries;
  }
  if (total >= 0xffffffffll) return;
  if (ctx->gather_rec32.n != ctx->gather_rec.n) ctx->gather_rec32.alloc(ctx->gather_rec.n);
  launch_compact_gather_records(ctx->gather_rec.p, (long long)ctx->gather_rec.n, ctx->gather_doms.p,
                                ctx->gather_rec32.p, ctx->stream);
  ++ctx->launches;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->rec32_valid = true;
}

// Chunk tables of the slot-centric pass, from the 32-bit class records (every domain
// class-tabulated).  Rebuilt when the element-block classes change.  The records are converted in
// place to the chunk format (head bit), so the plain compact gather no longer applies afterwards.
// Rows with identical record streams share one template (VLFS_ASM_TEMPLATES=0 switches that off).
void rebuild_chunks(vlfs_ctx* ctx) {
  ctx->chunks_valid = ctx->tmpl_valid = false;
  static const bool enabled = !(getenv("VLFS_ASM_CHUNKS") && atoi(getenv("VLFS_ASM_CHUNKS")) == 0);
  static const bool templates = !(getenv("VLFS_ASM_TEMPLATES") && atoi(getenv("VLFS_ASM_TEMPLATES")) == 0);
  if (!enabled || !ctx->rec32_valid || ctx->nnz == 0) return;
  long long total = 0;
  for (auto& d : ctx->doms) {
    if (d->dev.ncells * (long long)d->dev.entries_per_cell > 0 && !d->terms_host.empty() && !d->cls_full) return;
    total += d->ctab_entries;
  }
  if (total >= 0x7fffffffll) return;
  const unsigned maxrun = chunk_longest_run(ctx->run_start.p, ctx->nnz, ctx->stream);
  if (maxrun == 0) return;                    // a run longer than the shared-memory window: plain gather
  std::vector<long long> lo, hi;
  bool any_early = false;
  for (auto& d : ctx->doms) {
    if (d->ncls == 0 || d->ctab_entries == 0) continue;
    if (d->cls_full) { lo.push_back(d->ctab_base); hi.push_back(d->ctab_base + d->ctab_entries); }
    else any_early = true;
  }
  DevBuf<long long> dlo, dhi;
  const long long none = 0;
  dlo.upload(lo.empty() ? &none : lo.data(), std::max<size_t>(lo.size(), 1), ctx->stream);
  dhi.upload(hi.empty() ? &none : hi.data(), std::max<size_t>(hi.size(), 1), ctx->stream);
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  launch_chunk_mark_heads(ctx->run_start.p, ctx->nnz, ctx->gather_rec32.p, ctx->stream);
  ++ctx->launches;
  ctx->rec32_valid = false;                   // the records now carry the head bit
  long long T = 0, Te = 0;
  if (templates && build_row_templates(ctx->rowptr.p, ctx->run_start.p, ctx->ndofs, ctx->nnz, ctx->gather_rec32.p, dlo.p, dhi.p,
                                       (int)lo.size(), ctx->trec, ctx->trun, &T, &Te, ctx->row_list, &ctx->rows_early,
                                       &ctx->tmpl_classes, ctx->stream, &ctx->launches)) {
    ctx->tmpl_slots = T;
    for (int m = 0; m < ctx->nmat; ++m)
      if (ctx->tmpl_vals[m].n != (size_t)T * ctx->scalar()) ctx->tmpl_vals[m].alloc((size_t)T * ctx->scalar());
    long long ne = 0;
    build_chunks(ctx->trun.p, 0, Te, maxrun, ctx->trec.p, dlo.p, dhi.p, 0, ctx->chunk_meta[0], &ctx->seg_count[0], &ne,
                 ctx->stream, &ctx->launches);
    build_chunks(ctx->trun.p, Te, T, maxrun, ctx->trec.p, dlo.p, dhi.p, 0, ctx->chunk_meta[1], &ctx->seg_count[1], &ne,
                 ctx->stream, &ctx->launches);
    ctx->seg_meta[0] = ctx->chunk_meta[0].p; ctx->seg_meta[1] = ctx->chunk_meta[1].p;
    ctx->seg_first[0] = ctx->seg_first[1] = 0;
    ctx->tmpl_valid = true;
    if (getenv("VLFS_DEBUG"))
      fprintf(stderr, "[vlfs] row templates: %lld classes of rows for %lld rows, %lld template slots for %lld values (%lld early)\n",
              ctx->tmpl_classes, ctx->ndofs, T, ctx->nnz, Te);
  } else {
    ctx->trec.release(); ctx->trun.release(); ctx->row_list.release();
    for (int m = 0; m < 4; ++m) ctx->tmpl_vals[m].release();
    long long nchunks = 0, ne = 0;
    build_chunks(ctx->run_start.p, 0, ctx->nnz, maxrun, ctx->gather_rec32.p, dlo.p, dhi.p, (int)lo.size(), ctx->chunk_meta[0],
                 &nchunks, &ne, ctx->stream, &ctx->launches);
    ctx->chunk_meta[1].release();
    ctx->seg_meta[0] = ctx->seg_meta[1] = ctx->chunk_meta[0].p;
    ctx->seg_first[0] = 0; ctx->seg_count[0] = ne;
    ctx->seg_first[1] = ne; ctx->seg_count[1] = nchunks - ne;
  }
  ctx->chunks_two_phase = !lo.empty() && any_early && ctx->seg_count[0] > 0 && ctx->seg_count[1] > 0;
  ctx->chunks_valid = true;
}

// slot-centric fallback records, built on demand from the sorted COO order
void ensure_gather_records(vlfs_ctx* ctx) {
  if (ctx->gather_rec.n == (size_t)ctx->ncoo || ctx->ncoo == 0) return;
  if (ctx->perm.n != (size_t)ctx->ncoo) throw std::string("internal: sorted COO order no longer available");
  ctx->gather_rec.alloc((size_t)ctx->ncoo);
  launch_pack_gather_records(ctx->perm.p, ctx->ncoo, ctx->gather_doms.p, (int)ctx->doms.size(), ctx->gather_rec.p, ctx->stream);
  ++ctx->launches;
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
}

// Which numeric pass runs: with every domain class-tabulated the chunked slot-centric pass, else the
// plain slot-centric gather (32- or 64-bit records) followed by the cell-centric kernels.
void choose_numeric_pass(vlfs_ctx* ctx) {
  ctx->chunks_valid = false;
  ensure_gather_records(ctx);
  rebuild_compact_records(ctx);
  rebuild_chunks(ctx);
}

void assemble_impl(vlfs_ctx* ctx, bool do_mat, bool do_vec) {
  if (!ctx->pattern_built) throw std::string("vlfs_build_pattern must be called first");
  if (ctx->imported) throw std::string("the matrix was imported (vlfs_import_csr): nothing to assemble");
  cudaStream_t s = ctx->stream;
  std::vector<double*> mp, vp;
  for (int m = 0; m < ctx->nmat; ++m) mp.push_back(ctx->mats[m].p);     // work matrices are not assembled
  for (auto& v : ctx->vecs) vp.push_back(v.p);
  for (auto& dp : ctx->doms) refresh_coefs(ctx, *dp);
  {
    bool reclass = false;
    for (auto& dp : ctx->doms)
      if (dp->cls_dirty) { classify_domain(ctx, *dp); reclass = true; }
    if (reclass) {
      upload_gather_doms(ctx);
      choose_numeric_pass(ctx);
    }
  }
  const bool use_chunks = do_mat && ctx->chunks_valid;
  const bool two_phase = use_chunks && ctx->chunks_two_phase;
  // linear forms of constant-Jacobian domains touch only the vectors: they run on their own stream
  // beside the matrix passes (when both are assembled and nobody is timing kernels one by one)
  const bool vec_aside = do_vec && do_mat && !ctx->profile;
  bool vec_join_early = false;
  if (vec_aside) {
    if (!ctx->aux_vec) {
      CUDA_TRY(cudaStreamCreateWithFlags(&ctx->aux_vec, cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_vec_fork, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_vec_join, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_vec_fork, s));
    CUDA_TRY(cudaStreamWaitEvent(ctx->aux_vec, ctx->ev_vec_fork, 0));
    for (auto& v : ctx->vecs) v.zero(ctx->aux_vec);
    for (size_t di = 0; di < ctx->doms.size(); ++di) {
      Domain& d = *ctx->doms[di];
      if (d.rhs_host.empty() || !d.rhs_fast) continue;
      launch_rhs_affine(d.dev, d.rhs_dev.p, (int)d.rhs_host.size(), d.cellgeo.p, vp.data(), (int)vp.size(),
                        ctx->is_complex, ctx->nsm, ctx->aux_vec);
      ++ctx->launches;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_vec_join, ctx->aux_vec));
    // general vector passes (if any) run on the main stream and need the zero fill: join now;
    // otherwise the join is the last thing this function does
    for (auto& dp : ctx->doms) if (!dp->rhs_host.empty() && !dp->rhs_fast) vec_join_early = true;
    if (vec_join_early) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_vec_join, 0));
  }
  bool used[VLFS_NAUX] = {};
  if (do_mat) {
    // per-class element blocks: reference products tabulated directly, general (quadrature) terms
    // integrated on the class representatives only, reduced into the class tables.  These are
    // small latency-bound launches: the domains run side by side on auxiliary streams.
    ProfScope ps(ctx, "class_tables");
    if (!ctx->ev_fork) {
      CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
      for (int k = 0; k < VLFS_NAUX; ++k) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join[k], cudaEventDisableTiming));
      }
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_fork, s));
    int k = 0;
    for (size_t di = 0; di < ctx->doms.size(); ++di) {
      Domain& d = *ctx->doms[di];
      if (d.ncls == 0) continue;
      // row-centric pass in two phases: the directly tabulated domains share stream 0 (their tables
      // are ready first), the quadrature-tabulated ones take the other streams in turn
      const int q = two_phase ? (d.cls_full ? 1 + (k++ % (VLFS_NAUX - 1)) : 0) : (k++ % VLFS_NAUX);
      cudaStream_t a = ctx->aux[q];
      if (!used[q]) { CUDA_TRY(cudaStreamWaitEvent(a, ctx->ev_fork, 0)); used[q] = true; }
      launch_build_class_tables(ctx->gather_doms.p, (int)di, d.ctab_work, ctx->nmat, ctx->is_complex, a);
      ++ctx->launches;
      if (d.cls_full) {
        double* cm[4] = {d.ctab_ptr[0], d.ctab_ptr[1], d.ctab_ptr[2], d.ctab_ptr[3]};
        launch_integrate(d.dev, d.rhs_dev.p, 0, d.dest_cls.p, d.reps.p, d.ncls, cm, ctx->nmat, vp.data(),
                         (int)vp.size(), ctx->is_complex, 1, 0, ctx->nsm, a, 1);
        ++ctx->launches;
      }
    }
    for (int q = 0; q < VLFS_NAUX; ++q)
      if (used[q]) {
        CUDA_TRY(cudaEventRecord(ctx->ev_join[q], ctx->aux[q]));
        if (!(two_phase && q > 0)) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_join[q], 0));
      }
  }
  if (use_chunks) {
    double* ca[4] = {ctx->ctab_all[0].p, ctx->ctab_all[1].p, ctx->ctab_all[2].p, ctx->ctab_all[3].p};
    double* tv[4] = {ctx->tmpl_vals[0].p, ctx->tmpl_vals[1].p, ctx->tmpl_vals[2].p, ctx->tmpl_vals[3].p};
    const bool tm = ctx->tmpl_valid;
    auto chunks = [&](int seg) {
      if (ctx->seg_count[seg] == 0) return;
      launch_gather_chunks(ctx->seg_meta[seg], ctx->seg_first[seg], ctx->seg_count[seg], tm ? ctx->trec.p : ctx->gather_rec32.p,
                           ca, tm ? tv : mp.data(), (int)mp.size(), ctx->is_complex, s);
      ++ctx->launches;
    };
    auto expand = [&](long long first, long long n) {
      if (n == 0) return;
      launch_expand_rows(ctx->row_list.p, first, n, tv, mp.data(), (int)mp.size(), ctx->is_complex, s);
      ++ctx->launches;
    };
    if (two_phase) {
      { ProfScope p1(ctx, "gather_chunks:early"); chunks(0); }
      if (tm) { ProfScope p1(ctx, "expand_rows:early"); expand(0, ctx->rows_early); }
      {
        ProfScope pw(ctx, "class_tables:quadrature_wait");
        for (int q = 1; q < VLFS_NAUX; ++q) if (used[q]) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_join[q], 0));
      }
      { ProfScope p2(ctx, "gather_chunks:late"); chunks(1); }
      if (tm) { ProfScope p2(ctx, "expand_rows:late"); expand(ctx->rows_early, ctx->ndofs - ctx->rows_early); }
    } else {
      { ProfScope p0(ctx, "gather_chunks"); chunks(0); chunks(1); }
      if (tm) { ProfScope p0(ctx, "expand_rows"); expand(0, ctx->ndofs); }
    }
  }
  {
    // slot-centric pass: writes every CSR value (zero fill included) with the tabulated /
    // reference-type contributions summed in COO order; the cell-centric kernels then add the rest
    ProfScope ps(ctx, use_chunks ? "vector_zero" : "gather_reference");
    if (do_mat && !use_chunks) {
      double* ca[4] = {ctx->ctab_all[0].p, ctx->ctab_all[1].p, ctx->ctab_all[2].p, ctx->ctab_all[3].p};
      launch_gather_reference(ctx->run_start.p, ctx->gather_rec.p,
                              ctx->rec32_valid ? ctx->gather_rec32.p : nullptr, ca, ctx->nnz, ctx->gather_doms.p,
                              (int)ctx->doms.size(), ctx->nrp_total, ctx->tab_total, mp.data(), (int)mp.size(),
                              ctx->is_complex, ctx->nsm, s);
      ++ctx->launches;
    }
    if (do_vec && !vec_aside) for (auto& v : ctx->vecs) v.zero(s);
  }
  for (size_t di = 0; di < ctx->doms.size(); ++di) {
    Domain& d = *ctx->doms[di];
    bool gen_mat = do_mat && !d.terms_host.empty() && !d.cls_full;
    bool gen_vec = do_vec && !d.rhs_host.empty();
    if (gen_vec && d.rhs_fast && vec_aside) gen_vec = false;      // done on the side stream
    if (gen_vec && d.rhs_fast) {
      ProfScope ps(ctx, "rhs_affine:dom" + std::to_string(di));
      launch_rhs_affine(d.dev, d.rhs_dev.p, (int)d.rhs_host.size(), d.cellgeo.p, vp.data(), (int)vp.size(),
                        ctx->is_complex, ctx->nsm, s);
      ++ctx->launches;
      gen_vec = false;
    }
    if (gen_mat || gen_vec) {
      ProfScope ps(ctx, "integrate_cells:dom" + std::to_string(di));
      unsigned mask = ~0u;
      if (!gen_mat) {                      // vector-only pass: stage the linear terms' operators only
        mask = 0u;
        for (const RhsDev& R : d.rhs_host)
          for (int sl = 0; sl < d.dev.nslots; ++sl) if (R.opi[sl] >= 0) mask |= 1u << R.opi[sl];
      }
      launch_integrate(d.dev, d.rhs_dev.p, (int)d.rhs_host.size(), d.dest_tiled.p, nullptr, d.dev.ncells,
                       mp.data(), (int)mp.size(), vp.data(), (int)vp.size(), ctx->is_complex,
                       gen_mat, gen_vec, ctx->nsm, s, 0, mask);
      ++ctx->launches;
    }
  }
  if (vec_aside && !vec_join_early) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_vec_join, 0));
}

// status of the factorisation behind ctx->direct (after direct_numeric)
int pivot_status(vlfs_ctx* ctx) {
  long long pert = 0, bad = 0;
  direct_pivot_status(*ctx->direct, &pert, &bad);
  char buf[160];
  if (bad > 0) {
    snprintf(buf, sizeof buf, "sparse LU: %lld non-finite pivots (singular or overflowing matrix)", bad);
    return fail(ctx, VLFS_ERR_SINGULAR, buf);
  }
  if (pert > 0) {
    snprintf(buf, sizeof buf, "sparse LU: %lld tiny pivots replaced by a static perturbation", pert);
    return fail(ctx, VLFS_PIVOT_PERTURBED, buf);
  }
  return VLFS_OK;
}

int ilu_status(vlfs_ctx* ctx) {
  char buf[160];
  if (ctx->ilu->nonfinite > 0) {
    snprintf(buf, sizeof buf, "ILU(0): %lld non-finite pivots", ctx->ilu->nonfinite);
    return fail(ctx, VLFS_ERR_SINGULAR, buf);
  }
  if (ctx->ilu->perturbed > 0) {
    snprintf(buf, sizeof buf, "ILU(0): %lld tiny pivots replaced", ctx->ilu->perturbed);
    return fail(ctx, VLFS_PIVOT_PERTURBED, buf);
  }
  return VLFS_OK;
}

template <typename T>
void upload_opt(DevBuf<T>& b, const T* h, size_t n, cudaStream_t s) {
  if (h && n) b.upload(h, n, s);
}

}  // namespace

namespace {
// one host thread per rank: the collective entry points block on each other through NCCL
template <typename F>
int run_ranks(int n, vlfs_ctx** ctxs, F&& f) {
  if (!ctxs || n < 1) return VLFS_ERR_ARG;
  std::vector<int> rc(n, VLFS_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < n; ++r) th.emplace_back([&, r]() { rc[r] = f(r, ctxs[r]); });
  for (auto& t : th) t.join();
  int worst = VLFS_OK;
  for (int r = 0; r < n; ++r)
    if (rc[r] != VLFS_OK && (worst == VLFS_OK || worst == VLFS_NOT_CONVERGED || worst == VLFS_PIVOT_PERTURBED)) worst = rc[r];
  return worst;
}
}  // namespace

extern "C" {

int vlfs_version(void) { return 100; }

int vlfs_device_count(int* n) {
  if (!n) return VLFS_ERR_ARG;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) { *n = 0; cudaGetLastError(); return VLFS_ERR_CUDA; }
  return VLFS_OK;
}

int vlfs_create(vlfs_ctx** out, int device) {
  if (!out) return VLFS_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
    cudaGetLastError();
    return VLFS_ERR_CUDA;      // no CPU fallback: without a GPU the library refuses to exist
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VLFS_ERR_CUDA;
  if (prop.major < 10) return VLFS_ERR_CUDA;   // sm_100a code only
  vlfs_ctx* c = new (std::nothrow) vlfs_ctx();
  if (!c) return VLFS_ERR_STATE;
  c->device = device;
  c->nsm = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return VLFS_ERR_CUDA;
  }
  *out = c;
  return VLFS_OK;
}

int vlfs_destroy(vlfs_ctx* ctx) {
  if (!ctx) return VLFS_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->solver.reset();
  ctx->dist.reset();
  ctx->direct.reset(); ctx->ilu.reset();
  ctx->dense.reset();
  ctx->doms.clear();
  ctx->mats.clear();
  ctx->vecs.clear();
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->aux_vec) cudaStreamDestroy(ctx->aux_vec);
  if (ctx->ev_vec_fork) cudaEventDestroy(ctx->ev_vec_fork);
  if (ctx->ev_vec_join) cudaEventDestroy(ctx->ev_vec_join);
  for (int k = 0; k < VLFS_NAUX; ++k) {
    if (ctx->aux[k]) cudaStreamDestroy(ctx->aux[k]);
    if (ctx->ev_join[k]) cudaEventDestroy(ctx->ev_join[k]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  delete ctx;
  return VLFS_OK;
}

const char* vlfs_last_error(const vlfs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int vlfs_system_init(vlfs_ctx* ctx, int64_t ndofs, int is_complex, int nmat, int nvec) {
  return guarded(ctx, [&]() {
    if (ndofs <= 0 || ndofs >= (1ll << 31) || nmat < 1 || nmat > 4 || nvec < 0 || nvec > 4)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_system_init: bad sizes (1<=nmat<=4, 0<=nvec<=4)");
    ctx->ndofs = ndofs; ctx->is_complex = is_complex ? 1 : 0; ctx->nmat = nmat; ctx->nvec = nvec;
    ctx->nowned = ndofs;
    ctx->doms.clear(); ctx->mats.clear(); ctx->vecs.clear();
    ctx->pattern_built = false; ctx->imported = false; ctx->solver.reset(); ctx->direct.reset(); ctx->ilu.reset(); ctx->dense.reset();
    ctx->dof_coords.clear(); ctx->coord_dim = 0; ctx->roles.clear(); ctx->solver_mat = -1;
    ctx->chunks_valid = ctx->tmpl_valid = ctx->rec32_valid = false;       // numeric-pass tables of the old pattern
    ctx->gather_rec.release(); ctx->gather_rec32.release(); ctx->trec.release(); ctx->trun.release(); ctx->row_list.release();
    ctx->chunk_meta[0].release(); ctx->chunk_meta[1].release();
    for (int m = 0; m < 4; ++m) ctx->tmpl_vals[m].release();
    ctx->vecs.resize(nvec);
    for (auto& v : ctx->vecs) { v.alloc((size_t)ndofs * ctx->scalar()); v.zero(ctx->stream); }
    return VLFS_OK;
  });
}

int vlfs_add_domain(vlfs_ctx* ctx, const vlfs_domain_desc* desc, int* domain_id) {
  return guarded(ctx, [&]() {
    if (!desc || !domain_id) return fail(ctx, VLFS_ERR_ARG, "null argument");
    if (ctx->ndofs == 0) return fail(ctx, VLFS_ERR_STATE, "vlfs_system_init first");
    if (ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern already built");
    const vlfs_domain_desc& d = *desc;
    if (d.D < 1 || d.D > 3 || d.nq < 1 || d.nslots < 1 || d.nslots > VLFS_MAX_SLOTS ||
        d.ncells < 0 || d.measure_slot < 0 || d.measure_slot >= d.nslots ||
        d.nweights < 0 || d.nweights > VLFS_MAX_WEIGHTS || !d.qweights)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_add_domain: bad descriptor");
    auto dom = std::make_unique<Domain>();
    dom->desc = d;
    cudaStream_t s = ctx->stream;
    DomainDev& dev = dom->dev;
    dev.D = d.D; dev.nq = d.nq; dev.nslots = d.nslots; dev.measure_slot = d.measure_slot;
    dev.measure_kind = d.measure_kind; dev.nweights = d.nweights; dev.affine = d.affine;
    dev.ncells = d.ncells;
    dom->qweights.upload(d.qweights, d.nq, s);
    dev.qweights = dom->qweights.p;
    if (d.nweights) {
      if (!d.weights) return fail(ctx, VLFS_ERR_ARG, "weights missing");
      dom->weights.upload(d.weights, (size_t)d.nweights * d.ncells * d.nq, s);
      dom->weights_host.assign(d.weights, d.weights + (size_t)d.nweights * d.ncells * d.nq);
    }
    dev.weights = dom->weights.p;
    if (d.C) dom->C.upload(d.C, (size_t)d.D * d.D * d.D * d.D, s);
    dev.C = dom->C.p;
    for (int k = 0; k < d.nslots; ++k) {
      const vlfs_slot_desc& S = d.slot[k];
      if (S.dref < 0 || S.dref > d.D || S.nv < 1 || S.nvar < 1 || S.ndofs < 1 || !S.coords ||
          !S.geo_grad || !S.val || !S.dofs || (S.dref > 0 && !S.grad))
        return fail(ctx, VLFS_ERR_ARG, "vlfs_add_domain: bad slot descriptor");
      SlotStore& st = dom->st[k];
      size_t nc = (size_t)d.ncells;
      // a variant array with one distinct value (e.g. every boundary facet is the same local
      // face of its cell) selects one table set: upload only that set, so that the value
      // tables become cell-invariant for the kernel
      int nvar = S.nvar, v0 = 0;
      bool uniform = S.variant != nullptr && nc > 0;
      if (uniform) {
        v0 = S.variant[0];
        for (size_t c = 1; c < nc && uniform; ++c) uniform = S.variant[c] == v0;
      }
      if (S.variant)
        for (size_t c = 0; c < nc; ++c)
          if (S.variant[c] < 0 || S.variant[c] >= S.nvar)
            return fail(ctx, VLFS_ERR_ARG, "vlfs_add_domain: variant index out of range");
      if (uniform) nvar = 1; else v0 = 0;
      const size_t tq = (size_t)d.nq;
      st.coords.upload(S.coords, nc * S.nv * d.D, s);
      if (!uniform) upload_opt(st.variant, S.variant, nc, s);
      if (!uniform && S.variant) st.variant_host.assign(S.variant, S.variant + nc);
      st.geo_grad.upload(S.geo_grad + v0 * tq * S.nv * S.dref, (size_t)nvar * tq * S.nv * S.dref, s);
      st.val.upload(S.val + v0 * tq * S.ndofs, (size_t)nvar * tq * S.ndofs, s);
      st.grad.upload(S.grad + v0 * tq * S.ndofs * S.dref, (size_t)nvar * tq * S.ndofs * S.dref, s);
      if (S.hess)
        st.hess.upload(S.hess + v0 * tq * S.ndofs * S.dref * S.dref,
                       (size_t)nvar * tq * S.ndofs * S.dref * S.dref, s);
      st.dofs.upload(S.dofs, nc * S.ndofs, s);
      if (S.ref_normal) st.ref_normal.upload(S.ref_normal + (size_t)v0 * S.dref, (size_t)nvar * S.dref, s);
      if (S.ref_tangent) st.ref_tangent.upload(S.ref_tangent + (size_t)v0 * S.dref, (size_t)nvar * S.dref, s);
      SlotDev& sd = dev.slot[k];
      sd.dref = S.dref; sd.nv = S.nv; sd.nvar = nvar; sd.ndofs = S.ndofs;
      sd.has_hess = S.hess != nullptr;
      sd.coords = st.coords.p; sd.variant = st.variant.p; sd.geo_grad = st.geo_grad.p;
      sd.val = st.val.p; sd.grad = st.grad.p; sd.hess = st.hess.p; sd.dofs = st.dofs.p;
      sd.ref_normal = st.ref_normal.p; sd.ref_tangent = st.ref_tangent.p;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    *domain_id = (int)ctx->doms.size();
    ctx->doms.push_back(std::move(dom));
    return VLFS_OK;
  });
}

int vlfs_add_term(vlfs_ctx* ctx, int domain_id, const vlfs_term_desc* term, int* term_id) {
  return guarded(ctx, [&]() {
    if (!term || domain_id < 0 || domain_id >= (int)ctx->doms.size())
      return fail(ctx, VLFS_ERR_ARG, "vlfs_add_term: bad domain");
    if (ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern already built");
    Domain& d = *ctx->doms[domain_id];
    const int D = d.dev.D;
    if (term->mat < 0 || term->mat >= ctx->nmat || term->weight >= d.dev.nweights ||
        op_ncomp(term->test_op, D) < 0 || op_ncomp(term->trial_op, D) < 0 ||
        op_ncomp(term->test_op, D) != op_ncomp(term->trial_op, D))
      return fail(ctx, VLFS_ERR_ARG, "vlfs_add_term: bad term (operators must have equal component counts)");
    if ((int)d.terms.size() >= VLFS_MAX_TERMS) return fail(ctx, VLFS_ERR_ARG, "too many terms");
    if (term_id) *term_id = (int)d.terms.size();
    d.terms.push_back(*term);
    return VLFS_OK;
  });
}

int vlfs_add_rhs_term(vlfs_ctx* ctx, int domain_id, const vlfs_rhs_desc* term, int* term_id) {
  return guarded(ctx, [&]() {
    if (!term || domain_id < 0 || domain_id >= (int)ctx->doms.size())
      return fail(ctx, VLFS_ERR_ARG, "vlfs_add_rhs_term: bad domain");
    if (ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern already built");
    Domain& d = *ctx->doms[domain_id];
    if (term->vec < 0 || term->vec >= ctx->nvec || term->weight_re >= d.dev.nweights ||
        term->weight_im >= d.dev.nweights || op_ncomp(term->op, d.dev.D) != 1)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_add_rhs_term: bad term (scalar operators only)");
    if ((int)d.rhs.size() >= VLFS_MAX_TERMS) return fail(ctx, VLFS_ERR_ARG, "too many terms");
    if (term_id) *term_id = (int)d.rhs.size();
    d.rhs.push_back(*term);
    return VLFS_OK;
  });
}

int vlfs_set_term_coef(vlfs_ctx* ctx, int domain_id, int term_id, double re, double im) {
  return guarded(ctx, [&]() {
    if (domain_id < 0 || domain_id >= (int)ctx->doms.size()) return fail(ctx, VLFS_ERR_ARG, "bad domain");
    Domain& d = *ctx->doms[domain_id];
    if (term_id < 0 || term_id >= (int)d.terms.size()) return fail(ctx, VLFS_ERR_ARG, "bad term");
    d.terms[term_id].coef_re = re; d.terms[term_id].coef_im = im; d.coef_dirty = true;
    return VLFS_OK;
  });
}

int vlfs_set_rhs_coef(vlfs_ctx* ctx, int domain_id, int term_id, double re, double im) {
  return guarded(ctx, [&]() {
    if (domain_id < 0 || domain_id >= (int)ctx->doms.size()) return fail(ctx, VLFS_ERR_ARG, "bad domain");
    Domain& d = *ctx->doms[domain_id];
    if (term_id < 0 || term_id >= (int)d.rhs.size()) return fail(ctx, VLFS_ERR_ARG, "bad term");
    d.rhs[term_id].coef_re = re; d.rhs[term_id].coef_im = im; d.coef_dirty = true;
    return VLFS_OK;
  });
}

int vlfs_set_weights(vlfs_ctx* ctx, int domain_id, int weight, const double* values) {
  return guarded(ctx, [&]() {
    if (domain_id < 0 || domain_id >= (int)ctx->doms.size() || !values) return fail(ctx, VLFS_ERR_ARG, "bad domain");
    Domain& d = *ctx->doms[domain_id];
    if (weight < 0 || weight >= d.dev.nweights) return fail(ctx, VLFS_ERR_ARG, "bad weight index");
    size_t n = (size_t)d.dev.ncells * d.dev.nq;
    CUDA_TRY(cudaMemcpyAsync(d.weights.p + (size_t)weight * n, values, n * sizeof(double),
                             cudaMemcpyHostToDevice, ctx->stream));
    // a changed field that a bilinear term uses changes the element-block classes; fields that only
    // linear forms use (the incident-wave data of a frequency sweep) need no host copy and no compare
    bool in_matrix = false;
    for (const vlfs_term_desc& T : d.terms) in_matrix |= T.weight == weight;
    if (in_matrix && memcmp(d.weights_host.data() + (size_t)weight * n, values, n * sizeof(double)) != 0) {
      memcpy(d.weights_host.data() + (size_t)weight * n, values, n * sizeof(double));
      // The classes stay valid when the new field is still constant over every class (e.g. μ₂ = k μ₁ in
      // a frequency sweep: the same partition, new values): the class tables are recomputed from the
      // representatives at every assembly anyway, so nothing has to be rebuilt.
      bool same_partition = d.ncls > 0 && d.cls_host.size() == (size_t)d.dev.ncells;
      if (same_partition) {
        const size_t nq = (size_t)d.dev.nq;
        const double* w = values;
        for (long long e = 0; e < d.dev.ncells && same_partition; ++e) {
          const size_t r = (size_t)d.reps_host[(size_t)d.cls_host[(size_t)e]];
          same_partition = memcmp(w + (size_t)e * nq, w + r * nq, nq * sizeof(double)) == 0;
        }
      }
      if (!same_partition) d.cls_dirty = true;
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_build_pattern(vlfs_ctx* ctx, int64_t* nnz, int64_t* ncoo) {
  return guarded(ctx, [&]() {
    if (ctx->doms.empty()) return fail(ctx, VLFS_ERR_STATE, "no domains");
    // solver, LU and interface objects hold raw pointers into the pattern and value arrays: a
    // second build would leave them dangling, so it is refused (vlfs_system_init starts over)
    if (ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern already built (vlfs_system_init starts a new system)");
    std::vector<DomainDev> devs;
    for (auto& d : ctx->doms) { finalize_domain(ctx, *d); devs.push_back(d->dev); }
    long long n1 = 0, n2 = 0;
    build_pattern_device(devs, ctx->ndofs, ctx->rowptr, ctx->colind, ctx->rowind, ctx->dest_all,
                         ctx->dest_off, ctx->perm, ctx->run_start, &n1, &n2, ctx->stream, &ctx->launches);
    ctx->nnz = n1; ctx->ncoo = n2;
    {
      for (auto& d : ctx->doms) classify_domain(ctx, *d);
      upload_gather_doms(ctx);
      ctx->nrp_total = 0;
      ctx->tab_total = 0;
      for (auto& d : ctx->doms) { ctx->nrp_total += (int)d->refprods_host.size(); ctx->tab_total += (long long)d->reftabs.n; }
      if ((int)ctx->doms.size() > 16) throw std::string("more than 16 integration domains");
      choose_numeric_pass(ctx);
    }
    ctx->nnz = n1; ctx->ncoo = n2;
    for (size_t di = 0; di < ctx->doms.size(); ++di) {
      Domain& d = *ctx->doms[di];
      // (kept also for class-tabulated domains: a coefficient field may later break the classes)
      if (!d.terms_host.empty())
        build_tiled_dest(d.dev, ctx->dest_all.p + ctx->dest_off[di], d.dest_tiled, ctx->stream, &ctx->launches);
    }
    ctx->mats.clear();
    ctx->mats.resize(ctx->nmat);
    for (auto& m : ctx->mats) { m.alloc((size_t)n1 * ctx->scalar()); m.zero(ctx->stream); }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->pattern_built = true;
    if (nnz) *nnz = n1;
    if (ncoo) *ncoo = n2;
    return VLFS_OK;
  });
}

int vlfs_import_csr(vlfs_ctx* ctx, int64_t ndofs, int is_complex, int nmat, const int32_t* rowptr,
                    const int32_t* colind) {
  return guarded(ctx, [&]() {
    if (!rowptr || !colind || ndofs <= 0 || ndofs >= (1ll << 31) || nmat < 1 || nmat > 4 || rowptr[0] != 0)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_import_csr: bad arguments (0-based CSR, 1<=nmat<=4)");
    const long long nnz = rowptr[ndofs];
    if (nnz <= 0) return fail(ctx, VLFS_ERR_ARG, "vlfs_import_csr: empty matrix");
    std::vector<int> rowind((size_t)nnz);
    for (long long i = 0; i < ndofs; ++i) {
      if (rowptr[i + 1] < rowptr[i]) return fail(ctx, VLFS_ERR_ARG, "vlfs_import_csr: rowptr not monotone");
      for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        if (colind[p] < 0 || colind[p] >= ndofs) return fail(ctx, VLFS_ERR_ARG, "vlfs_import_csr: column index out of range");
        if (p > rowptr[i] && colind[p] <= colind[p - 1])
          return fail(ctx, VLFS_ERR_ARG, "vlfs_import_csr: column indices must ascend strictly inside a row");
        rowind[p] = (int)i;
      }
    }
    int rc = vlfs_system_init(ctx, ndofs, is_complex, nmat, 1);
    if (rc != VLFS_OK) return rc;
    ctx->rowptr.upload(rowptr, (size_t)ndofs + 1, ctx->stream);
    ctx->colind.upload(colind, (size_t)nnz, ctx->stream);
    ctx->rowind.upload(rowind.data(), (size_t)nnz, ctx->stream);
    ctx->nnz = nnz; ctx->ncoo = nnz;
    ctx->mats.clear();
    ctx->mats.resize(nmat);
    for (auto& m : ctx->mats) { m.alloc((size_t)nnz * ctx->scalar()); m.zero(ctx->stream); }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->pattern_built = true;
    ctx->imported = true;
    return VLFS_OK;
  });
}

int vlfs_set_values(vlfs_ctx* ctx, int mat, const void* values) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || !values) return fail(ctx, VLFS_ERR_ARG, "bad matrix");
    CUDA_TRY(cudaMemcpyAsync(ctx->mats[mat].p, values, ctx->nnz * ctx->scalar() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_set_vector(vlfs_ctx* ctx, int vec, const void* values) {
  return guarded(ctx, [&]() {
    if (vec < 0 || vec >= ctx->nvec || !values) return fail(ctx, VLFS_ERR_ARG, "bad vector");
    CUDA_TRY(cudaMemcpyAsync(ctx->vecs[vec].p, values, ctx->ndofs * ctx->scalar() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_get_pattern(vlfs_ctx* ctx, int32_t* rowptr, int32_t* colind) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern not built");
    if (rowptr) CUDA_TRY(cudaMemcpy(rowptr, ctx->rowptr.p, (ctx->ndofs + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (colind) CUDA_TRY(cudaMemcpy(colind, ctx->colind.p, ctx->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    return VLFS_OK;
  });
}

int vlfs_get_pattern_csc(vlfs_ctx* ctx, int32_t* colptr, int32_t* rowval, int32_t* csr_slot) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "pattern not built");
    DevBuf<int> cp, rv, sl;
    build_csc_device(ctx->rowind.p, ctx->colind.p, ctx->nnz, ctx->ndofs, cp, rv, sl, ctx->stream,
                     &ctx->launches);
    if (colptr) CUDA_TRY(cudaMemcpy(colptr, cp.p, (ctx->ndofs + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (rowval) CUDA_TRY(cudaMemcpy(rowval, rv.p, ctx->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    if (csr_slot) CUDA_TRY(cudaMemcpy(csr_slot, sl.p, ctx->nnz * sizeof(int), cudaMemcpyDeviceToHost));
    return VLFS_OK;
  });
}

int vlfs_assemble(vlfs_ctx* ctx) {
  return guarded(ctx, [&]() { assemble_impl(ctx, true, true); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); return VLFS_OK; });
}
int vlfs_assemble_matrices(vlfs_ctx* ctx) {
  return guarded(ctx, [&]() { assemble_impl(ctx, true, false); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); return VLFS_OK; });
}
int vlfs_assemble_vectors(vlfs_ctx* ctx) {
  return guarded(ctx, [&]() { assemble_impl(ctx, false, true); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); return VLFS_OK; });
}

int vlfs_get_values(vlfs_ctx* ctx, int mat, void* values) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || !values) return fail(ctx, VLFS_ERR_ARG, "bad matrix");
    CUDA_TRY(cudaMemcpy(values, ctx->mats[mat].p, ctx->nnz * ctx->scalar() * sizeof(double), cudaMemcpyDeviceToHost));
    return VLFS_OK;
  });
}

int vlfs_get_vector(vlfs_ctx* ctx, int vec, void* values) {
  return guarded(ctx, [&]() {
    if (vec < 0 || vec >= ctx->nvec || !values) return fail(ctx, VLFS_ERR_ARG, "bad vector");
    CUDA_TRY(cudaMemcpy(values, ctx->vecs[vec].p, ctx->ndofs * ctx->scalar() * sizeof(double), cudaMemcpyDeviceToHost));
    return VLFS_OK;
  });
}

int vlfs_combine(vlfs_ctx* ctx, int target, int n, const int* src, const double* coef) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || target < 0 || target >= ctx->nmat_all() || n < 1 || n > 4 || !src || !coef)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_combine: bad arguments");
    double* sp[4];
    for (int k = 0; k < n; ++k) {
      if (src[k] < 0 || src[k] >= ctx->nmat_all()) return fail(ctx, VLFS_ERR_ARG, "vlfs_combine: bad source");
      sp[k] = ctx->mats[src[k]].p;
    }
    launch_axpby_vals(ctx->mats[target].p, (size_t)ctx->nnz, n, sp, coef, ctx->is_complex, ctx->stream);
    ++ctx->launches;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_add_work_matrix(vlfs_ctx* ctx, int* mat) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || !mat) return fail(ctx, VLFS_ERR_ARG, "vlfs_add_work_matrix: after vlfs_build_pattern / vlfs_import_csr");
    if (ctx->mats.size() >= 16) return fail(ctx, VLFS_ERR_STATE, "vlfs_add_work_matrix: at most 16 value arrays");
    if (ctx->mats.capacity() < 16) ctx->mats.reserve(16);
    ctx->mats.emplace_back();
    ctx->mats.back().alloc((size_t)ctx->nnz * ctx->scalar());
    ctx->mats.back().zero(ctx->stream);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *mat = (int)ctx->mats.size() - 1;
    return VLFS_OK;
  });
}

int vlfs_spmv(vlfs_ctx* ctx, int mat, const void* x, void* y) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || !x || !y) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    size_t nb = (size_t)ctx->ndofs * ctx->scalar();
    DevBuf<double> dx, dy;
    dx.upload((const double*)x, nb, ctx->stream);
    dy.alloc(nb); dy.zero(ctx->stream);
    const double one[2] = {1, 0}, zero[2] = {0, 0};
    if (ctx->dist_active()) dist_hooks(*ctx->dist)->spmv(ctx->mats[mat].p, dx.p, dy.p, one, zero);   // owned rows, halo first
    else launch_spmv(ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p, dx.p, dy.p, ctx->ndofs, ctx->nnz,
                     ctx->is_complex, one, zero, ctx->nsm, ctx->stream);
    ++ctx->launches;
    CUDA_TRY(cudaMemcpyAsync(y, dy.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_spmv_device(vlfs_ctx* ctx, int mat, int reps, float* ms_total) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || reps < 1) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    size_t nb = (size_t)ctx->ndofs * ctx->scalar();
    DevBuf<double> dx, dy;
    std::vector<double> hx(nb);
    for (size_t i = 0; i < (size_t)ctx->ndofs; ++i) {
      if (ctx->is_complex) { hx[2 * i] = sin(0.1 * (double)i); hx[2 * i + 1] = cos(0.1 * (double)i); }
      else hx[i] = sin(0.1 * (double)i);
    }
    dx.upload(hx.data(), nb, ctx->stream);
    dy.alloc(nb);
    const double one[2] = {1, 0}, zero[2] = {0, 0};
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    for (int r = 0; r < reps; ++r) {
      if (ctx->dist_active()) dist_hooks(*ctx->dist)->spmv(ctx->mats[mat].p, dx.p, dy.p, one, zero);
      else launch_spmv(ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p, dx.p, dy.p, ctx->ndofs, ctx->nnz,
                       ctx->is_complex, one, zero, ctx->nsm, ctx->stream);
    }
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->launches += reps;
    if (ms_total) *ms_total = ms;
    return VLFS_OK;
  });
}

int vlfs_assemble_timed(vlfs_ctx* ctx, int reps, float* ms_total) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || reps < 1) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    for (int r = 0; r < reps; ++r) assemble_impl(ctx, true, true);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_total) *ms_total = ms;
    return VLFS_OK;
  });
}

int vlfs_assemble_profile(vlfs_ctx* ctx, int reps, char* names, int names_cap, float* ms, int ms_cap, int* nrec) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || reps < 1 || !names || !ms || !nrec) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    std::vector<std::string> nm;
    std::vector<double> acc;
    for (int r = 0; r < reps; ++r) {
      ctx->profile = true;
      ctx->prof_events.clear();
      assemble_impl(ctx, true, true);
      ctx->profile = false;
      CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (r == 0) { nm.resize(ctx->prof_events.size()); acc.assign(ctx->prof_events.size(), 0.0); }
      for (size_t k = 0; k < ctx->prof_events.size(); ++k) {
        float t = 0;
        cudaEventElapsedTime(&t, ctx->prof_events[k].second.first, ctx->prof_events[k].second.second);
        acc[k] += t; nm[k] = ctx->prof_events[k].first;
        cudaEventDestroy(ctx->prof_events[k].second.first);
        cudaEventDestroy(ctx->prof_events[k].second.second);
      }
      ctx->prof_events.clear();
    }
    std::string joined;
    int n = 0;
    for (size_t k = 0; k < nm.size() && (int)k < ms_cap; ++k) {
      joined += nm[k]; joined += ";";
      ms[k] = (float)(acc[k] / reps); ++n;
    }
    snprintf(names, names_cap, "%s", joined.c_str());
    *nrec = n;
    return VLFS_OK;
  });
}

int vlfs_measure_fp64_peak(vlfs_ctx* ctx, double* tflops) {
  return guarded(ctx, [&]() {
    if (!tflops) return fail(ctx, VLFS_ERR_ARG, "null argument");
    *tflops = measure_fp64_fma_tflops(ctx->nsm, 5, ctx->stream);
    ctx->launches += 7;
    return VLFS_OK;
  });
}

int vlfs_last_kernel_launches(vlfs_ctx* ctx, int64_t* n) {
  if (!ctx || !n) return VLFS_ERR_ARG;
  *n = ctx->launches;
  return VLFS_OK;
}

int vlfs_device_ptrs(vlfs_ctx* ctx, int mat, void** rowptr, void** colind, void** values) {
  if (!ctx || !ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all()) return VLFS_ERR_ARG;
  if (rowptr) *rowptr = ctx->rowptr.p;
  if (colind) *colind = ctx->colind.p;
  if (values) *values = ctx->mats[mat].p;
  return VLFS_OK;
}

int vlfs_solver_setup(vlfs_ctx* ctx, int mat, const vlfs_solver_opts* opts) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || !opts) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    if (opts->precond < VLFS_PRECOND_NONE || opts->precond > VLFS_PRECOND_SPARSE_LU)
      return fail(ctx, VLFS_ERR_ARG, "vlfs_solver_setup: unknown preconditioner");
    if (opts->precond == VLFS_PRECOND_BLOCK_ILU0) {
      // zero-fill incomplete LU of the block of owned rows x owned columns; level analysis once per pattern
      // block_size_hint = B > 0: block Jacobi of ILU(0) blocks of B consecutive rows, one CTA per block
      const int B = opts->block_size_hint > 0 && opts->block_size_hint < ctx->nowned ? opts->block_size_hint : 0;
      if (!ctx->ilu || ctx->ilu->block_rows != B)
        ctx->ilu = ilu0_symbolic(ctx->rowptr.p, ctx->colind.p, ctx->nowned, ctx->is_complex, B, ctx->stream);
      ilu0_numeric(*ctx->ilu, ctx->mats[mat].p, &ctx->launches);
    }
    if (opts->precond == VLFS_PRECOND_SPARSE_LU) {
      const bool chain = ctx->dist_active() && dist_nranks(*ctx->dist) > 1;
      if (chain && !ctx->direct) dist_roles(*ctx->dist, ctx->roles);     // cuts kept, lower ghosts ignored
      if (!ctx->direct) {
        if (ctx->dof_coords.empty()) {
          // no DOF coordinates (an imported matrix): graph distances stand in for them
          std::vector<int> rp((size_t)ctx->ndofs + 1), ci((size_t)ctx->nnz);
          CUDA_TRY(cudaMemcpy(rp.data(), ctx->rowptr.p, rp.size() * sizeof(int), cudaMemcpyDeviceToHost));
          CUDA_TRY(cudaMemcpy(ci.data(), ctx->colind.p, ci.size() * sizeof(int), cudaMemcpyDeviceToHost));
          vlfs_sym::pseudo_coords(ctx->ndofs, rp.data(), ci.data(), ctx->dof_coords);
          ctx->coord_dim = 3;
        }
        std::vector<signed char> role = ctx->roles;
        if (role.empty()) {
          role.assign((size_t)ctx->ndofs, 2);
          for (long long i = 0; i < ctx->nowned; ++i) role[i] = 0;
        }
        ctx->direct = direct_symbolic(ctx->rowptr.p, ctx->colind.p, ctx->ndofs, role.data(), ctx->nnz, ctx->dof_coords.data(),
                                      ctx->coord_dim, opts->block_size_hint, ctx->is_complex, ctx->nsm,
                                      ctx->stream, &ctx->launches);
      }
      direct_numeric(*ctx->direct, ctx->rowind.p, ctx->colind.p, ctx->mats[mat].p);
      if (chain) dist_schur_factor(*ctx->dist, *ctx->direct, ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p);
    }
    ctx->solver = solver_create(ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p, ctx->nowned, ctx->nnz,
                                ctx->is_complex, *opts, ctx->nsm, ctx->stream, &ctx->launches,
                                ctx->direct.get(), ctx->dist_active() ? dist_hooks(*ctx->dist) : nullptr, ctx->ilu.get());
    ctx->solver_mat = mat;
    ctx->precond_kind = opts->precond;
    if (opts->precond == VLFS_PRECOND_BLOCK_ILU0) return ilu_status(ctx);
    return opts->precond == VLFS_PRECOND_SPARSE_LU ? pivot_status(ctx) : VLFS_OK;
  });
}

int vlfs_set_dof_coords(vlfs_ctx* ctx, int dim, const double* coords) {
  return guarded(ctx, [&]() {
    if (dim < 1 || dim > 3 || !coords || ctx->ndofs == 0) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    ctx->dof_coords.assign(coords, coords + (size_t)ctx->ndofs * dim);
    ctx->coord_dim = dim;
    ctx->direct.reset(); ctx->ilu.reset();
    return VLFS_OK;
  });
}

int vlfs_solver_refactor(vlfs_ctx* ctx, int mat) {
  return guarded(ctx, [&]() {
    if (ctx->solver && ctx->ilu && ctx->precond_kind == VLFS_PRECOND_BLOCK_ILU0 && mat >= 0 && mat < ctx->nmat_all()) {
      ilu0_numeric(*ctx->ilu, ctx->mats[mat].p, &ctx->launches);      // same levels, new values
      solver_set_matrix(*ctx->solver, ctx->mats[mat].p);
      ctx->solver_mat = mat;
      return ilu_status(ctx);
    }
    if (!ctx->solver || !ctx->direct || mat < 0 || mat >= ctx->nmat_all())
      return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup with VLFS_PRECOND_SPARSE_LU or VLFS_PRECOND_BLOCK_ILU0 first");
    direct_numeric(*ctx->direct, ctx->rowind.p, ctx->colind.p, ctx->mats[mat].p);
    if (ctx->dist_active() && dist_nranks(*ctx->dist) > 1)
      dist_schur_factor(*ctx->dist, *ctx->direct, ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p);
    // the Krylov residual must use the matrix that was just factorised
    solver_set_matrix(*ctx->solver, ctx->mats[mat].p);
    ctx->solver_mat = mat;
    return pivot_status(ctx);
  });
}

int vlfs_direct_status(vlfs_ctx* ctx, int64_t* perturbed, int64_t* nonfinite) {
  if (!ctx || (!ctx->direct && !ctx->ilu)) return VLFS_ERR_ARG;
  long long a = 0, b = 0;
  if (ctx->ilu && ctx->precond_kind == VLFS_PRECOND_BLOCK_ILU0) { a = ctx->ilu->perturbed; b = ctx->ilu->nonfinite; }
  else if (ctx->direct) direct_pivot_status(*ctx->direct, &a, &b);
  if (perturbed) *perturbed = a;
  if (nonfinite) *nonfinite = b;
  return VLFS_OK;
}

int vlfs_direct_info(vlfs_ctx* ctx, double* out8) {
  if (!ctx || !out8 || !ctx->direct) return VLFS_ERR_ARG;
  direct_info(*ctx->direct, out8);
  return VLFS_OK;
}

int vlfs_solve(vlfs_ctx* ctx, const void* b, void* x, vlfs_solve_stats* stats) {
  return guarded(ctx, [&]() {
    if (!ctx->solver || !b || !x) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup first");
    size_t nb = (size_t)ctx->ndofs * ctx->scalar();
    DevBuf<double> db, dx;
    db.upload((const double*)b, nb, ctx->stream);
    dx.upload((const double*)x, nb, ctx->stream);       // initial guess
    vlfs_solve_stats st{};
    int rc = solver_solve(*ctx->solver, db.p, dx.p, &st);
    CUDA_TRY(cudaMemcpyAsync(x, dx.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) *stats = st;
    return rc;
  });
}

int vlfs_solve_vec(vlfs_ctx* ctx, int vec, void* x, vlfs_solve_stats* stats) {
  return guarded(ctx, [&]() {
    if (!ctx->solver || vec < 0 || vec >= ctx->nvec || !x) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup first");
    size_t nb = (size_t)ctx->ndofs * ctx->scalar();
    DevBuf<double> dx;
    dx.upload((const double*)x, nb, ctx->stream);
    vlfs_solve_stats st{};
    int rc = solver_solve(*ctx->solver, ctx->vecs[vec].p, dx.p, &st);
    CUDA_TRY(cudaMemcpyAsync(x, dx.p, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) *stats = st;
    return rc;
  });
}

int vlfs_newmark_run(vlfs_ctx* ctx, int matK, int matC, int matM, int nsteps, double dt,
                     double gamma, double beta, int nb, const double* tcoef, double* u0,
                     double* v0, double* a0, int64_t out_lo, int64_t out_hi, int out_stride,
                     double* out, vlfs_solve_stats* stats) {
  return guarded(ctx, [&]() {
    if (!ctx->solver) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup first");
    if (ctx->is_complex) return fail(ctx, VLFS_ERR_ARG, "Newmark stepping is real-valued");
    if (matK < 0 || matC < 0 || matM < 0 || matK >= ctx->nmat_all() || matC >= ctx->nmat_all() || matM >= ctx->nmat_all() ||
        nsteps < 1 || nb < 0 || nb > ctx->nvec || (nb && !tcoef) || !u0 || !v0 || !a0 ||
        out_lo < 0 || out_hi < out_lo || out_hi > ctx->ndofs || out_stride < 1 || (out_hi > out_lo && !out))
      return fail(ctx, VLFS_ERR_ARG, "vlfs_newmark_run: bad arguments");
    std::vector<double*> bv;
    for (int k = 0; k < nb; ++k) bv.push_back(ctx->vecs[k].p);
    return newmark_run(*ctx->solver, ctx->rowptr.p, ctx->colind.p, ctx->mats[matK].p, ctx->mats[matC].p, ctx->mats[matM].p,
                       ctx->ndofs, ctx->nnz, nsteps, dt, gamma, beta, nb, bv.data(), tcoef, u0, v0, a0,
                       out_lo, out_hi, out_stride, out, stats, ctx->nsm, ctx->stream, &ctx->launches);
  });
}

/* ---- distributed building blocks: device pointers in, device pointers out ------------------ */
int vlfs_set_stream(vlfs_ctx* ctx, void* stream) {
  return guarded(ctx, [&]() {
    if (ctx->solver || ctx->direct) return fail(ctx, VLFS_ERR_STATE, "vlfs_set_stream before vlfs_solver_setup");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = false;
    return VLFS_OK;
  });
}

int vlfs_set_owned_rows(vlfs_ctx* ctx, int64_t n_owned) {
  return guarded(ctx, [&]() {
    if (n_owned < 1 || n_owned > ctx->ndofs) return fail(ctx, VLFS_ERR_ARG, "vlfs_set_owned_rows: 1 <= n_owned <= ndofs");
    ctx->nowned = n_owned;
    ctx->solver.reset(); ctx->direct.reset(); ctx->ilu.reset();
    return VLFS_OK;
  });
}

int vlfs_dev_spmv(vlfs_ctx* ctx, int mat, const void* x_dev, void* y_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->pattern_built || mat < 0 || mat >= ctx->nmat_all() || !x_dev || !y_dev) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    const double one[2] = {1, 0}, zero[2] = {0, 0};
    if (ctx->dist_active()) dist_hooks(*ctx->dist)->spmv(ctx->mats[mat].p, (double*)const_cast<void*>(x_dev), (double*)y_dev, one, zero);
    else launch_spmv(ctx->rowptr.p, ctx->colind.p, ctx->mats[mat].p, x_dev, y_dev, ctx->nowned, ctx->nnz,
                     ctx->is_complex, one, zero, ctx->nsm, ctx->stream);
    ++ctx->launches;
    return VLFS_OK;
  });
}

int vlfs_dev_precond(vlfs_ctx* ctx, const void* r_dev, void* z_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->solver || !r_dev || !z_dev) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup first");
    solver_precond(*ctx->solver, (const double*)r_dev, (double*)z_dev);
    return VLFS_OK;
  });
}

int vlfs_precond_apply(vlfs_ctx* ctx, const void* r, void* z) {
  return guarded(ctx, [&]() {
    if (!ctx->solver || !r || !z) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup first");
    const size_t nb = (size_t)ctx->ndofs * ctx->scalar(), no = (size_t)ctx->nowned * ctx->scalar();
    DevBuf<double> dr, dz;
    dr.upload((const double*)r, nb, ctx->stream);
    dz.alloc(nb); dz.zero(ctx->stream);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, ctx->stream));
    solver_precond(*ctx->solver, dr.p, dz.p);
    CUDA_TRY(cudaEventRecord(e1, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(z, dz.p, no * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ctx->precond_apply_ms = ms;
    return VLFS_OK;
  });
}

int vlfs_precond_info(vlfs_ctx* ctx, double* out8) {
  if (!ctx || !out8 || !ctx->ilu) return VLFS_ERR_ARG;
  out8[0] = (double)ctx->ilu->nlevels_lower; out8[1] = (double)ctx->ilu->nlevels_upper;
  out8[2] = (double)ctx->ilu->lplan.size(); out8[3] = (double)ctx->ilu->uplan.size();
  out8[4] = ctx->ilu->symbolic_ms; out8[5] = ctx->ilu->numeric_ms; out8[6] = ctx->precond_apply_ms;
  out8[7] = (double)ctx->ilu->nnz;
  return VLFS_OK;
}

int vlfs_dev_multidot(vlfs_ctx* ctx, const void* V_dev, int64_t ldv, int nv, const void* w_dev, int64_t n,
                      void* out_host) {
  return guarded(ctx, [&]() {
    if (!V_dev || !w_dev || !out_host || nv < 1 || nv > 4096 || n < 0) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    const int nred = ctx->nsm * 2;
    const size_t sc = ctx->scalar();
    if (ctx->dev_partial.n < (size_t)nred * nv * sc) ctx->dev_partial.alloc((size_t)nred * nv * sc);
    if (ctx->dev_h.n < (size_t)nv * sc) ctx->dev_h.alloc((size_t)nv * sc);
    dev_multidot((const double*)V_dev, (size_t)ldv, nv, (const double*)w_dev, n, ctx->is_complex,
                 ctx->dev_partial.p, nred, ctx->dev_h.p, ctx->stream);
    ctx->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(out_host, ctx->dev_h.p, nv * sc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

int vlfs_dev_gemv_sub(vlfs_ctx* ctx, const void* V_dev, int64_t ldv, int nv, const void* h_host, void* w_dev,
                      int64_t n) {
  return guarded(ctx, [&]() {
    if (!V_dev || !w_dev || !h_host || nv < 1 || nv > 4096 || n < 0) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    const size_t sc = ctx->scalar();
    if (ctx->dev_h.n < (size_t)nv * sc) ctx->dev_h.alloc((size_t)nv * sc);
    CUDA_TRY(cudaMemcpyAsync(ctx->dev_h.p, h_host, nv * sc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    dev_gemv_sub((const double*)V_dev, (size_t)ldv, nv, ctx->dev_h.p, (double*)w_dev, n, ctx->is_complex,
                 ctx->nsm, ctx->stream);
    ++ctx->launches;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));     // h_host may be reused by the caller
    return VLFS_OK;
  });
}

int vlfs_dev_axpby(vlfs_ctx* ctx, double a_re, double a_im, const void* x_dev, double b_re, double b_im,
                   void* y_dev, int64_t n) {
  return guarded(ctx, [&]() {
    if (!x_dev || !y_dev || n < 0) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    dev_axpby(a_re, a_im, (const double*)x_dev, b_re, b_im, (double*)y_dev, n, ctx->is_complex, ctx->nsm, ctx->stream);
    ++ctx->launches;
    return VLFS_OK;
  });
}

int vlfs_dev_gather(vlfs_ctx* ctx, int64_t n, const int32_t* idx_dev, const void* src_dev, void* dst_dev) {
  return guarded(ctx, [&]() {
    if (n < 0 || (n && (!idx_dev || !src_dev || !dst_dev))) return fail(ctx, VLFS_ERR_ARG, "bad arguments");
    dev_gather(idx_dev, (const double*)src_dev, (double*)dst_dev, n, ctx->is_complex, ctx->nsm, ctx->stream);
    ++ctx->launches;
    return VLFS_OK;
  });
}

int vlfs_get_vector_device(vlfs_ctx* ctx, int vec, void** ptr) {
  if (!ctx || vec < 0 || vec >= ctx->nvec || !ptr) return VLFS_ERR_ARG;
  *ptr = ctx->vecs[vec].p;
  return VLFS_OK;
}

int vlfs_eval_functional(vlfs_ctx* ctx, int domain_id, const vlfs_functional_desc* f, const void* x_host,
                         double* value) {
  return guarded(ctx, [&]() {
    if (!f || !x_host || !value || domain_id < 0 || domain_id >= (int)ctx->doms.size())
      return fail(ctx, VLFS_ERR_ARG, "vlfs_eval_functional: bad arguments");
    if (ctx->is_complex) return fail(ctx, VLFS_ERR_ARG, "vlfs_eval_functional: real systems only");
    Domain& d = *ctx->doms[domain_id];
    const DomainDev& dev = d.dev;
    const int nc = op_ncomp(f->op, dev.D);
    if (nc < 0 || f->weight_f >= dev.nweights || (f->weight_f >= 0 && nc != 1))
      return fail(ctx, VLFS_ERR_ARG, "vlfs_eval_functional: bad operator / field");
    OpInst O[2];
    double sc[2] = {0.0, 0.0};
    for (int s = 0; s < 2; ++s) {
      O[s].slot = s < dev.nslots ? s : 0; O[s].op = f->op; O[s].ncomp = nc; O[s].smem_off = 0; O[s].ndp = 1;
      O[s].buf_stride = 0;
      for (int k = 0; k < 3; ++k) O[s].dir[k] = f->dir[k];
      if (s < dev.nslots) sc[s] = f->scale[s];
      if (sc[s] != 0.0) {
        if (op_needs_hess(f->op) && !dev.slot[s].hess) return fail(ctx, VLFS_ERR_ARG, "hess table missing");
        if (f->op == VLFS_OP_GRAD_NOWN && !dev.slot[s].ref_normal) return fail(ctx, VLFS_ERR_ARG, "ref_normal missing");
        if ((f->op == VLFS_OP_CHESS || f->op == VLFS_OP_CHESS_NPLUS) && !dev.C) return fail(ctx, VLFS_ERR_ARG, "C missing");
      }
    }
    DevBuf<double> dx, res;
    dx.upload((const double*)x_host, (size_t)ctx->ndofs, ctx->stream);
    res.alloc((size_t)ctx->nsm * 8 + 1); res.zero(ctx->stream);   // [0] the value, then one partial per CTA
    launch_functional(dev, O[0], O[1], sc[0], sc[1], f->weight_f, dx.p, res.p, ctx->nsm, ctx->stream);
    ctx->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(value, res.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VLFS_OK;
  });
}

/* ---- substructuring: partial factorisation with a kept boundary + dense interface solve -------- */
int vlfs_set_roles(vlfs_ctx* ctx, const int8_t* role) {
  return guarded(ctx, [&]() {
    if (!role || ctx->ndofs == 0) return fail(ctx, VLFS_ERR_ARG, "vlfs_set_roles: bad arguments");
    for (long long i = 0; i < ctx->ndofs; ++i)
      if (role[i] < 0 || role[i] > 2) return fail(ctx, VLFS_ERR_ARG, "vlfs_set_roles: role must be 0, 1 or 2");
    ctx->roles.assign(role, role + ctx->ndofs);
    ctx->solver.reset(); ctx->direct.reset(); ctx->ilu.reset();
    return VLFS_OK;
  });
}

int vlfs_schur_size(vlfs_ctx* ctx, int64_t* n_keep) {
  if (!ctx || !n_keep || !ctx->direct) return VLFS_ERR_ARG;
  *n_keep = direct_nkeep(*ctx->direct);
  return VLFS_OK;
}

int vlfs_schur_get(vlfs_ctx* ctx, void* S_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->direct || !S_dev) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup (sparse LU) first");
    direct_schur(*ctx->direct, (double*)S_dev);
    return VLFS_OK;
  });
}

int vlfs_schur_forward(vlfs_ctx* ctx, const void* b_dev, void* g_keep_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->direct || !b_dev) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup (sparse LU) first");
    direct_forward(*ctx->direct, (const double*)b_dev, (double*)g_keep_dev);
    return VLFS_OK;
  });
}

int vlfs_schur_backward(vlfs_ctx* ctx, const void* x_keep_dev, void* x_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->direct || !x_dev) return fail(ctx, VLFS_ERR_STATE, "vlfs_solver_setup (sparse LU) first");
    direct_backward(*ctx->direct, (const double*)x_keep_dev, (double*)x_dev);
    return VLFS_OK;
  });
}

int vlfs_dense_factor(vlfs_ctx* ctx, int64_t n, const void* S_dev) {
  return guarded(ctx, [&]() {
    if (n < 1 || n > 200000 || !S_dev) return fail(ctx, VLFS_ERR_ARG, "vlfs_dense_factor: bad arguments");
    ctx->dense = direct_dense(n, ctx->is_complex, ctx->nsm, ctx->stream, &ctx->launches);
    direct_dense_factor(*ctx->dense, (const double*)S_dev);
    return VLFS_OK;
  });
}

int vlfs_chain_factor(vlfs_ctx* ctx, int nblocks, const int64_t* sizes, const void* S_dev) {
  return guarded(ctx, [&]() {
    if (nblocks < 1 || nblocks > 4096 || !sizes || !S_dev) return fail(ctx, VLFS_ERR_ARG, "vlfs_chain_factor: bad arguments");
    std::vector<long long> sz(sizes, sizes + nblocks);
    for (long long v : sz) if (v < 1) return fail(ctx, VLFS_ERR_ARG, "vlfs_chain_factor: empty block");
    ctx->dense = direct_chain(nblocks, sz.data(), ctx->is_complex, ctx->nsm, ctx->stream, &ctx->launches);
    direct_chain_factor(*ctx->dense, (const double*)S_dev);
    return VLFS_OK;
  });
}

/* ---- multi-GPU: communicator, halo plan, collective calls ------------------------------------ */
int vlfs_dist_unique_id(void* id128) {
  if (!id128) return VLFS_ERR_ARG;
  try { dist_unique_id(id128); } catch (...) { return VLFS_ERR_NCCL; }
  return VLFS_OK;
}

int vlfs_dist_init(vlfs_ctx* ctx, int rank, int nranks, const void* id128) {
  return guarded(ctx, [&]() {
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !id128)) return fail(ctx, VLFS_ERR_ARG, "vlfs_dist_init: bad rank / id");
    if (ctx->solver || ctx->direct) return fail(ctx, VLFS_ERR_STATE, "vlfs_dist_init before vlfs_solver_setup");
    ctx->dist = dist_init_rank(rank, nranks, id128, ctx->device);
    dist_bind(*ctx->dist, ctx->stream, ctx->is_complex, ctx->nsm, &ctx->launches);
    return VLFS_OK;
  });
}

int vlfs_multi_create(vlfs_ctx** ctxs, int ngpu, const int* devices) {
  if (!ctxs || ngpu < 1 || ngpu > 64 || !devices) return VLFS_ERR_ARG;
  for (int r = 0; r < ngpu; ++r) ctxs[r] = nullptr;
  int rc = VLFS_OK;
  for (int r = 0; r < ngpu && rc == VLFS_OK; ++r) rc = vlfs_create(&ctxs[r], devices[r]);
  if (rc == VLFS_OK) {
    try {
      std::vector<DistPtr> d;
      dist_init_all(ngpu, devices, d);
      for (int r = 0; r < ngpu; ++r) ctxs[r]->dist = std::move(d[r]);
    } catch (const NcclError& e) { ctxs[0]->err = e.msg; rc = VLFS_ERR_NCCL; }
    catch (...) { rc = VLFS_ERR_NCCL; }
  }
  if (rc != VLFS_OK) for (int r = 0; r < ngpu; ++r) if (ctxs[r]) { vlfs_destroy(ctxs[r]); ctxs[r] = nullptr; }
  return rc;
}

int vlfs_dist_set_halo(vlfs_ctx* ctx, int64_t n_owned, int npeers, const int32_t* peers, const int64_t* send_ptr,
                       const int32_t* send_idx, const int64_t* recv_ptr) {
  return guarded(ctx, [&]() {
    if (!ctx->dist) return fail(ctx, VLFS_ERR_STATE, "vlfs_dist_init / vlfs_multi_create first");
    if (!ctx->pattern_built) return fail(ctx, VLFS_ERR_STATE, "vlfs_build_pattern first");
    if (n_owned < 1 || n_owned > ctx->ndofs || npeers < 0 || (npeers && (!peers || !send_ptr || !recv_ptr)))
      return fail(ctx, VLFS_ERR_ARG, "vlfs_dist_set_halo: bad arguments");
    ctx->nowned = n_owned;
    ctx->solver.reset(); ctx->direct.reset(); ctx->ilu.reset(); ctx->roles.clear();
    dist_bind(*ctx->dist, ctx->stream, ctx->is_complex, ctx->nsm, &ctx->launches);
    std::vector<long long> sp(npeers + 1, 0), rp(npeers + 1, 0);
    for (int p = 0; p <= npeers && npeers > 0; ++p) { sp[p] = send_ptr[p]; rp[p] = recv_ptr[p]; }
    dist_set_halo(*ctx->dist, n_owned, ctx->ndofs, npeers, peers, sp.data(), send_idx, rp.data(),
                  ctx->rowptr.p, ctx->colind.p, ctx->nnz);
    return VLFS_OK;
  });
}

int vlfs_dist_info(vlfs_ctx* ctx, double* out4) {
  if (!ctx || !ctx->dist || !out4) return VLFS_ERR_ARG;
  dist_schur_info(*ctx->dist, out4);
  return VLFS_OK;
}


int vlfs_multi_build_pattern(vlfs_ctx** ctxs, int n, int64_t* nnz, int64_t* ncoo) {
  return run_ranks(n, ctxs, [&](int r, vlfs_ctx* c) { return vlfs_build_pattern(c, nnz ? nnz + r : nullptr, ncoo ? ncoo + r : nullptr); });
}
int vlfs_multi_assemble(vlfs_ctx** ctxs, int n) {
  return run_ranks(n, ctxs, [&](int, vlfs_ctx* c) { return vlfs_assemble(c); });
}
int vlfs_multi_solver_setup(vlfs_ctx** ctxs, int n, int mat, const vlfs_solver_opts* opts) {
  return run_ranks(n, ctxs, [&](int, vlfs_ctx* c) { return vlfs_solver_setup(c, mat, opts); });
}
int vlfs_multi_solver_refactor(vlfs_ctx** ctxs, int n, int mat) {
  return run_ranks(n, ctxs, [&](int, vlfs_ctx* c) { return vlfs_solver_refactor(c, mat); });
}
int vlfs_multi_solve_vec(vlfs_ctx** ctxs, int n, int vec, void** x_hosts, vlfs_solve_stats* stats) {
  if (!x_hosts) return VLFS_ERR_ARG;
  return run_ranks(n, ctxs, [&](int r, vlfs_ctx* c) { return vlfs_solve_vec(c, vec, x_hosts[r], stats ? stats + r : nullptr); });
}
int vlfs_multi_spmv(vlfs_ctx** ctxs, int n, int mat, const void* const* x_hosts, void** y_hosts) {
  if (!x_hosts || !y_hosts) return VLFS_ERR_ARG;
  return run_ranks(n, ctxs, [&](int r, vlfs_ctx* c) { return vlfs_spmv(c, mat, x_hosts[r], y_hosts[r]); });
}
int vlfs_multi_spmv_device(vlfs_ctx** ctxs, int n, int mat, int reps, float* ms_total) {
  return run_ranks(n, ctxs, [&](int r, vlfs_ctx* c) { return vlfs_spmv_device(c, mat, reps, ms_total ? ms_total + r : nullptr); });
}

int vlfs_dense_solve(vlfs_ctx* ctx, const void* b_dev, void* x_dev) {
  return guarded(ctx, [&]() {
    if (!ctx->dense || !b_dev || !x_dev) return fail(ctx, VLFS_ERR_STATE, "vlfs_dense_factor first");
    direct_solve(*ctx->dense, (const double*)b_dev, (double*)x_dev);
    return VLFS_OK;
  });
}

}  // extern "C"
