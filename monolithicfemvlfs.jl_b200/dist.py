"""Row-partitioned (multi-GPU) assembly and solve: SURVEY.md §8(e).

The reference is a serial Julia program; this module is the partitioned twin of its
`AffineFEOperator` + `solve` path (src/Yago_freq_domain.jl:167-171, src/Multi_geo_freq_domain.jl:131-135):

* DOFs are split into slabs along one coordinate axis (owner-computes rows).  A rank registers
  with its own libvlfs_b200 context only the cells that touch one of its DOFs, with local DOF
  ids [owned | ghost]; every owned row of the local CSR is then complete - no matrix entry ever
  crosses GPUs and nothing is reduced between ranks during assembly.
* y = A x needs the ghost values of x: one halo exchange (pack kernel + NCCL send/recv through
  torch.distributed) before the local SpMV over the owned rows.
* The solve is right-preconditioned GMRES (CGS2, all dots of a pass batched into one all-reduce)
  with the GPU multifrontal LU of each rank's owned block as block-Jacobi preconditioner.

`Fleet` runs any number of ranks inside one process in lockstep (all members local: used by the
1-GPU tests, no waiting between kernels of different ranks) or one member per process with
torch.distributed carrying halos and reductions (real multi-GPU).  Arithmetic goes through a
`LocalOps` object: `GpuOps` below (libvlfs_b200 device entry points on torch CUDA tensors); the
NumPy twin used by the gloo CPU tests lives in oracle/dist_cpu.py and is test infrastructure.
"""
import ctypes as C
import numpy as np


# ---------------------------------------------------------------------------------------------
# host logic: ownership, local numbering, recorded problem definition
# ---------------------------------------------------------------------------------------------
def slab_owner(x, nranks):
    """Owner rank of every DOF: `nranks` slabs of (nearly) equal DOF count along coordinate x;
    DOFs with equal coordinate stay together (ties go to the lower rank)."""
    x = np.asarray(x, dtype=np.float64)
    if nranks == 1:
        return np.zeros(len(x), dtype=np.int32)
    xs = np.sort(x)
    cuts = xs[(np.arange(1, nranks) * len(x)) // nranks]
    return np.searchsorted(cuts, x, side="left").astype(np.int32)


class _DomainProxy:
    """Records what a case module registers on a domain; forwards coefficient updates to the
    rank-local domain once it exists."""

    def __init__(self, system, kwargs):
        self.system, self.kwargs = system, kwargs
        self.terms, self.rhs, self.local, self.cells = [], [], None, None
        self.ncells = len(kwargs["slots"][0]["dofs"])
        self.nq = kwargs["nq"]
        self.weights = [np.asarray(w, dtype=np.float64).reshape(self.ncells, self.nq).copy()
                        for w in kwargs.get("weights", ())]

    def add_term(self, *a, **k):
        self.terms.append((a, k))
        return len(self.terms) - 1

    def add_rhs(self, *a, **k):
        self.rhs.append((a, k))
        return len(self.rhs) - 1

    def set_coef(self, term, coef):
        a, k = self.terms[term]
        self.terms[term] = ((a[0], coef) + tuple(a[2:]), k)
        if self.local is not None:
            self.local.set_coef(term, coef)

    def set_rhs_coef(self, term, coef):
        a, k = self.rhs[term]
        self.rhs[term] = ((a[0], coef) + tuple(a[2:]), k)
        if self.local is not None:
            self.local.set_rhs_coef(term, coef)

    def set_weights(self, index, values):
        self.weights[index] = np.asarray(values, dtype=np.float64).reshape(self.ncells, self.nq).copy()
        if self.local is not None:
            self.local.set_weights(index, self.weights[index][self.cells])


class DistributedSystem:
    """Same surface as `B200System` for the case modules; builds the rank-local system."""

    def __init__(self, ndofs, is_complex, nmat=1, nvec=1, device=0, rank=0, nranks=1, axis=0,
                 backend=None):
        if backend is None:
            from .assembler import B200System as backend
        self.backend, self.device = backend, device
        self.ndofs_global, self.is_complex, self.nmat, self.nvec = int(ndofs), bool(is_complex), nmat, nvec
        self.rank, self.nranks, self.axis = rank, nranks, axis
        self.dtype = np.complex128 if is_complex else np.float64
        self.domains, self.coords, self.local = [], None, None

    def add_domain(self, D, nq, qweights, slots, **kw):
        kw = dict(kw, D=D, nq=nq, qweights=qweights, slots=slots)
        d = _DomainProxy(self, kw)
        self.domains.append(d)
        return d

    def set_dof_coords(self, coords):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)

    def build_pattern(self):
        if self.coords is None:
            raise ValueError("DistributedSystem needs set_dof_coords before build_pattern")
        N = self.ndofs_global
        self.owner = owner = slab_owner(self.coords[:, self.axis], self.nranks)
        mine = owner == self.rank
        self.owned = np.nonzero(mine)[0]
        touched = np.zeros(N, dtype=bool)
        for d in self.domains:
            keep = np.zeros(d.ncells, dtype=bool)
            for s in d.kwargs["slots"]:
                keep |= mine[np.asarray(s["dofs"])].any(axis=1)
            d.cells = np.nonzero(keep)[0]
            for s in d.kwargs["slots"]:
                touched[np.asarray(s["dofs"])[d.cells].reshape(-1)] = True
        gh = np.nonzero(touched & ~mine)[0]
        # ghost block grouped by owner (then ascending id): every peer's halo is contiguous
        self.ghosts = gh[np.lexsort((gh, owner[gh]))]
        self.n_owned, self.n_ghost = len(self.owned), len(self.ghosts)
        self.n_local = self.n_owned + self.n_ghost
        g2l = np.full(N, -1, dtype=np.int64)
        g2l[self.owned] = np.arange(self.n_owned)
        g2l[self.ghosts] = self.n_owned + np.arange(self.n_ghost)
        self.g2l = g2l
        self.local_gids = np.concatenate([self.owned, self.ghosts])
        loc = self.local = self.backend(self.n_local, self.is_complex, nmat=self.nmat, nvec=self.nvec,
                                        device=self.device)
        loc.set_owned_rows(self.n_owned)
        for d in self.domains:
            kw = dict(d.kwargs)
            slots = []
            for s in kw.pop("slots"):
                t = dict(s)
                t["coords"] = np.asarray(s["coords"])[d.cells]
                t["dofs"] = g2l[np.asarray(s["dofs"])[d.cells]]
                if s.get("variant") is not None:
                    t["variant"] = np.asarray(s["variant"])[d.cells]
                slots.append(t)
            kw.pop("weights", None)
            D_, nq, qw = kw.pop("D"), kw.pop("nq"), kw.pop("qweights")
            d.local = loc.add_domain(D_, nq, qw, slots, weights=[w[d.cells] for w in d.weights], **kw)
            for a, k in d.terms:
                d.local.add_term(*a, **k)
            for a, k in d.rhs:
                d.local.add_rhs(*a, **k)
        loc.set_dof_coords(self.coords[self.local_gids])
        self.nnz = loc.build_pattern()
        self.ncoo = getattr(loc, "ncoo", None)
        return self.nnz

    def assemble(self, *a, **k):
        self.local.assemble(*a, **k)

    # global <-> local vectors (host)
    def to_local(self, xg):
        return np.ascontiguousarray(np.asarray(xg, dtype=self.dtype)[self.local_gids])

    def scatter_owned(self, xl, out):
        out[self.owned] = np.asarray(xl)[:self.n_owned]


def halo_plan(owner, rank, ghosts):
    """{peer: ghost global ids owned by peer}, in the order of the (owner-grouped) ghost block."""
    own = owner[ghosts]
    return {int(p): ghosts[own == p] for p in np.unique(own)}


# ---------------------------------------------------------------------------------------------
# device arithmetic of one rank (libvlfs_b200 vlfs_dev_* on torch CUDA tensors)
# ---------------------------------------------------------------------------------------------
class GpuOps:
    def __init__(self, dsys, mat=0):
        import torch
        self.torch = torch
        self.d, self.sys, self.mat = dsys, dsys.local, mat
        self.sc = 2 if dsys.is_complex else 1
        self.device = torch.device("cuda", dsys.device)
        self.n_owned, self.n_local = dsys.n_owned, dsys.n_local
        self.lib, self.ctx = self.sys.lib, self.sys.ctx

    def bind_stream(self):
        """Run the library on torch's current stream (call before solver_setup)."""
        s = self.torch.cuda.current_stream(self.device).cuda_stream
        self.sys._check(self.lib.vlfs_set_stream(self.ctx, C.c_void_p(s)))

    def zeros(self, rows=None):
        shape = (self.n_local * self.sc,) if rows is None else (rows, self.n_local * self.sc)
        return self.torch.zeros(shape, dtype=self.torch.float64, device=self.device)

    def from_host(self, x):
        x = np.ascontiguousarray(x, dtype=self.d.dtype)
        return self.torch.from_numpy(x.view(np.float64).copy()).to(self.device)

    def to_host(self, t):
        a = t.detach().cpu().numpy()
        return a.view(np.complex128) if self.d.is_complex else a

    def index_tensor(self, idx):
        return self.torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int32)).to(self.device)

    def _p(self, t):
        return C.c_void_p(t.data_ptr())

    def spmv(self, x, y):
        self.sys._check(self.lib.vlfs_dev_spmv(self.ctx, self.mat, self._p(x), self._p(y)))

    def precond(self, r, z):
        self.sys._check(self.lib.vlfs_dev_precond(self.ctx, self._p(r), self._p(z)))

    def multidot(self, V, nv, w):
        out = np.zeros(nv, dtype=self.d.dtype)
        self.sys._check(self.lib.vlfs_dev_multidot(self.ctx, self._p(V), self.n_local, nv, self._p(w),
                                                   self.n_owned, out.ctypes.data_as(C.c_void_p)))
        return out

    def gemv_sub(self, V, nv, h, w):
        h = np.ascontiguousarray(h, dtype=self.d.dtype)
        self.sys._check(self.lib.vlfs_dev_gemv_sub(self.ctx, self._p(V), self.n_local, nv,
                                                   h.ctypes.data_as(C.c_void_p), self._p(w), self.n_owned))

    def axpby(self, a, x, b, y):
        a, b = complex(a), complex(b)
        self.sys._check(self.lib.vlfs_dev_axpby(self.ctx, a.real, a.imag, self._p(x), b.real, b.imag,
                                                self._p(y), self.n_owned))

    def gather(self, idx_t, x):
        buf = self.torch.empty(len(idx_t) * self.sc, dtype=self.torch.float64, device=self.device)
        self.sys._check(self.lib.vlfs_dev_gather(self.ctx, len(idx_t), self._p(idx_t), self._p(x), self._p(buf)))
        return buf

    def ghost_view(self, x, offset, count):
        lo = (self.n_owned + offset) * self.sc
        return x[lo:lo + count * self.sc]

    def rhs_vector(self, vec=0):
        """Assembled right-hand side of the local system as a device tensor."""
        return self.from_host(self.sys.get_vector(vec))

    def set_ghost(self, x, offset, buf):
        self.ghost_view(x, offset, len(buf) // self.sc).copy_(buf)

    def as_tensor(self, t):
        return t

    def fill_zero(self, x):
        x.zero_()

    # -- substructuring (partial factorisation with a kept boundary) -------------------------
    def set_roles(self, role):
        r = np.ascontiguousarray(role, dtype=np.int8)
        self.sys._check(self.lib.vlfs_set_roles(self.ctx, r.ctypes.data_as(C.POINTER(C.c_int8))))

    def factor(self, **opts):
        from . import _lib as L
        self.sys.solver_setup(self.mat, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=2, maxit=2,
                              rtol=1e-12, **opts)

    def refactor(self):
        self.sys.refactor(self.mat)

    def schur_block(self, nk):
        S = self.torch.zeros(nk * nk * self.sc, dtype=self.torch.float64, device=self.device)
        self.sys._check(self.lib.vlfs_schur_get(self.ctx, self._p(S)))
        return S

    def schur_forward(self, b, g):
        self.sys._check(self.lib.vlfs_schur_forward(self.ctx, self._p(b), self._p(g)))

    def schur_backward(self, xk, x):
        self.sys._check(self.lib.vlfs_schur_backward(self.ctx, self._p(xk), self._p(x)))

    def dense_factor(self, n, S):
        self.sys._check(self.lib.vlfs_dense_factor(self.ctx, n, self._p(S)))

    def chain_factor(self, sizes, S):
        sz = np.ascontiguousarray(sizes, dtype=np.int64)
        self.sys._check(self.lib.vlfs_chain_factor(self.ctx, len(sz), sz.ctypes.data_as(C.POINTER(C.c_int64)), self._p(S)))

    def dense_solve(self, b, x):
        self.sys._check(self.lib.vlfs_dense_solve(self.ctx, self._p(b), self._p(x)))

    def matrix_values(self, slots_t):
        """Values of the local matrix at the given CSR slots (int32 device tensor), gathered on the
        device: the interface rows never travel through the host."""
        rp, ci, vp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self.sys._check(self.lib.vlfs_device_ptrs(self.ctx, self.mat, C.byref(rp), C.byref(ci), C.byref(vp)))
        buf = self.torch.empty(len(slots_t) * self.sc, dtype=self.torch.float64, device=self.device)
        if len(slots_t):
            self.sys._check(self.lib.vlfs_dev_gather(self.ctx, len(slots_t), self._p(slots_t), vp, self._p(buf)))
        return buf

    def add_triplets(self, S, n, rows_t, cols_t, vals):
        """S[rows, cols] += vals on the device (flat row-major n x n; index tensors from long_index,
        vals a flat device tensor)."""
        S.view(n, n, self.sc).index_put_((rows_t, cols_t), vals.view(-1, self.sc), accumulate=True)

    def vec(self, n):
        return self.torch.zeros(n * self.sc, dtype=self.torch.float64, device=self.device)

    def dense_from_host(self, a):
        """complex/real 2-D NumPy array -> flat device tensor (row-major, interleaved)."""
        a = np.ascontiguousarray(a, dtype=self.d.dtype)
        return self.torch.from_numpy(a.reshape(-1).view(np.float64).copy()).to(self.device)

    def add_block(self, S, n, idx, block, nb):
        """S[idx,idx] += block for flat row-major S (n x n) and block (nb x nb)."""
        t = self.torch
        idx_t = t.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(self.device)
        Sv = S.view(n, n, self.sc)
        Sv.index_put_((idx_t[:, None], idx_t[None, :]), block.view(nb, nb, self.sc), accumulate=True)

    def take(self, x, idx_t):
        """entries idx of an interleaved vector -> new flat tensor."""
        return x.view(-1, self.sc)[idx_t].reshape(-1).contiguous()

    def put(self, x, idx_t, vals, accumulate=False):
        xv = x.view(-1, self.sc)
        if accumulate:
            xv.index_add_(0, idx_t, vals.view(-1, self.sc))
        else:
            xv[idx_t] = vals.view(-1, self.sc)

    def long_index(self, idx):
        return self.torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(self.device)

    def to_index_host(self, idx_t):
        return idx_t.cpu().numpy()

    def empty_buf(self, count):
        return self.torch.empty(count * self.sc, dtype=self.torch.float64, device=self.device)

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()


# ---------------------------------------------------------------------------------------------
# a group of ranks: all local (lockstep emulation) or one local + torch.distributed
# ---------------------------------------------------------------------------------------------
class Member:
    def __init__(self, dsys, ops):
        self.d, self.ops = dsys, ops
        self.rank = dsys.rank
        self.recv = {}      # peer -> (offset in ghost block, count)
        self.send = {}      # peer -> device index tensor of owned local ids


class Fleet:
    def __init__(self, members, nranks, dist=None):
        """members: local Member objects; dist: the torch.distributed module when the other
        ranks live in other processes (then len(members) == 1)."""
        self.members, self.nranks, self.dist = members, nranks, dist
        self.by_rank = {m.rank: m for m in members}
        plans = {m.rank: halo_plan(m.d.owner, m.rank, m.d.ghosts) for m in members}
        if dist is not None:
            gathered = [None] * nranks
            dist.all_gather_object(gathered, plans[members[0].rank])
            allplans = {r: gathered[r] for r in range(nranks)}
        else:
            allplans = plans
        for m in members:
            off = 0
            for p in sorted(plans[m.rank]):
                cnt = len(plans[m.rank][p])
                m.recv[p] = (off, cnt)
                off += cnt
            for r in range(nranks):                 # what rank r wants from me
                want = allplans[r].get(m.rank)
                if r != m.rank and want is not None and len(want):
                    m.send[r] = m.ops.index_tensor(m.d.g2l[want])
        self.halo_bytes = sum(len(v) for m in members for v in m.send.values()) * 8 * (2 if members[0].d.is_complex else 1)

    # -- collectives ------------------------------------------------------------------------
    def allreduce(self, parts):
        """parts: one NumPy array per local member -> the global sum (same on every rank)."""
        tot = np.sum(np.stack(parts), axis=0)
        if self.dist is not None:
            import torch
            m = self.members[0]
            dev = getattr(m.ops, "device", "cpu")
            t = torch.from_numpy(np.ascontiguousarray(tot).view(np.float64).copy()).to(dev)
            self.dist.all_reduce(t)
            tot = t.cpu().numpy().view(tot.dtype).reshape(tot.shape)
        return tot

    def exchange(self, xs):
        """Fill the ghost block of every member's vector from the owners."""
        if self.dist is None:
            for m, x in zip(self.members, xs):
                for peer, idx in m.send.items():
                    dst = self.by_rank[peer]
                    off, cnt = dst.recv[m.rank]
                    dst.ops.set_ghost(xs[self.members.index(dst)], off, m.ops.gather(idx, x))
            return
        m, x = self.members[0], xs[0]
        ops = []
        keep = []
        for peer, idx in sorted(m.send.items()):
            buf = m.ops.gather(idx, x)
            keep.append(buf)
            ops.append(self.dist.P2POp(self.dist.isend, m.ops.as_tensor(buf), peer))
        for peer, (off, cnt) in sorted(m.recv.items()):
            ops.append(self.dist.P2POp(self.dist.irecv, m.ops.as_tensor(m.ops.ghost_view(x, off, cnt)), peer))
        if ops:
            for r in self.dist.batch_isend_irecv(ops):
                r.wait()

    # -- distributed operators ------------------------------------------------------------------
    def matvec(self, xs, ys):
        self.exchange(xs)
        for m, x, y in zip(self.members, xs, ys):
            m.ops.spmv(x, y)

    def dots(self, Vs, nv, ws):
        return self.allreduce([m.ops.multidot(V, nv, w) for m, V, w in zip(self.members, Vs, ws)])

    def norm(self, xs):
        return float(np.sqrt(abs(self.dots(xs, 1, xs)[0].real)))

    def gmres(self, bs, xs, restart=50, maxit=500, rtol=1e-10, precond=True, apply=None, stagnation=None):
        """Right-preconditioned restarted GMRES with CGS2 on distributed vectors (lists with one
        device vector per local member).  Returns dict(iterations, rel_residual, converged)."""
        M = self.members
        m_ = restart
        V = [mm.ops.zeros(m_ + 1) for mm in M]
        w = [mm.ops.zeros() for mm in M]
        z = [mm.ops.zeros() for mm in M]
        r = [mm.ops.zeros() for mm in M]
        row = lambda Vk, j: Vk[j]
        bnorm = self.norm(bs)
        if bnorm == 0.0:
            for mm, x in zip(M, xs):
                mm.ops.fill_zero(x)
            return dict(iterations=0, rel_residual=0.0, converged=True)
        it, rel = 0, 1.0
        dtype = M[0].d.dtype
        # stagnation=f: with an exact preconditioner every iteration is a refinement step; stop
        # when one gains less than a factor 1/f (attainable accuracy eps*|A||x|/|b| reached)
        stalled, prev_cycle = False, np.inf
        while True:
            self.matvec(xs, r)                                   # r = b - A x
            for mm, rr, b in zip(M, r, bs):
                mm.ops.axpby(1.0, b, -1.0, rr)
            beta = self.norm(r)
            rel = beta / bnorm
            if rel <= rtol or it >= maxit or stalled:
                break
            if stagnation and rel > stagnation * prev_cycle:
                stalled = True                                   # a whole cycle gained less than 1/f
                break
            prev_cycle = rel
            for mm, Vk, rr in zip(M, V, r):
                mm.ops.axpby(1.0 / beta, rr, 0.0, row(Vk, 0))
            H = np.zeros((m_ + 1, m_), dtype=dtype)
            cs = np.zeros(m_, dtype=dtype); sn = np.zeros(m_, dtype=dtype)
            g = np.zeros(m_ + 1, dtype=dtype); g[0] = beta
            j = 0
            while j < m_ and it < maxit:
                if apply is not None:                      # e.g. SchurSolver.apply (exact)
                    apply([row(Vk, j) for Vk in V], z)
                else:
                    for mm, Vk, zz in zip(M, V, z):
                        if precond:
                            mm.ops.precond(row(Vk, j), zz)
                        else:
                            mm.ops.axpby(1.0, row(Vk, j), 0.0, zz)
                self.matvec(z, w)
                h = self.dots(V, j + 1, w)
                for mm, Vk, ww in zip(M, V, w):
                    mm.ops.gemv_sub(Vk, j + 1, h, ww)
                h2 = self.dots(V, j + 1, w)
                for mm, Vk, ww in zip(M, V, w):
                    mm.ops.gemv_sub(Vk, j + 1, h2, ww)
                h = h + h2
                hn = self.norm(w)
                if hn > 0:
                    for mm, Vk, ww in zip(M, V, w):
                        mm.ops.axpby(1.0 / hn, ww, 0.0, row(Vk, j + 1))
                col = np.concatenate([h, [hn]]).astype(dtype)
                for k in range(j):
                    t = cs[k] * col[k] + sn[k] * col[k + 1]
                    col[k + 1] = -np.conj(sn[k]) * col[k] + cs[k] * col[k + 1]
                    col[k] = t
                a, bb = col[j], col[j + 1]
                na, nb = abs(a), abs(bb)
                t = np.hypot(na, nb)
                if na == 0.0:
                    cs[j], sn[j] = 0.0, 1.0
                else:
                    cs[j], sn[j] = na / t, (a / na) * np.conj(bb) / t
                col[j] = cs[j] * a + sn[j] * bb
                col[j + 1] = 0.0
                g[j + 1] = -np.conj(sn[j]) * g[j]
                g[j] = cs[j] * g[j]
                H[:j + 2, j] = col[:j + 2]
                prev = rel
                rel = abs(g[j + 1]) / bnorm
                j += 1
                it += 1
                if rel <= rtol or hn == 0.0:
                    break
                if stagnation and j >= 2 and rel > stagnation * prev:
                    stalled = True
                    break
            y = np.zeros(j, dtype=dtype)
            for k in range(j - 1, -1, -1):
                y[k] = (g[k] - H[k, k + 1:j] @ y[k + 1:j]) / H[k, k]
            for mm, Vk, ww, zz, x in zip(M, V, w, z, xs):        # x += M^-1 V y
                mm.ops.fill_zero(ww)
                mm.ops.gemv_sub(Vk, j, -y, ww)
            if apply is not None:
                apply(w, z)
            for mm, ww, zz, x in zip(M, w, z, xs):
                if apply is None:
                    if precond:
                        mm.ops.precond(ww, zz)
                    else:
                        mm.ops.axpby(1.0, ww, 0.0, zz)
                mm.ops.axpby(1.0, zz, 1.0, x)
        return dict(iterations=it, rel_residual=rel, converged=rel <= rtol, stagnated=bool(stalled))


# ---------------------------------------------------------------------------------------------
# exact distributed solve: substructuring on the slab interfaces
# ---------------------------------------------------------------------------------------------
class SchurSolver:
    """x = A^-1 b for the row-partitioned system, used as the (exact) preconditioner of
    `Fleet.gmres` so that one or two iterations give the LU-quality solution:

      * interface Γ = owned unknowns that a LOWER rank holds as ghosts (the first DOF layer of
        every slab); the remaining owned unknowns I_r of different ranks do not couple;
      * every rank factorises A[I_r,I_r] with the multifrontal LU keeping B_r = Γ_r ∪ (ghosts owned
        by higher ranks) as a Schur boundary: the root fronts end with -A[B,I] A[I,I]^-1 A[I,B];
      * rank 0 assembles S = A[Γ,Γ] + Σ_r blocks (a few thousand unknowns per cut), factorises it
        densely with the same front kernels, and per application solves S x_Γ = b_Γ + Σ_r g_r with
        the forward sweeps' boundary parts g_r; the backward sweeps finish x_I on every rank.
    """

    def __init__(self, fleet, mat=0):
        self.fleet, self.mat = fleet, mat
        F = fleet
        # Γ_r: owned unknowns requested by lower ranks
        infos = []
        for m in F.members:
            d = m.d
            gam = np.zeros(d.n_local, dtype=bool)
            for peer, idx in m.send.items():
                if peer < m.rank:
                    gam[np.asarray(m.ops.to_index_host(idx))] = True
            role = np.full(d.n_local, 2, dtype=np.int8)
            role[:d.n_owned] = 0
            role[gam] = 1
            higher = d.owner[d.ghosts] > m.rank
            role[d.n_owned:][higher] = 1
            m.role = role
            m.keep_local = np.nonzero(role == 1)[0]                 # ascending local id
            m.keep_gids = d.local_gids[m.keep_local]
            m.gamma_local = np.nonzero(gam)[0]
            infos.append((m.rank, d.local_gids[m.gamma_local]))
        if F.dist is not None:
            gathered = [None] * F.nranks
            F.dist.all_gather_object(gathered, infos[0])
            infos = gathered
        # interface unknowns grouped by cut (= owner rank), ascending global id inside a cut: the
        # interface matrix is block tridiagonal in this order (a cut couples only to its neighbours)
        per_rank = [np.sort(g) for _, g in sorted(infos, key=lambda t: t[0])]
        self.block_sizes = [len(g) for g in per_rank if len(g)]
        self.gamma = np.concatenate(per_rank) if per_rank else np.zeros(0, dtype=np.int64)
        self.block_of = np.repeat(np.arange(len(self.block_sizes)), self.block_sizes)
        self.ng = len(self.gamma)
        order = np.argsort(self.gamma, kind="stable")
        gsorted = self.gamma[order]
        self._gsorted, self._gorder = gsorted, order

        def position(gids):
            p = order[np.searchsorted(gsorted, gids)]
            assert np.array_equal(self.gamma[p], gids)
            return p
        self.position = position
        for m in F.members:
            m.keep_pos = position(m.keep_gids)                       # position of B_r in Γ
            m.gamma_pos = position(m.d.local_gids[m.gamma_local])
            m.keep_t = m.ops.long_index(m.keep_local)
            m.gamma_t = m.ops.long_index(m.gamma_local)
            m.ops.set_roles(m.role)
        self.root = F.by_rank.get(0)
        self.interface_solver = None

    # gather variable-size device vectors of every rank on rank 0 (list indexed by rank) ------
    def _gather(self, parts, sizes):
        F = self.fleet
        if F.dist is None:
            return {m.rank: p for m, p in zip(F.members, parts)}
        import torch
        m = F.members[0]
        out = {}
        if m.rank == 0:
            out[0] = parts[0]
            reqs = []
            for r in range(1, F.nranks):
                buf = m.ops.vec(sizes[r])
                out[r] = buf
                reqs.append(F.dist.irecv(m.ops.as_tensor(buf), src=r))
            for q in reqs:
                q.wait()
        else:
            F.dist.send(m.ops.as_tensor(parts[0]), dst=0)
        return out

    def setup(self):
        """Returns False (on every rank) when a rank could not factorise, e.g. out of memory."""
        F = self.fleet
        oks, self.error = [], None
        for m in F.members:
            try:
                m.ops.factor()
                oks.append(np.array([1.0]))
            except Exception as e:          # keep the ranks in step: agree on the outcome first
                oks.append(np.array([0.0]))
                self.error = repr(e)
        if F.allreduce(oks)[0] < F.nranks - 0.5:
            return False
        self._assemble_interface()
        return True

    def refactor(self):
        for m in self.fleet.members:
            m.ops.refactor()
        self._assemble_interface()

    def _interface_pattern(self):
        """Once per pattern: where the entries A[Γ_r rows, Γ columns] sit in every rank's local CSR
        (slots) and in the interface matrix (rows, cols); sizes and positions of the kept sets."""
        F = self.fleet
        static = []
        for m in F.members:
            rp, ci = m.d.local.get_pattern()
            gl = m.gamma_local
            cnt = (rp[gl + 1] - rp[gl]).astype(np.int64)
            start = np.repeat(rp[gl].astype(np.int64), cnt)
            within = np.arange(cnt.sum(), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            slot = start + within
            cols_g = m.d.local_gids[ci[slot]]
            inG = np.isin(cols_g, self._gsorted)
            m.if_slots = m.ops.index_tensor(slot[inG])
            rows = np.repeat(m.gamma_pos, cnt)[inG]
            cols = self.position(cols_g[inG])
            static.append((m.rank, len(m.keep_local), m.keep_pos, rows, cols))
        if F.dist is not None:
            meta = [None] * F.nranks
            F.dist.all_gather_object(meta, static[0])
            static = meta
        self.nks = {r: nk for r, nk, _, _, _ in static}
        self.pos = {r: p for r, _, p, _, _ in static}
        self.ntrip = {r: len(rows) for r, _, _, rows, _ in static}
        self.tridiag = True
        for r, nk, p, rows, cols in static:
            self.tridiag &= bool(np.all(np.abs(self.block_of[rows] - self.block_of[cols]) <= 1))
            bl = self.block_of[p]
            self.tridiag &= bool(len(bl) == 0 or bl.max() - bl.min() <= 1)
        if self.root is not None:
            ops = self.root.ops
            self.trip_t = {r: (ops.long_index(rows), ops.long_index(cols)) for r, _, _, rows, cols in static}
            self.pos_t = {r: ops.long_index(self.pos[r]) for r in self.pos}
            self.xg = ops.vec(self.ng)
            self.gg = ops.vec(self.ng)
        else:
            self.xg = F.members[0].ops.vec(self.ng)

    def _assemble_interface(self):
        """S = A[Γ,Γ] + Σ_r Schur blocks on rank 0, then its factorisation.  Only device buffers move:
        the interface entries of every rank (gathered from its CSR values on the device) and its
        Schur block."""
        F = self.fleet
        if not hasattr(self, "nks"):
            self._interface_pattern()
        nks, pos = self.nks, self.pos
        vals = self._gather([m.ops.matrix_values(m.if_slots) for m in F.members], self.ntrip)
        blocks = [m.ops.schur_block(len(m.keep_local)) for m in F.members]
        got = self._gather(blocks, {r: nks[r] * nks[r] for r in nks})
        if self.root is not None:
            ops = self.root.ops
            S = ops.vec(self.ng * self.ng)
            for r, v in vals.items():
                ops.add_triplets(S, self.ng, self.trip_t[r][0], self.trip_t[r][1], v)
            for r, blk in got.items():
                ops.add_block(S, self.ng, pos[r], blk, nks[r])
            if self.tridiag and len(self.block_sizes) > 1:
                ops.chain_factor(self.block_sizes, S)      # O(sum n_r^3): chain of fronts
                self.interface_solver = "block-tridiagonal chain of %d fronts" % len(self.block_sizes)
            else:
                ops.dense_factor(self.ng, S)
                self.interface_solver = "dense"

    def apply(self, rs, zs):
        """zs = A^-1 rs on distributed vectors (one device vector per local member)."""
        F = self.fleet
        gs = []
        for m, r in zip(F.members, rs):
            g = m.ops.vec(len(m.keep_local))
            m.ops.schur_forward(r, g)                               # g = -A[B,I] A[I,I]^-1 r_I
            # b_Γ part: the owner adds its own right-hand-side entries
            own = m.ops.take(r, m.gamma_t)
            full = m.ops.vec(len(m.keep_local))
            k_of_gamma = np.searchsorted(m.keep_local, m.gamma_local)
            m.ops.put(full, m.ops.long_index(k_of_gamma), own)
            g += full
            gs.append(g)
        got = self._gather(gs, self.nks)
        if self.root is not None:
            ops = self.root.ops
            ops.fill_zero(self.gg)
            for r, g in got.items():
                ops.put(self.gg, self.pos_t[r], g, accumulate=True)
            ops.dense_solve(self.gg, self.xg)
        if F.dist is not None:
            F.dist.broadcast(F.members[0].ops.as_tensor(self.xg), src=0)
        for m, z in zip(F.members, zs):          # (in-process fleets share one device)
            xk = m.ops.take(self.xg, m.ops.long_index(m.keep_pos))
            m.ops.schur_backward(xk, z)
