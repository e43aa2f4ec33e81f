"""Row-partitioned (multi-GPU) assembly and solve: SURVEY.md §8(e).

The reference is a serial Julia program; this module is the partitioned twin of its
`AffineFEOperator` + `solve` path (src/Yago_freq_domain.jl:167-171, src/Multi_geo_freq_domain.jl:131-135):

* DOFs are split into slabs along one coordinate axis (owner-computes rows).  A rank registers
  with its own libvlfs_b200 context only the cells that touch one of its DOFs, with local DOF
  ids [owned | ghost]; every owned row of the local CSR is then complete - no matrix entry ever
  crosses GPUs and nothing is reduced between ranks during assembly.
* The PRODUCT path is `MultiSystem` (bottom of this file): partition and local numbering here on the
  host, everything else - halo exchange overlapped with the interior rows, all-reduce, distributed
  GMRES, substructuring on the chain of slab cuts - inside the library (csrc/dist.cu, behind
  `vlfs_multi_*` / `vlfs_dist_*`).  `bench.py --gpus N` and tests/test_gpu_multi.py use it.

`Fleet` / `GpuOps` / `SchurSolver` are the round-1 Python prototype of the same algorithms (lockstep
ranks inside one process, torch.distributed between processes).  They are kept as the executable
specification of the host logic: the gloo world_size-2 CPU tests (tests/test_dist_cpu.py, with the
NumPy arithmetic twin oracle/dist_cpu.py, test infrastructure) and tests/test_gpu_dist.py run them;
nothing in the product path or the bench does.
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


# ---------------------------------------------------------------------------------------------
# cell-based partition: recursive coordinate bisection of the cell centroids
# ---------------------------------------------------------------------------------------------
def rcb_tree(points, nparts, axes=None):
    """Recursive coordinate bisection of `points` (n, dim) into `nparts` parts of (nearly) equal
    size; a split is made along the axis (among `axes`, default all) of the largest extent, the
    part numbers grow along the split axis.  Returns the tree ("leaf", part) |
    ("split", axis, cut, left, right); classify other points with `rcb_classify`."""
    points = np.asarray(points, dtype=np.float64)
    axes = list(range(points.shape[1])) if axes is None else list(axes)

    def rec(idx, first, count):
        if count == 1 or len(idx) == 0:
            return ("leaf", first)
        P = points[idx]
        ext = [P[:, a].max() - P[:, a].min() for a in axes]
        a = axes[int(np.argmax(ext))]
        nl = count // 2
        order = np.argsort(P[:, a], kind="stable")
        k = (len(idx) * nl) // count
        k = min(max(k, 1), len(idx) - 1)
        lo, hi = P[order[k - 1], a], P[order[k], a]
        cut = 0.5 * (lo + hi)
        left = idx[P[:, a] <= cut] if lo < hi else idx[order[:k]]
        right = idx[P[:, a] > cut] if lo < hi else idx[order[k:]]
        return ("split", a, cut, rec(left, first, nl), rec(right, first + nl, count - nl))
    return rec(np.arange(len(points)), 0, nparts)


def rcb_classify(tree, points):
    points = np.asarray(points, dtype=np.float64)
    out = np.zeros(len(points), dtype=np.int32)

    def rec(node, idx):
        if node[0] == "leaf":
            out[idx] = node[1]
            return
        _, a, cut, left, right = node
        m = points[idx, a] <= cut
        rec(left, idx[m])
        rec(right, idx[~m])
    rec(tree, np.arange(len(points)))
    return out


def _cell_centroids(slots):
    """Centroid of every integration cell, taken from the slot with the highest reference
    dimension (a boundary facet follows the volume cell it is glued to)."""
    s = max(slots, key=lambda t: t["dref"])
    return np.asarray(s["coords"], dtype=np.float64).mean(axis=1)


def cell_partition_owner(domain_slots, ndofs, nranks, axes=None):
    """Cell-based partition: the cells of the domain with the most cells are bisected recursively
    by centroid (`rcb_tree`), the cells of every other domain fall where their centroid lies, and a
    DOF belongs to the HIGHEST part among the cells that touch it.  The unknowns a lower rank then
    holds as ghosts are exactly those on the faces between parts - the smallest interface a
    cell partition allows.  Returns (owner[ndofs], [part of every cell, per domain])."""
    if nranks == 1:
        return np.zeros(ndofs, dtype=np.int32), [np.zeros(len(sl[0]["dofs"]), dtype=np.int32) for sl in domain_slots]
    cents = [_cell_centroids(sl) for sl in domain_slots]
    primary = int(np.argmax([len(c) for c in cents]))
    tree = rcb_tree(cents[primary], nranks, axes)
    parts = [rcb_classify(tree, c) for c in cents]
    owner = np.full(ndofs, -1, dtype=np.int32)
    for sl, part in zip(domain_slots, parts):
        for s in sl:
            dofs = np.asarray(s["dofs"])
            np.maximum.at(owner, dofs.reshape(-1), np.repeat(part, dofs.shape[1]))
    owner[owner < 0] = 0                      # DOFs no cell touches (none in the reference's cases)
    return owner, parts


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

    def used_slots(self):
        """Slots some bilinear or linear term acts on (a measure-only slot, e.g. the Q1 face of the
        inlet right-hand side, carries placeholder DOF ids that must not create couplings)."""
        used = set()

        def note(spec):
            if isinstance(spec, dict):
                used.update(int(k) for k, v in spec.items() if v != 0)
            else:
                used.add(int(spec))
        for a, k in self.terms:          # (mat, coef, test_op, test_slots, trial_op, trial_slots, ...)
            note(a[3]); note(a[5])
        for a, k in self.rhs:            # (vec, coef, op, slots, ...)
            note(a[3])
        return sorted(used)

    def touch_dofs(self):
        """(ncells, n) global DOF ids of the used slots of every cell."""
        sl = self.kwargs["slots"]
        return np.concatenate([np.asarray(sl[k]["dofs"]) for k in self.used_slots()], axis=1)

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
                 backend=None, partition="dofs", rcb_axes=None):
        """partition: "dofs" - slabs of equal DOF count along coordinate `axis` (slab_owner);
        "cells" - recursive coordinate bisection of the cells, DOFs to the highest touching part
        (cell_partition_owner; `rcb_axes` restricts the split axes, e.g. (0,) for slabs)."""
        if backend is None:
            from .assembler import B200System as backend
        self.backend, self.device = backend, device
        self.ndofs_global, self.is_complex, self.nmat, self.nvec = int(ndofs), bool(is_complex), nmat, nvec
        self.rank, self.nranks, self.axis = rank, nranks, axis
        self.partition, self.rcb_axes = partition, rcb_axes
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
        if self.partition == "cells":
            self.owner, self.cell_parts = cell_partition_owner(
                [[d.kwargs["slots"][k] for k in d.used_slots()] for d in self.domains], N, self.nranks, self.rcb_axes)
            owner = self.owner
        else:
            self.owner = owner = slab_owner(self.coords[:, self.axis], self.nranks)
        mine = owner == self.rank
        self.owned = np.nonzero(mine)[0]
        touched = np.zeros(N, dtype=bool)
        for d in self.domains:
            td = d.touch_dofs()
            d.cells = np.nonzero(mine[td].any(axis=1))[0]
            touched[td[d.cells].reshape(-1)] = True
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
            used = d.used_slots()
            for ks, s in enumerate(kw.pop("slots")):
                t = dict(s)
                t["coords"] = np.asarray(s["coords"])[d.cells]
                t["dofs"] = g2l[np.asarray(s["dofs"])[d.cells]] if ks in used else np.zeros_like(np.asarray(s["dofs"])[d.cells])
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


def local_halo_arrays(dsys):
    """The halo plan of `vlfs_dist_set_halo` for the rank of `dsys`, computed without communication
    (every rank knows the global tables): peers, send_ptr/send_idx (owned local ids each peer
    needs, in the order of ITS ghost block = ascending global id) and recv_ptr (ranges of this
    rank's ghost block, which is grouped by owner)."""
    owner, r = dsys.owner, dsys.rank
    own = owner[dsys.ghosts]
    recv = {int(p): int((own == p).sum()) for p in np.unique(own)}
    # what peer q needs from me: my owned DOFs inside cells that touch a DOF q owns
    need = {}
    for d in dsys.domains:
        alld = d.touch_dofs()                                 # (ncells, DOFs of the used slots)
        own_c = owner[alld]
        has_mine = (own_c == r).any(axis=1)
        for q in np.unique(own_c[has_mine]):
            if q == r:
                continue
            cells = has_mine & (own_c == q).any(axis=1)
            g = alld[cells][own_c[cells] == r]
            need.setdefault(int(q), []).append(np.unique(g))
    send = {q: np.unique(np.concatenate(v)) for q, v in need.items()}
    peers = sorted(set(recv) | set(send))
    send_ptr, recv_ptr, send_idx = [0], [0], []
    # the ghost block is grouped by owner in ascending owner order: recv ranges follow that order
    for p in peers:
        ids = send.get(p, np.zeros(0, dtype=np.int64))
        send_idx.append(dsys.g2l[ids])
        send_ptr.append(send_ptr[-1] + len(ids))
        recv_ptr.append(recv_ptr[-1] + recv.get(p, 0))
    send_idx = np.concatenate(send_idx) if send_idx else np.zeros(0, dtype=np.int64)
    return (np.asarray(peers, dtype=np.int32), np.asarray(send_ptr, dtype=np.int64),
            send_idx.astype(np.int32), np.asarray(recv_ptr, dtype=np.int64))


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
        # the torch ops of this class and the library's kernels must share ONE stream, or halo / Schur data
        # race silently: bound here, not left to the caller (vlfs_set_stream refuses once a solver exists,
        # so constructing GpuOps after solver_setup fails loudly instead)
        self.bind_stream()

    def bind_stream(self):
        """Run the library on torch's current stream (done by the constructor; before solver_setup)."""
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


# ---------------------------------------------------------------------------------------------
# the multi-GPU system behind the C ABI: communicator, halo exchange, Krylov reductions and the
# substructuring solve all live in libvlfs_b200 (csrc/dist.cu); this class only partitions the
# tables and marshals them
# ---------------------------------------------------------------------------------------------
class _FanoutDomain:
    def __init__(self, parts):
        self.parts = parts

    def _fan(self, name, *a, **k):
        out = [getattr(p, name)(*a, **k) for p in self.parts]
        return out[0]

    def add_term(self, *a, **k): return self._fan("add_term", *a, **k)
    def add_rhs(self, *a, **k): return self._fan("add_rhs", *a, **k)
    def set_coef(self, *a, **k): return self._fan("set_coef", *a, **k)
    def set_rhs_coef(self, *a, **k): return self._fan("set_rhs_coef", *a, **k)
    def set_weights(self, *a, **k): return self._fan("set_weights", *a, **k)

    @property
    def cells(self):
        return self.parts[0].cells


class MultiSystem:
    """`B200System` surface over `nranks` GPUs.  `local_ranks` are the ranks this process hosts:
    all of them (one process, one host thread per rank inside the library - what the Julia shim
    does through `ccall`: `vlfs_multi_create`) or one (one process per GPU: pass the 128-byte
    `unique_id` rank 0 got from `MultiSystem.unique_id()`).  Case modules use it unchanged:
    `YagoProblem(params, system_cls=lambda *a, **k: MultiSystem(*a, nranks=2, devices=[0, 1], **k))`."""

    @staticmethod
    def unique_id():
        from . import _lib as L
        buf = (C.c_char * 128)()
        rc = L.load().vlfs_dist_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise L.VlfsError("vlfs_dist_unique_id failed (status %d): NCCL not available" % rc)
        return bytes(buf)

    def __init__(self, ndofs, is_complex, nmat=1, nvec=1, device=None, nranks=1, devices=None, local_ranks=None,
                 unique_id=None, partition="cells", rcb_axes=(0,), axis=0):
        from . import _lib as L
        from .assembler import B200System
        self.lib = L.load()
        self.ndofs, self.is_complex, self.nmat, self.nvec = int(ndofs), bool(is_complex), nmat, nvec
        self.dtype = np.complex128 if is_complex else np.float64
        self.nranks = nranks
        self.local_ranks = list(range(nranks)) if local_ranks is None else list(local_ranks)
        nl = len(self.local_ranks)
        devices = list(devices) if devices is not None else [device or 0] * nl
        assert len(devices) == nl
        self._ctx_arr = (C.c_void_p * nl)()
        if nl == nranks:                                    # every rank in this process
            dv = (C.c_int * nl)(*devices)
            rc = self.lib.vlfs_multi_create(self._ctx_arr, nl, dv)
            if rc != 0:
                raise L.VlfsError("vlfs_multi_create failed (status %d): needs %d usable sm_100 GPUs and NCCL; "
                                  "there is no CPU fallback" % (rc, nl))
            self._created = "multi"
        else:                                               # one rank per process
            assert nl == 1 and unique_id is not None and len(unique_id) == 128
            ctx = C.c_void_p()
            rc = self.lib.vlfs_create(C.byref(ctx), devices[0])
            if rc != 0:
                raise L.VlfsError("vlfs_create failed (status %d): no usable sm_100 GPU" % rc)
            self._ctx_arr[0] = ctx
            self._created = "rank"
            self._uid = unique_id
        self.parts = []
        for k, r in enumerate(self.local_ranks):
            ctx = C.c_void_p(self._ctx_arr[k])
            mk = (lambda c: (lambda n, cplx, nmat=1, nvec=1, device=0: B200System.adopt(c, n, cplx, nmat, nvec)))(ctx)
            self.parts.append(DistributedSystem(ndofs, is_complex, nmat=nmat, nvec=nvec, device=devices[k], rank=r,
                                                nranks=nranks, axis=axis, backend=mk, partition=partition,
                                                rcb_axes=rcb_axes))
        self.nnz = self.ncoo = None

    # -- registration fans out to the ranks (each keeps its own cells) ------------------------------
    def add_domain(self, *a, **k):
        return _FanoutDomain([p.add_domain(*a, **k) for p in self.parts])

    def set_dof_coords(self, coords):
        for p in self.parts:
            p.set_dof_coords(coords)

    def _check(self, rc, allow=(0,)):
        if rc not in allow:
            msgs = [self.lib.vlfs_last_error(C.c_void_p(self._ctx_arr[k])) for k in range(len(self.parts))]
            from . import _lib as L
            raise L.VlfsError("libvlfs_b200 status %d: %s" % (rc, " | ".join(m.decode() for m in msgs if m)))
        return rc

    def build_pattern(self):
        from . import _lib as L
        for k, p in enumerate(self.parts):                 # host partition + table extraction per rank
            if self._created == "rank":
                # the communicator needs the context's scalar type: after the local system exists
                pass
            p.build_pattern()
        if self._created == "rank":
            ctx = C.c_void_p(self._ctx_arr[0])
            uid = (C.c_char * 128).from_buffer_copy(self._uid)
            self._check(self.lib.vlfs_dist_init(ctx, self.local_ranks[0], self.nranks, C.cast(uid, C.c_void_p)))
        for k, p in enumerate(self.parts):
            peers, sp, si, rp = local_halo_arrays(p)
            self._check(self.lib.vlfs_dist_set_halo(C.c_void_p(self._ctx_arr[k]), p.n_owned, len(peers), L.iptr(peers),
                                                    sp.ctypes.data_as(C.POINTER(C.c_int64)), L.iptr(si),
                                                    rp.ctypes.data_as(C.POINTER(C.c_int64))))
            p.halo = (peers, sp, si, rp)
        rps = [p.local.get_pattern()[0] for p in self.parts]
        self.nnz_owned = [int(rp[p.n_owned]) for rp, p in zip(rps, self.parts)]
        self.nnz = sum(self.nnz_owned)                     # of the ranks hosted here
        self.ncoo = sum(p.ncoo or 0 for p in self.parts)
        return self.nnz

    def assemble(self, matrices=True, vectors=True):
        if matrices and vectors:
            self._check(self.lib.vlfs_multi_assemble(self._ctx_arr, len(self.parts)))
        else:
            for p in self.parts:
                p.local.assemble(matrices, vectors)

    def solver_setup(self, mat=0, method=None, precond=None, restart=20, maxit=40, rtol=1e-10, block_size_hint=0, sweeps=0):
        from . import _lib as L
        o = L.SolverOpts(L.SOLVER_GMRES if method is None else method, L.PRECOND_SPARSE_LU if precond is None else precond,
                         restart, maxit, rtol, block_size_hint, sweeps)
        self.factor_status = self._check(self.lib.vlfs_multi_solver_setup(self._ctx_arr, len(self.parts), mat, C.byref(o)),
                                         allow=(0, L.PIVOT_PERTURBED))

    def refactor(self, mat=0):
        from . import _lib as L
        self.factor_status = self._check(self.lib.vlfs_multi_solver_refactor(self._ctx_arr, len(self.parts), mat),
                                         allow=(0, L.PIVOT_PERTURBED))

    def solve(self, vec=0):
        """Solves with the assembled right-hand side `vec`; returns the GLOBAL vector (entries owned by
        the ranks hosted here; the others are zero) and the statistics of the first local rank."""
        from . import _lib as L
        n = len(self.parts)
        xs = [np.zeros(p.n_local, dtype=self.dtype) for p in self.parts]
        ptrs = (C.c_void_p * n)(*[x.ctypes.data_as(C.c_void_p).value for x in xs])
        st = (L.SolveStats * n)()
        self._check(self.lib.vlfs_multi_solve_vec(self._ctx_arr, n, vec, ptrs, st), allow=(0, L.NOT_CONVERGED))
        xg = np.zeros(self.ndofs, dtype=self.dtype)
        for p, x in zip(self.parts, xs):
            p.scatter_owned(x, xg)
        s0 = st[0]
        return xg, dict(iterations=s0.iterations, converged=bool(s0.converged), rel_residual=s0.rel_residual,
                        setup_ms=s0.setup_ms, solve_ms=max(s.solve_ms for s in st), backward_error=s0.backward_error,
                        at_attainable_accuracy=s0.converged == 2)

    def spmv(self, x, mat=0):
        """y = A x for a GLOBAL host vector (entries of the rows owned here)."""
        n = len(self.parts)
        xs = [p.to_local(x) for p in self.parts]
        ys = [np.zeros(p.n_local, dtype=self.dtype) for p in self.parts]
        xp = (C.c_void_p * n)(*[v.ctypes.data_as(C.c_void_p).value for v in xs])
        yp = (C.c_void_p * n)(*[v.ctypes.data_as(C.c_void_p).value for v in ys])
        self._check(self.lib.vlfs_multi_spmv(self._ctx_arr, n, mat, xp, yp))
        yg = np.zeros(self.ndofs, dtype=self.dtype)
        for p, y in zip(self.parts, ys):
            p.scatter_owned(y, yg)
        return yg

    def spmv_device_ms(self, mat=0, reps=10):
        n = len(self.parts)
        ms = (C.c_float * n)()
        self._check(self.lib.vlfs_multi_spmv_device(self._ctx_arr, n, mat, reps, ms))
        return max(ms) / reps

    def get_vector(self, vec=0):
        bg = np.zeros(self.ndofs, dtype=self.dtype)
        for p in self.parts:
            p.scatter_owned(p.local.get_vector(vec), bg)
        return bg

    def dist_info(self):
        out = []
        for k in range(len(self.parts)):
            a = np.zeros(4)
            self.lib.vlfs_dist_info(C.c_void_p(self._ctx_arr[k]), a.ctypes.data_as(C.POINTER(C.c_double)))
            out.append(dict(n_left=int(a[0]), n_right=int(a[1]), chain_ms=float(a[2]), boundary_rows=int(a[3])))
        return out

    def launches(self):
        return sum(p.local.launches() for p in self.parts)

    def close(self):
        for k in range(len(self.parts)):
            if self._ctx_arr[k]:
                self.lib.vlfs_destroy(C.c_void_p(self._ctx_arr[k]))
                self._ctx_arr[k] = None
        self.parts = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
