"""NumPy/SciPy twin of the rank-local device arithmetic of `vlfs_b200.dist` (TEST
INFRASTRUCTURE ONLY): lets the `-m "not gpu"` tests run the partition / halo / distributed GMRES
host logic - in-process fleets and world_size-2 gloo - without a GPU.  `CpuLocalSystem` has the
surface `DistributedSystem` expects from its backend; `CpuOps` the surface `Fleet` expects from
`GpuOps`.  Nothing in `monolithicfemvlfs.jl_b200/` imports this module."""
import numpy as np
import scipy.sparse.linalg as spla
from .descriptor_eval import RecordingSystem


class CpuLocalSystem(RecordingSystem):
    def set_owned_rows(self, n):
        self.n_owned = int(n)

    def set_dof_coords(self, coords):
        self.coords = coords

    def build_pattern(self):
        return None

    def assemble(self):
        self.mats, self.vecs = self.evaluate()

    def get_vector(self, vec=0):
        return self.vecs[vec]

    def get_matrix(self, mat=0):
        return self.mats[mat].tocsr()

    def get_pattern(self):
        A = self.canonical(0)
        return A.indptr, A.indices

    def canonical(self, mat):
        """CSR with sorted column indices: the slot numbering `get_pattern` refers to."""
        A = self.mats[mat].tocsr()
        A.sort_indices()
        return A


class CpuOps:
    def __init__(self, dsys, mat=0):
        self.d, self.sys = dsys, dsys.local
        self.n_owned, self.n_local = dsys.n_owned, dsys.n_local
        self.A = self.sys.mats[mat][:self.n_owned]                 # owned rows, all local columns
        self.lu = spla.splu(self.A[:, :self.n_owned].tocsc())      # block-Jacobi: owned block
        self.device = "cpu"

    def zeros(self, rows=None):
        shape = (self.n_local,) if rows is None else (rows, self.n_local)
        return np.zeros(shape, dtype=self.d.dtype)

    def from_host(self, x):
        return np.array(x, dtype=self.d.dtype)

    def to_host(self, t):
        return np.asarray(t)

    def index_tensor(self, idx):
        return np.asarray(idx, dtype=np.int64)

    def spmv(self, x, y):
        y[:self.n_owned] = self.A @ x

    def precond(self, r, z):
        z[:self.n_owned] = self.lu.solve(np.ascontiguousarray(r[:self.n_owned]))

    def multidot(self, V, nv, w):
        V = V.reshape(-1, self.n_local)
        return np.conj(V[:nv, :self.n_owned]) @ w[:self.n_owned]

    def gemv_sub(self, V, nv, h, w):
        V = V.reshape(-1, self.n_local)
        w[:self.n_owned] -= np.asarray(h)[:nv] @ V[:nv, :self.n_owned]

    def axpby(self, a, x, b, y):
        n = self.n_owned
        y[:n] = a * x[:n] + (b * y[:n] if b != 0 else 0)

    def gather(self, idx, x):
        return x[idx].copy()

    def ghost_view(self, x, offset, count):
        return x[self.n_owned + offset:self.n_owned + offset + count]

    def set_ghost(self, x, offset, buf):
        x[self.n_owned + offset:self.n_owned + offset + len(buf)] = buf

    def as_tensor(self, a):
        import torch
        return torch.from_numpy(a.view(np.float64))      # shares memory with the NumPy buffer

    def fill_zero(self, x):
        x[...] = 0

    def rhs_vector(self, vec=0):
        return np.array(self.sys.vecs[vec], dtype=self.d.dtype)

    def sync(self):
        pass

    # -- substructuring twin -----------------------------------------------------------------
    def set_roles(self, role):
        self.role = np.asarray(role)
        self.E = np.nonzero(self.role == 0)[0]
        self.K = np.nonzero(self.role == 1)[0]

    def factor(self, **_):
        Afull = self.sys.mats[0].tocsr()
        self.AEE = spla.splu(Afull[self.E][:, self.E].tocsc())
        self.AKE = Afull[self.K][:, self.E]
        self.AEK = Afull[self.E][:, self.K]

    refactor = factor

    def schur_block(self, nk):
        X = self.AEE.solve(self.AEK.toarray().astype(self.d.dtype))
        return (-(self.AKE @ X)).reshape(-1)

    def schur_forward(self, b, g):
        self._bE = np.array(b[self.E])
        g += -(self.AKE @ self.AEE.solve(self._bE))

    def schur_backward(self, xk, x):
        x[self.E] = self.AEE.solve(self._bE - self.AEK @ xk)
        x[self.K] = xk

    def dense_factor(self, n, S):
        import scipy.linalg as sla
        self._dense = sla.lu_factor(np.asarray(S).reshape(n, n))

    def chain_factor(self, sizes, S):
        assert sum(sizes) * sum(sizes) == len(S)
        self.dense_factor(int(sum(sizes)), S)

    def matrix_values(self, slots):
        return np.array(self.sys.canonical(0).data[slots], dtype=self.d.dtype)

    def add_triplets(self, S, n, rows, cols, vals):
        np.add.at(S.reshape(n, n), (rows, cols), vals)

    def dense_solve(self, b, x):
        import scipy.linalg as sla
        x[...] = sla.lu_solve(self._dense, b)

    def vec(self, n):
        return np.zeros(n, dtype=self.d.dtype)

    def dense_from_host(self, a):
        return np.asarray(a, dtype=self.d.dtype).reshape(-1).copy()

    def add_block(self, S, n, idx, block, nb):
        S.reshape(n, n)[np.ix_(idx, idx)] += np.asarray(block).reshape(nb, nb)

    def take(self, x, idx):
        return np.array(x[idx])

    def put(self, x, idx, vals, accumulate=False):
        if accumulate:
            np.add.at(x, idx, vals)
        else:
            x[idx] = vals

    def long_index(self, idx):
        return np.asarray(idx, dtype=np.int64)

    def to_index_host(self, idx):
        return np.asarray(idx)
