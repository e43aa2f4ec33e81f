"""COO collection and `sparse(I,J,V)` semantics (oracle).

Restates what `AffineFEOperator(a,l,X,Y)` does inside Gridap 0.17 after the
cell-wise integration (reference call sites `src/Yago_freq_domain.jl:167`,
`src/Khabakhpasheva_freq_domain.jl:240`, `src/Multi_geo_freq_domain.jl:131`):
for every touched (test field, trial field) block of every cell, all entries
are pushed as (i,j,v) triplets - including basis functions that vanish on the
integration face - and `sparse(I,J,V,N,N)` sums duplicates while KEEPING
explicit zeros.
"""
import numpy as np
import scipy.sparse as sp


class COO:
    def __init__(self, n, dtype):
        self.n = n
        self.dtype = dtype
        self.I, self.J, self.V = [], [], []

    def add_block(self, rows, cols, vals):
        """rows (ne,nr), cols (ne,nc), vals (ne,nr,nc)."""
        ne, nr = rows.shape
        nc = cols.shape[1]
        assert vals.shape == (ne, nr, nc), (vals.shape, ne, nr, nc)
        self.I.append(np.broadcast_to(rows[:, :, None], (ne, nr, nc)).reshape(-1))
        self.J.append(np.broadcast_to(cols[:, None, :], (ne, nr, nc)).reshape(-1))
        self.V.append(np.asarray(vals, dtype=self.dtype).reshape(-1))

    @property
    def ncoo(self):
        return int(sum(len(v) for v in self.V))

    def tocsc(self):
        I = np.concatenate(self.I) if self.I else np.zeros(0, dtype=np.int64)
        J = np.concatenate(self.J) if self.J else np.zeros(0, dtype=np.int64)
        V = np.concatenate(self.V) if self.V else np.zeros(0, dtype=self.dtype)
        A = sp.coo_matrix((V, (I, J)), shape=(self.n, self.n)).tocsc()
        A.sum_duplicates()          # sums, never drops explicit zeros
        A.sort_indices()
        return A

    def tocsr(self):
        A = self.tocsc().tocsr()
        A.sort_indices()
        return A


class VEC:
    def __init__(self, n, dtype):
        self.v = np.zeros(n, dtype=dtype)

    def add(self, rows, vals):
        np.add.at(self.v, rows.reshape(-1), np.asarray(vals).reshape(-1))
