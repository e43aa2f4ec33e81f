"""Lagrangian reference FE, `ReferenceFE(lagrangian,Float64,order)` (oracle).

Reference call sites: `src/Yago_freq_domain.jl:147-149`,
`src/Khabakhpasheva_freq_domain.jl:216`, `src/Periodic_Beam.jl:86-87`,
`src/Multi_geo_freq_domain.jl:111-113`.

Gridap semantics restated: equispaced nodes; node = DOF order is
vertices, then the interior nodes of each edge (from the edge's first to its
second vertex), then face interiors, then the cell interior (lattice order,
x fastest; simplices keep the lattice points with positive barycentrics).
Shape functions = monomial prebasis times the inverse nodal Vandermonde matrix.
"""
import itertools
import numpy as np
from . import polytopes as pt


def _face_polytope(cell, d):
    if d == 0:
        return pt.VERTEX
    if d == cell.dim:
        return cell
    return pt.simplex(d) if cell.is_simplex else pt.ncube(d)


def _interior_lattice(poly, order):
    """Interior equispaced nodes of `poly` in its own reference coordinates."""
    d = poly.dim
    if d == 0:
        return np.zeros((1, 0))
    rng = range(1, order)
    out = []
    # x fastest: iterate the last axis in the outer loop
    for idx in itertools.product(*([rng] * d)):
        idx = idx[::-1]
        if poly.is_simplex and d > 1 and sum(idx) >= order:
            continue
        out.append([i / order for i in idx])
    return np.array(out, dtype=float).reshape(-1, d)


def _linear_shape(poly, x):
    """Values of the vertex (Q1/P1) shape functions of `poly` at points x."""
    d = poly.dim
    if d == 0:
        return np.ones((len(x), 1))
    if poly.is_simplex and d > 1:
        return np.concatenate([1.0 - x.sum(axis=1, keepdims=True), x], axis=1)
    out = np.ones((len(x), len(poly.verts)))
    for v, corner in enumerate(poly.verts):
        for a in range(d):
            out[:, v] *= x[:, a] if corner[a] == 1 else (1.0 - x[:, a])
    return out


def _exponents(poly, order):
    d = poly.dim
    out = []
    for idx in itertools.product(*([range(order + 1)] * d)):
        idx = idx[::-1]
        if poly.is_simplex and d > 1 and sum(idx) > order:
            continue
        out.append(idx)
    return np.array(out, dtype=int).reshape(-1, d)


class LagrangianRefFE:
    def __init__(self, poly, order):
        self.poly = poly
        self.order = order
        self.dim = poly.dim
        nodes = []
        self.face_own_nodes = {}          # (d, iface) -> list of local node ids
        for d in range(0, poly.dim + 1):
            fpoly = _face_polytope(poly, d)
            own = _interior_lattice(fpoly, order) if d > 0 else np.zeros((1, 0))
            for iface, fv in enumerate(poly.face_vertices(d)):
                fx = poly.verts[list(fv)]
                x = _linear_shape(fpoly, own) @ fx if len(own) else np.zeros((0, poly.dim))
                ids = list(range(len(nodes), len(nodes) + len(x)))
                nodes.extend(x.tolist())
                self.face_own_nodes[(d, iface)] = ids
        self.nodes = np.array(nodes, dtype=float).reshape(-1, poly.dim)
        self.ndofs = len(self.nodes)
        self.exps = _exponents(poly, order)
        assert len(self.exps) == self.ndofs, (len(self.exps), self.ndofs)
        self.change = self._exact_change()

    def _exact_change(self):
        """Inverse nodal Vandermonde matrix.  Gridap takes a Float64 `inv`, which carries
        1e-13..1e-12 relative round-off at order 4 (SURVEY.md hard part 3); the nodes are
        rational (i/order), so the oracle inverts exactly in rational arithmetic and rounds
        once, and evaluates the monomial sums in extended precision - it is then the more
        accurate side of every 1e-12 comparison."""
        from fractions import Fraction
        p = self.order
        n = self.ndofs
        lat = np.rint(self.nodes * p).astype(int)
        V = [[Fraction(1) for _ in range(n)] for _ in range(n)]
        for i in range(n):
            for j in range(n):
                v = Fraction(1)
                for a in range(self.dim):
                    v *= Fraction(int(lat[i, a]), p) ** int(self.exps[j, a])
                V[i][j] = v
        # Gauss-Jordan on [V | I]
        M = [row + [Fraction(int(i == j)) for j in range(n)] for i, row in enumerate(V)]
        for c in range(n):
            piv = next(r for r in range(c, n) if M[r][c] != 0)
            M[c], M[piv] = M[piv], M[c]
            inv = 1 / M[c][c]
            M[c] = [x * inv for x in M[c]]
            for r in range(n):
                if r != c and M[r][c] != 0:
                    f = M[r][c]
                    M[r] = [x - f * y for x, y in zip(M[r], M[c])]
        ld = np.longdouble
        return np.array([[ld(x.numerator) / ld(x.denominator) for x in row[n:]] for row in M], dtype=ld)

    def _monomials(self, x, deriv=None):
        """Monomial values (npts, nmono); deriv = tuple of axes to differentiate."""
        x = np.asarray(x, dtype=np.longdouble).reshape(-1, self.dim)
        out = np.ones((len(x), len(self.exps)), dtype=np.longdouble)
        e = self.exps.copy()
        coef = np.ones(len(self.exps), dtype=np.longdouble)
        for a in (deriv or ()):
            coef = coef * e[:, a]
            e = e.copy()
            e[:, a] = np.maximum(e[:, a] - 1, 0)
        for a in range(self.dim):
            out *= x[:, a:a + 1] ** e[:, a][None, :]
        return out * coef[None, :]

    def values(self, x):
        """(npts, ndofs)."""
        return (self._monomials(x) @ self.change).astype(float)

    def gradients(self, x):
        """(npts, ndofs, dim) reference gradients."""
        return np.stack([(self._monomials(x, (a,)) @ self.change).astype(float)
                         for a in range(self.dim)], axis=2)

    def hessians(self, x):
        """(npts, ndofs, dim, dim) reference Hessians."""
        n = len(np.asarray(x).reshape(-1, self.dim))
        H = np.zeros((n, self.ndofs, self.dim, self.dim))
        for a in range(self.dim):
            for b in range(a, self.dim):
                h = (self._monomials(x, (a, b)) @ self.change).astype(float)
                H[:, :, a, b] = h
                H[:, :, b, a] = h
        return H


_cache = {}


def lagrangian(poly, order):
    key = (poly.name, order)
    if key not in _cache:
        _cache[key] = LagrangianRefFE(poly, order)
    return _cache[key]


def face_to_cell_points(cell_poly, d, iface, xface):
    """Map points given in the reference d-face into the reference cell."""
    fpoly = _face_polytope(cell_poly, d)
    fv = cell_poly.face_vertices(d)[iface]
    return _linear_shape(fpoly, xface) @ cell_poly.verts[list(fv)]
