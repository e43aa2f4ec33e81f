"""Quadrature rules as Gridap's `Measure(trian, degree)` builds them (oracle).

Reference call sites: `Measure(Ω,degree)` with `degree = 2*order`
(`src/Yago_freq_domain.jl:135-140`, `src/Khabakhpasheva_freq_domain.jl:198-209`,
`src/Periodic_Beam.jl:79-82`, `src/Multi_geo_freq_domain.jl:99-104`).

n-cubes: tensor Gauss-Legendre with ceil((degree+1)/2) points per axis on
[0,1], x fastest.  Simplices: Duffy-collapsed Gauss-Jacobi x Gauss-Legendre
(the rule of the Gridap 0.17.12 era; stated rule for MultiGeo parity, see
SURVEY.md A13).  A 0-dimensional face carries one point of weight 1.
"""
import numpy as np
from scipy.special import roots_jacobi


def gauss_legendre_01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _lex_indices(n, dim):
    """Index tuples of an n^dim lattice, first index fastest."""
    return [g.reshape(-1, order="F") for g in np.indices((n,) * dim)]


def npoints_for_degree(degree):
    return int(np.ceil((degree + 1.0) / 2.0))


def tensor_quadrature(dim, degree):
    """Points (nq, dim) with x fastest, weights (nq,)."""
    if dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    n = npoints_for_degree(degree)
    x1, w1 = gauss_legendre_01(n)
    idx = _lex_indices(n, dim)
    pts = np.stack([x1[i] for i in idx], axis=1)
    wts = np.prod(np.stack([w1[i] for i in idx], axis=1), axis=1)
    return pts, wts


def _gauss_jacobi_01(n, a, b):
    x, w = roots_jacobi(n, a, b)
    return 0.5 * (x + 1.0), 0.5 * w


def duffy_quadrature(dim, degree):
    """Collapsed-coordinate rule on the unit simplex, exact to `degree`."""
    n = npoints_for_degree(degree)
    xs, ws = [], []
    for d in range(1, dim):
        a = (dim - 1) - (d - 1)
        x, w = _gauss_jacobi_01(n, a, 0.0)
        # Jacobi weight (1-t)^a on [-1,1] -> (1-x)^a 2^a on [0,1]
        xs.append(x)
        ws.append(w * 0.5 ** a)
    x, w = gauss_legendre_01(n)
    xs.append(x)
    ws.append(w)
    idx = _lex_indices(n, dim)
    q = np.stack([xs[d][idx[d]] for d in range(dim)], axis=1)
    wts = np.prod(np.stack([ws[d][idx[d]] for d in range(dim)], axis=1), axis=1)
    pts = np.empty_like(q)
    pts[:, 0] = q[:, 0]
    for i in range(1, dim):
        pts[:, i] = q[:, i] * np.prod(1.0 - q[:, :i], axis=1)
    return pts, wts


def quadrature(polytope, degree):
    if polytope.dim == 0:
        return np.zeros((1, 0)), np.ones(1)
    if polytope.is_simplex and polytope.dim > 1:
        return duffy_quadrature(polytope.dim, degree)
    return tensor_quadrature(polytope.dim, degree)
