"""Reference FEs, quadratures, FE spaces and measures: producers of the shape-function and
DOF tables the CUDA library consumes.

Mirrors (names kept):
  ReferenceFE(lagrangian,Float64,order), TestFESpace(trian,reffe;vector_type),
  TrialFESpace, MultiFieldFESpace, Measure(trian,degree)
      src/Yago_freq_domain.jl:135-157, src/Khabakhpasheva_freq_domain.jl:198-224,
      src/Periodic_Beam.jl:79-93, src/Multi_geo_freq_domain.jl:99-121
n-cube bases are evaluated as tensor products of 1-D Lagrange polynomials (product form),
simplex bases through a monomial Vandermonde solve; node = DOF order follows Gridap
(vertices, edge interiors, face interiors, cell interior).
"""
import itertools
import numpy as np
from scipy.special import roots_jacobi
from .geometry import Polytope, Interior, BoundaryTriangulation, Skeleton

lagrangian = "lagrangian"


# ---------------------------------------------------------------------------------------
def _interior_lattice(poly, order):
    d = poly.dim
    if d == 0:
        return np.zeros((1, 0))
    pts = []
    for idx in itertools.product(*([range(1, order)] * d)):
        idx = idx[::-1]                                   # first axis fastest
        if poly.is_simplex and d > 1 and sum(idx) >= order:
            continue
        pts.append([i / order for i in idx])
    return np.array(pts, dtype=np.float64).reshape(-1, d)


def _vertex_shape(poly, x):
    d = poly.dim
    if d == 0:
        return np.ones((len(x), 1))
    if poly.is_simplex and d > 1:
        return np.concatenate([1.0 - x.sum(axis=1, keepdims=True), x], axis=1)
    out = np.ones((len(x), len(poly.verts)))
    for v, c in enumerate(poly.verts):
        for a in range(d):
            out[:, v] *= x[:, a] if c[a] == 1 else 1.0 - x[:, a]
    return out


def _vertex_shape_grad(poly, x):
    """(npts, nv, d) gradients of the Q1/P1 geometry map basis."""
    d = poly.dim
    n = len(x)
    if poly.is_simplex and d > 1:
        g = np.zeros((n, d + 1, d))
        g[:, 0, :] = -1.0
        for a in range(d):
            g[:, a + 1, a] = 1.0
        return g
    nv = len(poly.verts)
    g = np.ones((n, nv, d))
    for v, c in enumerate(poly.verts):
        for a in range(d):
            for b in range(d):
                if a == b:
                    g[:, v, a] *= 1.0 if c[b] == 1 else -1.0
                else:
                    g[:, v, a] *= x[:, b] if c[b] == 1 else 1.0 - x[:, b]
    return g


def face_points_in_cell(cell_poly, d, lface, xface):
    """Points of the reference d-face `lface` expressed in the cell's reference space."""
    fpoly = cell_poly.face_polytope(d)
    return _vertex_shape(fpoly, xface) @ cell_poly.verts[cell_poly.lfaces(d)[lface]]


class LagrangianRefFE:
    """Scalar Lagrangian reference element of a given order on a polytope."""

    def __init__(self, poly, order):
        self.poly, self.order, self.dim = poly, order, poly.dim
        nodes, own = [], {}
        for d in range(poly.dim + 1):
            fpoly = poly.face_polytope(d)
            lat = _interior_lattice(fpoly, order) if d > 0 else np.zeros((1, 0))
            for lf, fv in enumerate(poly.lfaces(d)):
                x = _vertex_shape(fpoly, lat) @ poly.verts[fv] if len(lat) else np.zeros((0, poly.dim))
                own[(d, lf)] = list(range(len(nodes), len(nodes) + len(x)))
                nodes.extend(x.tolist())
        self.nodes = np.array(nodes, dtype=np.float64).reshape(-1, poly.dim)
        self.face_own_nodes = own
        self.ndofs = len(self.nodes)
        if poly.is_simplex and poly.dim > 1:
            ex = []
            for idx in itertools.product(*([range(order + 1)] * poly.dim)):
                idx = idx[::-1]
                if sum(idx) <= order:
                    ex.append(idx)
            self._exps = np.array(ex, dtype=np.int64)
            self._change = np.linalg.inv(self._mono(self.nodes, ()))
        else:
            self._lat = np.rint(self.nodes * order).astype(np.int64)      # (ndofs, dim)
            self._x1 = np.arange(order + 1) / order

    # ---- simplices: monomials ------------------------------------------------------
    def _mono(self, x, deriv):
        e = self._exps.copy()
        coef = np.ones(len(e))
        for a in deriv:
            coef = coef * e[:, a]
            e[:, a] = np.maximum(e[:, a] - 1, 0)
        out = np.ones((len(x), len(e)))
        for a in range(self.dim):
            out *= x[:, a:a + 1] ** e[:, a][None, :]
        return out * coef[None, :]

    # ---- n-cubes: 1-D Lagrange in product form ---------------------------------------
    def _lag1d(self, x):
        """values, first and second derivatives of the 1-D basis: each (npts, order+1)."""
        xs = self._x1
        p = len(xs)
        n = len(x)
        L = np.ones((n, p)); dL = np.zeros((n, p)); ddL = np.zeros((n, p))
        for k in range(p):
            others = [m for m in range(p) if m != k]
            den = np.prod([xs[k] - xs[m] for m in others])
            L[:, k] = np.prod([x - xs[m] for m in others], axis=0) / den
            s1 = np.zeros(n); s2 = np.zeros(n)
            for a in others:
                ra = [m for m in others if m != a]
                s1 += np.prod([x - xs[m] for m in ra], axis=0) if ra else np.ones(n)
                for b in ra:
                    rb = [m for m in ra if m != b]
                    s2 += np.prod([x - xs[m] for m in rb], axis=0) if rb else np.ones(n)
            dL[:, k] = s1 / den
            ddL[:, k] = s2 / den
        return L, dL, ddL

    def tabulate(self, x, need_hess=False):
        """val (npts,ndofs), grad (npts,ndofs,dim), hess (npts,ndofs,dim,dim) or None."""
        x = np.asarray(x, dtype=np.float64).reshape(-1, self.dim)
        n, nd, D = len(x), self.ndofs, self.dim
        if D == 0:
            return np.ones((n, 1)), np.zeros((n, 1, 0)), None
        if self.poly.is_simplex and D > 1:
            val = self._mono(x, ()) @ self._change
            grad = np.stack([self._mono(x, (a,)) @ self._change for a in range(D)], axis=2)
            hess = None
            if need_hess:
                hess = np.zeros((n, nd, D, D))
                for a in range(D):
                    for b in range(a, D):
                        h = self._mono(x, (a, b)) @ self._change
                        hess[:, :, a, b] = h
                        hess[:, :, b, a] = h
            return val, grad, hess
        tabs = [self._lag1d(x[:, a]) for a in range(D)]
        val = np.ones((n, nd)); grad = np.ones((n, nd, D))
        hess = np.ones((n, nd, D, D)) if need_hess else None
        for a in range(D):
            La, dLa, ddLa = (t[:, self._lat[:, a]] for t in tabs[a])
            val *= La
            for b in range(D):
                grad[:, :, b] *= dLa if a == b else La
            if need_hess:
                for b in range(D):
                    for c in range(D):
                        if a == b and a == c:
                            hess[:, :, b, c] *= ddLa
                        elif a == b or a == c:
                            hess[:, :, b, c] *= dLa
                        else:
                            hess[:, :, b, c] *= La
        return val, grad, hess


_reffes = {}


def get_reffe(poly, order):
    key = (poly.name, order)
    if key not in _reffes:
        _reffes[key] = LagrangianRefFE(poly, order)
    return _reffes[key]


# ---------------------------------------------------------------------------------------
def _gl01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def quadrature(poly, degree):
    """`Quadrature(polytope,degree)`: tensor Gauss-Legendre on n-cubes, Duffy-collapsed
    Gauss-Jacobi on simplices, one unit-weight point on a vertex."""
    d = poly.dim
    if d == 0:
        return np.zeros((1, 0)), np.ones(1)
    n = int(np.ceil((degree + 1.0) / 2.0))
    idx = [g.reshape(-1, order="F") for g in np.indices((n,) * d)]
    if not (poly.is_simplex and d > 1):
        x1, w1 = _gl01(n)
        return (np.stack([x1[i] for i in idx], axis=1),
                np.prod(np.stack([w1[i] for i in idx], axis=1), axis=1))
    xs, ws = [], []
    for a in range(1, d):
        al = (d - 1) - (a - 1)
        x, w = roots_jacobi(n, al, 0.0)
        xs.append(0.5 * (x + 1.0))
        ws.append(w * 0.5 ** (al + 1))
    x, w = _gl01(n)
    xs.append(x); ws.append(w)
    q = np.stack([xs[a][idx[a]] for a in range(d)], axis=1)
    wts = np.prod(np.stack([ws[a][idx[a]] for a in range(d)], axis=1), axis=1)
    pts = q.copy()
    for a in range(1, d):
        pts[:, a] = q[:, a] * np.prod(1.0 - q[:, :a], axis=1)
    return pts, wts


# ---------------------------------------------------------------------------------------
class FESpace:
    """`TestFESpace(trian, ReferenceFE(lagrangian,Float64,order))`: conforming numbering,
    vertices first, then edges, faces, cell interiors; identity orientation."""

    def __init__(self, trian, order, vector_type=np.float64):
        self.trian, self.order, self.vector_type = trian, order, vector_type
        self.reffe = get_reffe(trian.poly, order)
        d = trian.dim
        if isinstance(trian, Interior):
            m = trian.model
            cell_faces = m.cell_faces
            nfaces = {k: m.nfaces(k) for k in range(d + 1)}
        else:
            act = trian.active_model()
            cell_faces, nfaces = act["cell_faces"], act["nfaces"]
        cell_dofs = np.empty((trian.ncells, self.reffe.ndofs), dtype=np.int64)
        offset = 0
        for k in range(d + 1):
            nlf = trian.poly.nfaces(k)
            nown = len(self.reffe.face_own_nodes[(k, 0)])
            if nown == 0:
                continue
            for lf in range(nlf):
                base = offset + cell_faces[k][:, lf] * nown
                for r, ln in enumerate(self.reffe.face_own_nodes[(k, lf)]):
                    cell_dofs[:, ln] = base + r
            offset += nfaces[k] * nown
        self.ndofs = int(offset)
        self.cell_dofs = cell_dofs

    def cell_dof_coords(self):
        X = self.trian.cell_coords()
        N = _vertex_shape(self.trian.poly, self.reffe.nodes)
        return np.einsum("lv,evd->eld", N, X)


def ReferenceFE(name, T, order):
    """`ReferenceFE(lagrangian,Float64,order)`: a spec; the polytope comes from the triangulation."""
    assert name == lagrangian
    return (name, T, int(order))


def TestFESpace(trian, reffe, conformity="H1", vector_type=np.float64):
    """`TestFESpace(trian,reffe;conformity=:H1,vector_type=Vector{ComplexF64})`; no Dirichlet
    tags exist anywhere in the reference, so every DOF is free."""
    assert conformity in ("H1", ":H1")
    return FESpace(trian, reffe[2], vector_type=vector_type)


def TrialFESpace(V):
    return V


class MultiFieldFESpace:
    def __init__(self, spaces):
        self.spaces = list(spaces)
        self.offsets = np.concatenate([[0], np.cumsum([s.ndofs for s in self.spaces])]).astype(np.int64)
        self.ndofs = int(self.offsets[-1])
        for f, s in enumerate(self.spaces):
            s.field = f
            s.offset = int(self.offsets[f])

    def dof_coords(self):
        """(ndofs, D) physical coordinates of every DOF node of the monolithic system."""
        D = self.spaces[0].trian.cell_coords().shape[2]
        out = np.zeros((self.ndofs, D))
        for s in self.spaces:
            xc = s.cell_dof_coords()
            out[s.cell_dofs.reshape(-1) + s.offset] = xc.reshape(-1, D)
        return out

    def split(self, x):
        return [x[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self.spaces))]


# ---------------------------------------------------------------------------------------
class Measure:
    """`Measure(trian,degree)`; builds the per-slot tables of vlfs_slot_desc."""

    def __init__(self, trian, degree):
        self.trian, self.degree = trian, degree
        if isinstance(trian, Skeleton):
            self.qpoly = trian.poly.face_polytope(trian.dim - 1)
        else:
            self.qpoly = trian.poly
        self.xq, self.wq = quadrature(self.qpoly, degree)
        self.nq = len(self.wq)
        self.ncells = trian.ncells

    # geometry nodes of the cells a slot's basis lives on ------------------------------
    def _tables(self, poly, order, xref_variants, need_hess):
        reffe = get_reffe(poly, order)
        vals, grads, hesss, ggs = [], [], [], []
        for x in xref_variants:
            v, g, h = reffe.tabulate(x, need_hess)
            vals.append(v); grads.append(g); ggs.append(_vertex_shape_grad(poly, x))
            if need_hess:
                hesss.append(h)
        return (np.stack(vals), np.stack(grads), np.stack(hesss) if need_hess else None,
                np.stack(ggs), reffe.ndofs)

    def own_slot(self, space, need_hess=False):
        """A volume field integrated on its own cells (`dΩ`)."""
        t = self.trian
        val, grad, hess, gg, nd = self._tables(t.poly, space.order, [self.xq], need_hess)
        return dict(dref=t.dim, nv=len(t.poly.verts), nvar=1, ndofs=nd, coords=t.cell_coords(),
                    variant=None, geo_grad=gg, val=val, grad=grad, hess=hess,
                    dofs=space.cell_dofs + space.offset, ref_normal=None, ref_tangent=None)

    def surface_slot(self, space, need_hess=False):
        """A surface field defined on a (possibly larger) triangulation, integrated here."""
        t = self.trian
        val, grad, hess, gg, nd = self._tables(t.poly, space.order, [self.xq], need_hess)
        pos = np.arange(t.ncells) if space.trian is t else t.position_in(space.trian)
        return dict(dref=t.dim, nv=len(t.poly.verts), nvar=1, ndofs=nd, coords=t.cell_coords(),
                    variant=None, geo_grad=gg, val=val, grad=grad, hess=hess,
                    dofs=space.cell_dofs[pos] + space.offset, ref_normal=None, ref_tangent=None)

    def trace_slot(self, space):
        """A volume field seen from a boundary triangulation through the facet->cell glue."""
        t = self.trian
        cp = t.model.poly
        xv = [face_points_in_cell(cp, t.dim, lf, self.xq) for lf in range(cp.nfaces(t.dim))]
        val, grad, _, gg, nd = self._tables(cp, space.order, xv, False)
        return dict(dref=cp.dim, nv=len(cp.verts), nvar=len(xv), ndofs=nd,
                    coords=t.parent_cell_coords(), variant=t.lface, geo_grad=gg, val=val, grad=grad,
                    hess=None, dofs=space.cell_dofs[t.cell] + space.offset, ref_normal=None,
                    ref_tangent=None)

    def skeleton_slot(self, space, side, need_hess=True):
        """One side of a skeleton: the surface field of the plus / minus cell."""
        sk = self.trian
        surf, sp = sk.surf, sk.poly
        k = sk.dim - 1
        cells = sk.plus if side == "plus" else sk.minus
        lfs = sk.plus_lf if side == "plus" else sk.minus_lf
        xv = [face_points_in_cell(sp, k, lf, self.xq) for lf in range(sp.nfaces(k))]
        val, grad, hess, gg, nd = self._tables(sp, space.order, xv, need_hess)
        if space.trian is surf:
            pos = cells
        else:       # look the field up through the common background facet (only these cells)
            o = np.argsort(space.trian.facets, kind="stable")
            pos = o[np.searchsorted(space.trian.facets[o], surf.facets[cells])]
            assert np.all(space.trian.facets[pos] == surf.facets[cells])
        return dict(dref=sp.dim, nv=len(sp.verts), nvar=len(xv), ndofs=nd,
                    coords=surf.cell_coords()[cells], variant=lfs, geo_grad=gg, val=val, grad=grad,
                    hess=hess, dofs=space.cell_dofs[pos] + space.offset,
                    ref_normal=sp.facet_normals(),
                    ref_tangent=sp.facet_tangents() if sp.dim == 2 else None)

    def points(self):
        """Physical quadrature points (ncells, nq, D) (`get_cell_points(dΩ)`)."""
        t = self.trian
        if isinstance(t, Skeleton):
            sp = t.poly
            k = t.dim - 1
            xv = np.stack([face_points_in_cell(sp, k, lf, self.xq) for lf in range(sp.nfaces(k))])
            N = np.stack([_vertex_shape(sp, x) for x in xv])            # (nlf, nq, nv)
            X = t.surf.cell_coords()[t.plus]
            return np.einsum("eqv,evd->eqd", N[t.plus_lf], X)
        N = _vertex_shape(t.poly, self.xq)
        return np.einsum("qv,evd->eqd", N, t.cell_coords())
