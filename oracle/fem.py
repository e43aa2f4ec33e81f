"""Triangulations, conforming DOF numbering and cell-wise operator tables (oracle).

Restates the Gridap 0.17 machinery behind the reference calls
`Boundary(model,tags=...)`, `Triangulation(Γ,ids)`, `Skeleton(Γb)`
(`src/Yago_freq_domain.jl:116-132`, `src/Khabakhpasheva_freq_domain.jl:114-181`),
`TestFESpace(trian,reffe)` / `MultiFieldFESpace` (`src/Yago_freq_domain.jl:147-157`)
and the evaluation of `∇`, `∇∇`, `Δ`, `jump`, `mean`, normals at the points of
`Measure(trian,degree)`.  Semantics: SURVEY.md A6-A13.  0-based ids.

Numbering rules restated here
  * Boundary(model,tags): facets whose own label entity is tagged, ascending
    facet id; each is glued to its first cell around (boundary facets have one).
  * FE space on a surface triangulation: DOFs live on the triangulation's
    active model = the facets' sub-faces keeping the PARENT's relative order
    (vertices, then edges, then facet interiors in triangulation cell order).
  * Conforming numbering: all vertex DOFs, then (p-1) per edge in edge order,
    then face-interior, then cell-interior; identity orientation (oriented
    meshes: Cartesian, and simplices with ascending vertex ids).
  * Skeleton(Γ): interior (d-1)-faces of Γ's active model in ascending order;
    plus = first surface cell around, minus = second.
"""
import numpy as np
from . import polytopes as pt
from .reffe import lagrangian, face_to_cell_points
from .quadrature import quadrature


# ---------------------------------------------------------------------------
# Triangulations
# ---------------------------------------------------------------------------
class VolumeTriangulation:
    def __init__(self, model):
        self.model = model
        self.poly = model.poly
        self.dim = model.D
        self.ncells = model.ncells

    def cell_coords(self):
        m = self.model
        return m.node_coords[m.cell_nodes]            # (ncells, nv, D)


class SurfaceTriangulation:
    """A set of facets of the background model, in a given order."""

    def __init__(self, model, facets):
        self.model = model
        D = model.D
        self.dim = D - 1
        self.poly = pt.facet_polytope(model.poly)
        self.facets = np.asarray(facets, dtype=np.int64)
        self.ncells = len(self.facets)
        around = getattr(model, "_around_facets", None)
        if around is None:
            around = model.cells_around(D - 1)
            model._around_facets = around
        self.cell = np.array([around[f][0][0] for f in self.facets], dtype=np.int64)
        self.lface = np.array([around[f][0][1] for f in self.facets], dtype=np.int64)

    def cell_coords(self):
        """Facet vertex coordinates taken from the glued cell (ncells, nvf, D)."""
        m = self.model
        lfv = np.array(m.poly.faces[self.dim], dtype=np.int64)      # (nlf, nvf)
        nodes = np.take_along_axis(m.cell_nodes[self.cell], lfv[self.lface], axis=1)
        return m.node_coords[nodes]

    def parent_cell_coords(self):
        m = self.model
        return m.node_coords[m.cell_nodes[self.cell]]

    def sub(self, ids):
        return SurfaceTriangulation(self.model, self.facets[np.asarray(ids, dtype=np.int64)])

    def position_of(self, other):
        """Index in `self` of each cell of `other` (same background facet)."""
        lut = {int(f): i for i, f in enumerate(self.facets)}
        return np.array([lut[int(f)] for f in other.facets], dtype=np.int64)

    # -- active model -------------------------------------------------------
    def active_model(self):
        if hasattr(self, "_active"):
            return self._active
        m = self.model
        d = self.dim
        act = {"cell_faces": {}, "nfaces": {}, "parent": {}}
        fv = m.face_vertices[d][self.facets]                   # (ncells, nvf)
        for k in range(0, d):
            lf = self.poly.face_vertices(k)
            if k == 0:
                parent_ids = fv
            else:
                lookup = getattr(m, "_lookup_%d" % k, None)
                if lookup is None:
                    lookup = {tuple(sorted(v)): i
                              for i, v in enumerate(m.face_vertices[k].tolist())}
                    setattr(m, "_lookup_%d" % k, lookup)
                parent_ids = np.array(
                    [[lookup[tuple(sorted(int(row[i]) for i in lv))] for lv in lf]
                     for row in fv], dtype=np.int64).reshape(self.ncells, len(lf))
            touched = np.unique(parent_ids)
            renum = {int(p): i for i, p in enumerate(touched)}
            act["parent"][k] = touched
            act["nfaces"][k] = len(touched)
            act["cell_faces"][k] = np.vectorize(renum.get, otypes=[np.int64])(parent_ids) \
                if parent_ids.size else parent_ids
        act["nfaces"][d] = self.ncells
        act["cell_faces"][d] = np.arange(self.ncells, dtype=np.int64)[:, None]
        self._active = act
        return act


def boundary(model, tags):
    mask = model.face_mask(tags, model.D - 1)
    return SurfaceTriangulation(model, np.nonzero(mask)[0])


class SkeletonTriangulation:
    """Interior (d-1)-faces of a surface triangulation's active model."""

    def __init__(self, surf, face_mask=None):
        self.surf = surf
        act = surf.active_model()
        k = surf.dim - 1
        c2f = act["cell_faces"][k]
        around = [[] for _ in range(act["nfaces"][k])]
        for c in range(surf.ncells):
            for lf, f in enumerate(c2f[c]):
                around[f].append((c, lf))
        faces = [f for f in range(len(around)) if len(around[f]) == 2]
        if face_mask is not None:
            faces = [f for f in faces if face_mask[f]]
        self.faces = np.array(faces, dtype=np.int64)
        self.nfaces = len(faces)
        self.plus = np.array([around[f][0][0] for f in faces], dtype=np.int64)
        self.plus_lf = np.array([around[f][0][1] for f in faces], dtype=np.int64)
        self.minus = np.array([around[f][1][0] for f in faces], dtype=np.int64)
        self.minus_lf = np.array([around[f][1][1] for f in faces], dtype=np.int64)


# ---------------------------------------------------------------------------
# FE spaces
# ---------------------------------------------------------------------------
class FESpace:
    """Conforming Lagrangian space on a volume or surface triangulation."""

    def __init__(self, trian, order):
        self.trian = trian
        self.order = order
        self.reffe = lagrangian(trian.poly, order)
        d = trian.dim
        if isinstance(trian, VolumeTriangulation):
            m = trian.model
            cell_faces = {k: m.cell_faces[k] for k in range(d + 1)}
            nfaces = {k: m.nfaces(k) for k in range(d + 1)}
        else:
            act = trian.active_model()
            cell_faces, nfaces = act["cell_faces"], act["nfaces"]
        nloc = self.reffe.ndofs
        cell_dofs = np.empty((trian.ncells, nloc), dtype=np.int64)
        offset = 0
        for k in range(d + 1):
            nlf = trian.poly.nfaces(k)
            own = [self.reffe.face_own_nodes[(k, lf)] for lf in range(nlf)]
            nown = len(own[0])
            if nown == 0:
                continue
            for lf in range(nlf):
                base = offset + cell_faces[k][:, lf] * nown
                for r, ln in enumerate(own[lf]):
                    cell_dofs[:, ln] = base + r
            offset += nfaces[k] * nown
        self.ndofs = offset
        self.cell_dofs = cell_dofs

    def dof_coords(self):
        """Physical coordinates of the DOF nodes (last writer wins)."""
        X = self.trian.cell_coords()
        geo = lagrangian(self.trian.poly, 1)
        N = geo.values(self.reffe.nodes)                      # (nloc, nv)
        xc = np.einsum("lv,evd->eld", N, X)
        out = np.zeros((self.ndofs, X.shape[2]))
        out[self.cell_dofs.reshape(-1)] = xc.reshape(-1, X.shape[2])
        return out

    def cell_dof_coords(self):
        X = self.trian.cell_coords()
        geo = lagrangian(self.trian.poly, 1)
        N = geo.values(self.reffe.nodes)
        return np.einsum("lv,evd->eld", N, X)


class MultiFieldFESpace:
    def __init__(self, spaces):
        self.spaces = spaces
        self.offsets = np.concatenate([[0], np.cumsum([s.ndofs for s in spaces])]).astype(np.int64)
        self.ndofs = int(self.offsets[-1])


# ---------------------------------------------------------------------------
# Operator tables at quadrature points
# ---------------------------------------------------------------------------
def _pinv_rows(Jt):
    """Jt (..., d, D) -> pinv (..., D, d) with grad_phys = pinv @ grad_ref, and
    the measure sqrt(det(Jt Jt^T))."""
    d, D = Jt.shape[-2], Jt.shape[-1]
    if d == D:
        inv = np.linalg.inv(Jt)                 # (D,d): x_b <- ref a
        return inv, np.abs(np.linalg.det(Jt))
    G = Jt @ np.swapaxes(Jt, -1, -2)            # (d,d)
    Ginv = np.linalg.inv(G)
    pinv = np.swapaxes(Jt, -1, -2) @ Ginv       # (D,d)
    return pinv, np.sqrt(np.linalg.det(G))


class CellOps:
    """Physical basis values / derivatives of one field on a set of cells.

    X        (ncells, nv, D)   vertex coordinates of the cells owning the basis
    poly     polytope of those cells, order = FE order
    xref     (nq, dim)  evaluation points in the cell's reference space, or
             (ncells, nq, dim) when they differ per cell (face-dependent)
    """

    def __init__(self, X, poly, order, xref, need_hess=False):
        reffe = lagrangian(poly, order)
        geo = lagrangian(poly, 1)
        xref = np.asarray(xref, dtype=float)
        per_cell = xref.ndim == 3
        pts = xref.reshape(-1, poly.dim)
        N = reffe.values(pts)
        dN = reffe.gradients(pts)
        dG = geo.gradients(pts)
        if per_cell:
            ne, nq = xref.shape[:2]
            N = N.reshape(ne, nq, reffe.ndofs)
            dN = dN.reshape(ne, nq, reffe.ndofs, poly.dim)
            dG = dG.reshape(ne, nq, geo.ndofs, poly.dim)
            Jt = np.einsum("eqva,evb->eqab", dG, X)
            self.val = N
        else:
            Jt = np.einsum("qva,evb->eqab", dG, X)
            self.val = np.broadcast_to(N[None], (len(X),) + N.shape)
        pinv, meas = _pinv_rows(Jt)                      # (e,q,D,d)
        self.Jt = Jt
        self.pinv = pinv
        self.meas = meas                                 # (e,q)
        if per_cell:
            self.grad = np.einsum("eqba,eqia->eqib", pinv, dN)
        else:
            self.grad = np.einsum("eqba,qia->eqib", pinv, dN)
        if need_hess:
            H = reffe.hessians(pts)
            if per_cell:
                H = H.reshape(ne, nq, reffe.ndofs, poly.dim, poly.dim)
                self.hess = np.einsum("eqba,eqiac,eqdc->eqibd", pinv, H, pinv)
            else:
                self.hess = np.einsum("eqba,qiac,eqdc->eqibd", pinv, H, pinv)
            self.lapl = np.einsum("eqibb->eqi", self.hess)

    def phys_points(self, X, poly, xref):
        geo = lagrangian(poly, 1)
        xref = np.asarray(xref, dtype=float)
        if xref.ndim == 3:
            ne, nq = xref.shape[:2]
            N = geo.values(xref.reshape(-1, poly.dim)).reshape(ne, nq, geo.ndofs)
            return np.einsum("eqv,evb->eqb", N, X)
        return np.einsum("qv,evb->eqb", geo.values(xref), X)


def phys_points(X, poly, xref):
    geo = lagrangian(poly, 1)
    xref = np.asarray(xref, dtype=float)
    if xref.ndim == 3:
        ne, nq = xref.shape[:2]
        N = geo.values(xref.reshape(-1, poly.dim)).reshape(ne, nq, geo.ndofs)
        return np.einsum("eqv,evb->eqb", N, X)
    return np.einsum("qv,evb->eqb", geo.values(xref), X)


class VolumeMeasure:
    """`Measure(Ω,degree)` with one field's operators."""

    def __init__(self, trian, degree):
        self.trian = trian
        self.xq, self.wq = quadrature(trian.poly, degree)

    def ops(self, order, cells=None):
        X = self.trian.cell_coords()
        if cells is not None:
            X = X[cells]
        op = CellOps(X, self.trian.poly, order, self.xq)
        op.dV = op.meas * self.wq[None, :]
        op.x = phys_points(X, self.trian.poly, self.xq)
        return op


class SurfaceMeasure:
    """`Measure(Γ,degree)`: surface-field operators and volume-field traces."""

    def __init__(self, trian, degree):
        self.trian = trian
        self.xq, self.wq = quadrature(trian.poly, degree)

    def surface_ops(self, order, need_hess=False):
        X = self.trian.cell_coords()
        op = CellOps(X, self.trian.poly, order, self.xq, need_hess=need_hess)
        op.dV = op.meas * self.wq[None, :]
        op.x = phys_points(X, self.trian.poly, self.xq)
        return op

    def trace_ops(self, order):
        """Operators of a volume field seen from the glued cell."""
        t = self.trian
        cp = t.model.poly
        nlf = cp.nfaces(t.dim)
        per_lf = np.stack([face_to_cell_points(cp, t.dim, lf, self.xq)
                           for lf in range(nlf)])            # (nlf, nq, D)
        xref = per_lf[t.lface]                                # (ncells, nq, D)
        return CellOps(t.parent_cell_coords(), cp, order, xref)


class SkeletonMeasure:
    """`Measure(Λ,degree)` on the skeleton of a surface triangulation."""

    def __init__(self, skel, degree):
        self.skel = skel
        surf = skel.surf
        self.fpoly = pt.facet_polytope(surf.poly)
        self.xq, self.wq = quadrature(self.fpoly, degree)

    def side_ops(self, order, side, need_hess=True):
        skel = self.skel
        surf = skel.surf
        sp = surf.poly
        k = surf.dim - 1
        cells = skel.plus if side == "plus" else skel.minus
        lfs = skel.plus_lf if side == "plus" else skel.minus_lf
        per_lf = np.stack([face_to_cell_points(sp, k, lf, self.xq)
                           for lf in range(sp.nfaces(k))])
        xref = per_lf[lfs]
        X = surf.cell_coords()[cells]
        op = CellOps(X, sp, order, xref, need_hess=need_hess)
        # outward co-normal of this side's cell: normalise(pinv . n_ref)
        nref = sp.facet_normals()[lfs]                           # (nf, d)
        n = np.einsum("eqba,ea->eqb", op.pinv, nref)
        op.normal = n / np.linalg.norm(n, axis=-1, keepdims=True)
        # measure of the skeleton face seen from this side
        if k == 0:
            op.dS = np.ones((len(cells), 1)) * self.wq[None, :]
        else:
            lv = np.array(sp.faces[k], dtype=np.int64)[lfs]     # (nf, 2)
            tref = sp.verts[lv[:, 1]] - sp.verts[lv[:, 0]]       # (nf, d)
            tphys = np.einsum("ea,eqab->eqb", tref, op.Jt)
            op.dS = np.linalg.norm(tphys, axis=-1) * self.wq[None, :]
        op.x = phys_points(X, sp, xref)
        return op
