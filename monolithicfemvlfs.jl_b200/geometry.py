"""Discrete models and triangulations: the table-producing mirror of the Gridap calls the
reference makes before `AffineFEOperator` (vectorised NumPy; scales to 10^6-10^7 cells).

Mirrors (names kept):
  CartesianDiscreteModel(domain,partition;map,isperiodic)   src/Yago_freq_domain.jl:98-108,
                                                            src/Khabakhpasheva_freq_domain.jl:94-104,
                                                            src/Periodic_Beam.jl:60-62
  DiscreteModelFromFile(file)                                src/Multi_geo_freq_domain.jl:66
  add_tag_from_tags!, Interior, Boundary, Triangulation(Γ,ids), Skeleton
                                                            src/Yago_freq_domain.jl:110-132
In production these tables come from Gridap itself through the Julia shim;
here they are generated with the same numbering rules so that the
DOF numbering and sparsity pattern are a property of the INPUTS of the CUDA library.
All ids are 0-based.
"""
import json
import numpy as np

# ---------------------------------------------------------------------------------------
# reference polytopes (vertex coordinates + local faces per dimension)
# ---------------------------------------------------------------------------------------
_POLY = {
    "VERTEX": dict(dim=0, simplex=True, verts=np.zeros((1, 0)), faces={}),
    "SEGMENT": dict(dim=1, simplex=False, verts=np.array([[0.], [1.]]),
                    faces={0: [[0], [1]]}),
    "QUAD": dict(dim=2, simplex=False, verts=np.array([[0., 0], [1, 0], [0, 1], [1, 1]]),
                 faces={0: [[0], [1], [2], [3]], 1: [[0, 1], [2, 3], [0, 2], [1, 3]]}),
    "TRI": dict(dim=2, simplex=True, verts=np.array([[0., 0], [1, 0], [0, 1]]),
                faces={0: [[0], [1], [2]], 1: [[0, 1], [0, 2], [1, 2]]}),
    "HEX": dict(dim=3, simplex=False,
                verts=np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0],
                                [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]]),
                faces={0: [[i] for i in range(8)],
                       1: [[0, 1], [2, 3], [0, 2], [1, 3], [4, 5], [6, 7], [4, 6], [5, 7],
                           [0, 4], [1, 5], [2, 6], [3, 7]],
                       2: [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 4, 5], [2, 3, 6, 7],
                           [0, 2, 4, 6], [1, 3, 5, 7]]}),
    "TET": dict(dim=3, simplex=True, verts=np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]),
                faces={0: [[0], [1], [2], [3]],
                       1: [[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]],
                       2: [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]]}),
}


class Polytope:
    def __init__(self, name):
        p = _POLY[name]
        self.name, self.dim, self.is_simplex = name, p["dim"], p["simplex"]
        self.verts = p["verts"]
        self._faces = {d: np.array(f, dtype=np.int64) for d, f in p["faces"].items()}

    def lfaces(self, d):
        """(nlf, nvf) local vertex ids of the d-faces."""
        if d == self.dim:
            return np.arange(len(self.verts), dtype=np.int64)[None, :]
        return self._faces[d]

    def nfaces(self, d):
        return len(self.lfaces(d))

    def face_polytope(self, d):
        if d == self.dim:
            return self
        if d == 0:
            return Polytope("VERTEX")
        if d == 1:
            return Polytope("SEGMENT")
        return Polytope("TRI" if self.is_simplex else "QUAD")

    def facet_normals(self):
        d = self.dim
        c = self.verts.mean(axis=0)
        out = np.zeros((self.nfaces(d - 1), d))
        for k, fv in enumerate(self.lfaces(d - 1)):
            p = self.verts[fv]
            if d == 1:
                n = p[0] - c
            elif d == 2:
                t = p[1] - p[0]
                n = np.array([t[1], -t[0]])
            else:
                n = np.cross(p[1] - p[0], p[2] - p[0])
            n = n / np.linalg.norm(n)
            out[k] = n if np.dot(n, p.mean(axis=0) - c) > 0 else -n
        return out

    def facet_tangents(self):
        """Edge vector (second minus first vertex) of the local 1-faces of a 2-D cell."""
        lf = self.lfaces(1)
        return self.verts[lf[:, 1]] - self.verts[lf[:, 0]]


def ncube(d):
    return Polytope({0: "VERTEX", 1: "SEGMENT", 2: "QUAD", 3: "HEX"}[d])


def simplex(d):
    return Polytope({0: "VERTEX", 1: "SEGMENT", 2: "TRI", 3: "TET"}[d])


# ---------------------------------------------------------------------------------------
def _row_groups(keys):
    """Group equal rows. Returns (order, group id of each sorted row, first-row mask)."""
    order = np.lexsort(keys.T[::-1])
    ks = keys[order]
    new = np.ones(len(ks), dtype=bool)
    if len(ks) > 1:
        new[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    return order, np.cumsum(new) - 1, new


def first_encounter_faces(cell_vertices, lfaces):
    """Number the faces spanned by `lfaces` in order of first appearance over
    (cell, local face); a face keeps the vertex list of the first cell that sees it."""
    ncell = len(cell_vertices)
    nlf, nvf = lfaces.shape
    flat = cell_vertices[:, lfaces].reshape(-1, nvf)
    order, grp, new = _row_groups(np.sort(flat, axis=1))
    first_flat = order[new]                     # stable sort => smallest flat index of each group
    rank = np.empty(len(first_flat), dtype=np.int64)
    by_first = np.argsort(first_flat, kind="stable")
    rank[by_first] = np.arange(len(first_flat))
    face_of_flat = np.empty(len(flat), dtype=np.int64)
    face_of_flat[order] = rank[grp]
    return flat[first_flat[by_first]], face_of_flat.reshape(ncell, nlf)


def incidence_lists(cell_faces, nfaces, max_per_face=2):
    """For each face the first `max_per_face` (cell, local face) pairs in ascending cell id
    (-1 padded) and the number of cells around."""
    ncell, nlf = cell_faces.shape
    flat = cell_faces.reshape(-1)
    order = np.argsort(flat, kind="stable")
    fs = flat[order]
    start = np.searchsorted(fs, np.arange(nfaces))
    count = np.searchsorted(fs, np.arange(nfaces), side="right") - start
    cells = np.full((nfaces, max_per_face), -1, dtype=np.int64)
    lfs = np.full((nfaces, max_per_face), -1, dtype=np.int64)
    for k in range(max_per_face):
        has = count > k
        idx = order[start[has] + k]
        cells[has, k] = idx // nlf
        lfs[has, k] = idx % nlf
    return cells, lfs, count


class DiscreteModel:
    """node_coords (nnodes,D); cell_nodes (geometry) and cell_vertices (topology, differs
    only for periodic models); face_vertices[d], cell_faces[d], face_entity[d]; tags."""

    def __init__(self, poly, node_coords, cell_nodes, cell_vertices, nvertices):
        self.poly = poly
        self.D = poly.dim
        self.node_coords = np.ascontiguousarray(node_coords, dtype=np.float64)
        self.cell_nodes = np.ascontiguousarray(cell_nodes, dtype=np.int64)
        self.cell_vertices = np.ascontiguousarray(cell_vertices, dtype=np.int64)
        self.ncells = len(self.cell_nodes)
        self.nvertices = int(nvertices)
        self.face_vertices, self.cell_faces, self.face_entity, self.tags = {}, {}, {}, {}
        self._facet_inc = None
        self._keys = {}

    def build_topology(self):
        D = self.D
        self.face_vertices[0] = np.arange(self.nvertices, dtype=np.int64)[:, None]
        self.cell_faces[0] = self.cell_vertices
        self.face_vertices[D] = self.cell_vertices
        self.cell_faces[D] = np.arange(self.ncells, dtype=np.int64)[:, None]
        for d in range(1, D):
            fv, c2f = first_encounter_faces(self.cell_vertices, self.poly.lfaces(d))
            self.face_vertices[d], self.cell_faces[d] = fv, c2f
        return self

    def nfaces(self, d):
        return len(self.face_vertices[d])

    def facet_incidence(self):
        if self._facet_inc is None:
            d = self.D - 1
            self._facet_inc = incidence_lists(self.cell_faces[d], self.nfaces(d))
        return self._facet_inc

    def face_ids(self, d, vertex_rows):
        """Ids of the d-faces with the given vertex sets (rows, any vertex order)."""
        if d == 0:
            return np.asarray(vertex_rows, dtype=np.int64).reshape(-1)
        q = np.sort(np.asarray(vertex_rows, dtype=np.int64), axis=1)
        ncol = min(q.shape[1], 3)       # three vertices identify a face of a conforming mesh
        nv = self.nvertices
        if float(nv) ** ncol < 2.0 ** 62:
            if d not in self._keys:
                key = self._pack(np.sort(self.face_vertices[d], axis=1), ncol)
                o = np.argsort(key, kind="stable")
                self._keys[d] = (key[o], o)
            ks, o = self._keys[d]
            kq = self._pack(q, ncol)
        else:                            # rank-compress the leading pair to stay inside 64 bits
            t = np.sort(self.face_vertices[d], axis=1)
            kt, kq = t[:, 0] * nv + t[:, 1], q[:, 0] * nv + q[:, 1]
            _, inv = np.unique(np.concatenate([kt, kq]), return_inverse=True)
            kt, kq = inv[:len(kt)] * nv + t[:, 2], inv[len(kt):] * nv + q[:, 2]
            o = np.argsort(kt, kind="stable")
            ks = kt[o]
        pos = np.minimum(np.searchsorted(ks, kq), len(ks) - 1)
        assert np.all(ks[pos] == kq), "face not found"
        return o[pos]

    def _pack(self, rows, ncol):
        key = np.zeros(len(rows), dtype=np.int64)
        for k in range(ncol):
            key = key * self.nvertices + rows[:, k]
        return key

    def face_mask(self, tags, d):
        if isinstance(tags, str):
            tags = [tags]
        ents = np.unique(np.concatenate([np.asarray(self.tags[t], dtype=np.int64) for t in tags]))
        return np.isin(self.face_entity[d], ents)


def add_tag_from_tags(model, name, entities):
    """`add_tag_from_tags!(labels,name,[entity ids])` (src/Yago_freq_domain.jl:112-113)."""
    model.tags[name] = [int(e) for e in entities]


# ---------------------------------------------------------------------------------------
def CartesianDiscreteModel(domain, partition, map=None, isperiodic=None):
    """Nodes/cells lexicographic with x fastest; coordinate = x0 + i*dx then `map` (applied
    to the (nnodes,D) array, vectorised).  Label entities: box n-face index (1-based)."""
    D = len(partition)
    poly = ncube(D)
    n = np.array(partition, dtype=np.int64)
    x0 = np.array(domain[0::2], dtype=np.float64)
    x1 = np.array(domain[1::2], dtype=np.float64)
    dx = (x1 - x0) / n
    per = np.zeros(D, dtype=bool) if isperiodic is None else np.array(isperiodic, dtype=bool)
    nn = n + 1
    idx = [g.reshape(-1, order="F") for g in np.indices(tuple(nn))]
    coords = np.stack([x0[a] + idx[a] * dx[a] for a in range(D)], axis=1)
    if map is not None:
        coords = np.ascontiguousarray(map(coords), dtype=np.float64)
    cidx = [g.reshape(-1, order="F") for g in np.indices(tuple(n))]
    corners = poly.verts.astype(np.int64)
    nvd = np.where(per, n, nn)

    def lin(ix, dims):
        out, stride = 0, 1
        for a in range(D):
            out = out + ix[a] * stride
            stride *= dims[a]
        return out

    cell_nodes = np.stack([lin([cidx[a] + c[a] for a in range(D)], nn) for c in corners], axis=1)
    cell_vertices = np.stack([lin([(cidx[a] + c[a]) % nvd[a] for a in range(D)], nvd)
                              for c in corners], axis=1)
    m = DiscreteModel(poly, coords, cell_nodes, cell_vertices, int(np.prod(nvd)))
    m.partition, m.isperiodic = tuple(int(v) for v in n), tuple(bool(v) for v in per)
    m.build_topology()
    # labels
    vid = np.arange(m.nvertices)
    vpos = np.empty((m.nvertices, D), dtype=np.int64)
    for a in range(D):
        vpos[:, a] = vid % nvd[a]
        vid = vid // nvd[a]
    offsets = np.cumsum([0] + [poly.nfaces(d) for d in range(D)])
    interior = int(offsets[-1]) + 1
    table = np.full(3 ** D, interior, dtype=np.int64)          # signature -> entity
    for d in range(D):
        for k, fv in enumerate(poly.lfaces(d)):
            c = poly.verts[fv]
            code = 0
            for a in range(D):
                s = 0 if np.all(c[:, a] == 0) else (2 if np.all(c[:, a] == 1) else 1)
                code += s * 3 ** a
            table[code] = offsets[d] + k + 1
    for d in range(D + 1):
        p = vpos[m.face_vertices[d]]                           # (nf, nvf, D)
        code = np.zeros(len(p), dtype=np.int64)
        for a in range(D):
            if per[a]:
                s = np.ones(len(p), dtype=np.int64)
            else:
                lo = np.all(p[:, :, a] == 0, axis=1)
                hi = np.all(p[:, :, a] == n[a], axis=1)
                s = np.where(lo, 0, np.where(hi, 2, 1))
            code += s * 3 ** a
        m.face_entity[d] = table[code]
    m.tags = {"interior": [interior], "boundary": list(range(1, interior))}
    return m


def _table(t):
    ptrs = np.asarray(t["ptrs"], dtype=np.int64) - 1
    data = np.asarray(t["data"], dtype=np.int64) - 1
    w = int(ptrs[1] - ptrs[0])
    return data.reshape(-1, w)


def DiscreteModelFromFile(path):
    """Gridap JSON model (`to_json_file`, schema SURVEY.md A5) or a gmsh 4.1 ASCII mesh
    converted with GridapGmsh's numbering (reference: models/multi_geo.jl:5-7)."""
    if path.endswith(".msh"):
        from .gmsh_io import GmshDiscreteModel
        return GmshDiscreteModel(path)
    with open(path) as f:
        j = json.load(f)
    D = j["D"]
    g = j["grid"]
    poly = simplex(D) if all(e == 2 for e in g["reffes"][0]["extrusion"]) else ncube(D)
    coords = np.asarray(g["node_coordinates"], dtype=np.float64).reshape(-1, g["Dp"])
    cells = _table(g["cell_node_ids"])
    m = DiscreteModel(poly, coords, cells, cells, len(j["vertex_node"]))
    m.face_vertices[0] = np.arange(m.nvertices, dtype=np.int64)[:, None]
    m.cell_faces[0] = cells
    m.face_vertices[D] = cells
    m.cell_faces[D] = np.arange(m.ncells, dtype=np.int64)[:, None]
    for d in range(1, D):
        m.face_vertices[d] = _table(j["face_vertices_%d" % d])
        lf = poly.lfaces(d)
        m.cell_faces[d] = m.face_ids(d, cells[:, lf].reshape(-1, lf.shape[1])).reshape(m.ncells, len(lf))
    lab = j["labeling"]
    for d in range(D + 1):
        m.face_entity[d] = np.asarray(lab["entities_%d" % d], dtype=np.int64)
    m.tags = {name: [int(e) for e in ents] for name, ents in zip(lab["names"], lab["tags"])}
    return m


# ---------------------------------------------------------------------------------------
# Triangulations
# ---------------------------------------------------------------------------------------
class Interior:
    """`Interior(model)` / `Triangulation(model)`."""

    def __init__(self, model):
        self.model, self.poly, self.dim, self.ncells = model, model.poly, model.D, model.ncells

    def cell_coords(self):
        m = self.model
        return m.node_coords[m.cell_nodes]


class BoundaryTriangulation:
    """Facets of the background model in a given order, each glued to its first cell."""

    def __init__(self, model, facets):
        self.model = model
        self.dim = model.D - 1
        self.poly = model.poly.face_polytope(self.dim)
        self.facets = np.asarray(facets, dtype=np.int64)
        self.ncells = len(self.facets)
        cells, lfs, _ = model.facet_incidence()
        self.cell = cells[self.facets, 0]
        self.lface = lfs[self.facets, 0]
        self._active = None

    def cell_coords(self):
        """Facet vertex coordinates read through the glued cell (periodic-safe)."""
        m = self.model
        lfv = m.poly.lfaces(self.dim)[self.lface]
        nodes = np.take_along_axis(m.cell_nodes[self.cell], lfv, axis=1)
        return m.node_coords[nodes]

    def parent_cell_coords(self):
        m = self.model
        return m.node_coords[m.cell_nodes[self.cell]]

    def position_in(self, other):
        """Index in `other` of each of this triangulation's cells (same background facet)."""
        o = np.argsort(other.facets, kind="stable")
        pos = np.searchsorted(other.facets[o], self.facets)
        assert np.all(other.facets[o][pos] == self.facets)
        return o[pos]

    def active_model(self):
        """Sub-faces of the triangulation's facets, keeping the parent's relative order."""
        if self._active is not None:
            return self._active
        m, d = self.model, self.dim
        fv = m.face_vertices[d][self.facets]
        act = {"cell_faces": {}, "nfaces": {}, "parent": {}}
        for k in range(d):
            lf = self.poly.lfaces(k)
            if k == 0:
                parent = fv
            else:
                parent = m.face_ids(k, fv[:, lf].reshape(-1, lf.shape[1])).reshape(self.ncells, len(lf))
            touched = np.unique(parent)
            act["parent"][k] = touched
            act["nfaces"][k] = len(touched)
            act["cell_faces"][k] = np.searchsorted(touched, parent)
        act["nfaces"][d] = self.ncells
        act["cell_faces"][d] = np.arange(self.ncells, dtype=np.int64)[:, None]
        self._active = act
        return act


def Boundary(model, tags):
    """`Boundary(model,tags=...)`: tagged facets in ascending facet id."""
    return BoundaryTriangulation(model, np.nonzero(model.face_mask(tags, model.D - 1))[0])


def Triangulation(parent, ids):
    """`Triangulation(Γ,ids)`: keeps the order of `ids` (src/Yago_freq_domain.jl:127-130)."""
    return BoundaryTriangulation(parent.model, parent.facets[np.asarray(ids, dtype=np.int64)])


class Skeleton:
    """`Skeleton(Γ)` / `Skeleton(Γ,mask)`: interior (d-1)-faces of Γ's active model, ascending;
    plus = first surface cell around, minus = second (src/Yago_freq_domain.jl:132,
    src/Khabakhpasheva_freq_domain.jl:173-181)."""

    def __init__(self, surf, face_mask=None):
        self.surf = surf
        act = surf.active_model()
        k = surf.dim - 1
        cells, lfs, count = incidence_lists(act["cell_faces"][k], act["nfaces"][k])
        keep = count == 2
        if face_mask is not None:
            keep &= np.asarray(face_mask, dtype=bool)
        f = np.nonzero(keep)[0]
        self.faces, self.ncells = f, len(f)
        self.plus, self.minus = cells[f, 0], cells[f, 1]
        self.plus_lf, self.minus_lf = lfs[f, 0], lfs[f, 1]
        self.dim = surf.dim
        self.poly = surf.poly

    def cell_measure(self):
        """`get_cell_measure(Λ)`: length of each skeleton edge (src/Multi_geo_freq_domain.jl:82)."""
        X = self.surf.cell_coords()[self.plus]
        if self.dim == 1:
            return np.ones(self.ncells)
        lv = self.poly.lfaces(1)[self.plus_lf]
        a = np.take_along_axis(X, lv[:, 0][:, None, None].repeat(X.shape[2], 2), axis=1)[:, 0]
        b = np.take_along_axis(X, lv[:, 1][:, None, None].repeat(X.shape[2], 2), axis=1)[:, 0]
        return np.linalg.norm(b - a, axis=1)


# ---------------------------------------------------------------------------------------
# Uniform (red) refinement of simplicial models - synthetic scaling meshes (SURVEY.md §8d "M2":
# the MultiGeo tetrahedral mesh refined 1-3 levels; every level multiplies the cells by 2^D)
# ---------------------------------------------------------------------------------------
def refine_uniform(model):
    """One level of regular refinement of a triangular / tetrahedral model: every edge is split at
    its midpoint, a triangle becomes 4 triangles, a tetrahedron 4 corner tetrahedra + 4 from the
    inner octahedron (cut along the midpoint diagonal m02-m13).  Vertices 0..nv-1 are kept, the
    midpoint of edge k becomes vertex nv+k; every new face inherits the entity of the smallest
    face of the parent model that contains it, so `Boundary(model,tags=...)`, `Skeleton`, FE
    spaces and the case modules work unchanged on the refined model."""
    if not model.poly.is_simplex or model.D not in (2, 3):
        raise ValueError("refine_uniform: triangular or tetrahedral models only")
    if not np.array_equal(model.cell_nodes, model.cell_vertices):
        raise ValueError("refine_uniform: periodic models are not supported")
    D, nv = model.D, model.nvertices
    ev = model.face_vertices[1]                                   # (nedges, 2)
    coords = np.concatenate([model.node_coords, 0.5 * (model.node_coords[ev[:, 0]] + model.node_coords[ev[:, 1]])])
    cv = model.cell_vertices
    ce = model.cell_faces[1]                                      # local edge k of a cell = poly.lfaces(1)[k]
    le = [tuple(r) for r in model.poly.lfaces(1)]
    mid = {e: nv + ce[:, k] for k, e in enumerate(le)}            # midpoint vertex of local edge (a,b), a<b
    m = lambda a, b: mid[(a, b) if a < b else (b, a)]
    v = lambda a: cv[:, a]
    if D == 2:
        kids = [(v(0), m(0, 1), m(0, 2)), (m(0, 1), v(1), m(1, 2)), (m(0, 2), m(1, 2), v(2)),
                (m(0, 1), m(1, 2), m(0, 2))]
    else:
        kids = [(v(0), m(0, 1), m(0, 2), m(0, 3)), (m(0, 1), v(1), m(1, 2), m(1, 3)),
                (m(0, 2), m(1, 2), v(2), m(2, 3)), (m(0, 3), m(1, 3), m(2, 3), v(3)),
                # inner octahedron around the diagonal m02-m13
                (m(0, 2), m(1, 3), m(0, 1), m(0, 3)), (m(0, 2), m(1, 3), m(0, 3), m(2, 3)),
                (m(0, 2), m(1, 3), m(2, 3), m(1, 2)), (m(0, 2), m(1, 3), m(1, 2), m(0, 1))]
    cells = np.stack([np.stack(k, axis=1) for k in kids], axis=1).reshape(-1, D + 1)    # children of a cell together
    parent_cell = np.repeat(np.arange(model.ncells), len(kids))
    # GridapGmsh convention of the reference's models: simplex vertex ids ascending
    cells = np.sort(cells, axis=1)
    fine = DiscreteModel(model.poly, coords, cells, cells, len(coords)).build_topology()
    # support of every new vertex in the parent model: the old vertices it is a combination of
    sup = np.full((len(coords), 2), -1, dtype=np.int64)
    sup[:nv, 0] = np.arange(nv)
    sup[nv:] = ev
    first_cell = {}
    for d in range(1, D + 1):                                     # a fine cell that contains each fine d-face
        c2f = fine.cell_faces[d]
        fc = np.full(fine.nfaces(d), -1, dtype=np.int64)
        fc[c2f[::-1].reshape(-1)] = np.repeat(np.arange(fine.ncells)[::-1], c2f.shape[1])
        first_cell[d] = fc
    fine.face_entity[0] = np.concatenate([model.face_entity[0], model.face_entity[1]])
    for d in range(1, D + 1):
        fv = fine.face_vertices[d]
        s = sup[fv].reshape(len(fv), -1)                          # old vertices spanned (with -1 padding)
        s = np.sort(s, axis=1)[:, ::-1]                           # descending: real ids first
        ent = np.zeros(len(fv), dtype=np.int64)
        nuniq = np.array([len(np.unique(r[r >= 0])) for r in s]) if len(s) < 200000 else None
        if nuniq is None:                                         # vectorised count of distinct ids
            ss = np.sort(s, axis=1)
            nuniq = (ss >= 0).sum(axis=1) - ((ss[:, 1:] == ss[:, :-1]) & (ss[:, 1:] >= 0)).sum(axis=1)
        for k in range(1, D + 2):                                 # k distinct old vertices -> old (k-1)-face
            sel = np.nonzero(nuniq == k)[0]
            if len(sel) == 0:
                continue
            rows = np.array([np.unique(r[r >= 0]) for r in s[sel]]).reshape(len(sel), k)
            if k - 1 == D:
                ent[sel] = model.face_entity[D][parent_cell[first_cell[d][sel]]]
            elif k == 1:
                ent[sel] = model.face_entity[0][rows[:, 0]]
            else:
                ent[sel] = model.face_entity[k - 1][model.face_ids(k - 1, rows)]
        fine.face_entity[d] = ent
    fine.tags = {name: list(ents) for name, ents in model.tags.items()}
    return fine
