"""Discrete models: Cartesian (mapped / periodic) and Gridap-JSON (oracle).

Restates what the reference obtains from Gridap at
`CartesianDiscreteModel(domain,partition,map=map)` (`src/Yago_freq_domain.jl:98-108`,
`src/Khabakhpasheva_freq_domain.jl:94-104`), `...isperiodic=(true,false)`
(`src/Periodic_Beam.jl:60-62`) and `DiscreteModelFromFile(mesh_file)`
(`src/Multi_geo_freq_domain.jl:66`).  Semantics: SURVEY.md A1-A5.  All ids are
0-based here (Gridap: 1-based).

Topology rule (verified on `models/multi_geo_coarse.json`, assumed for the
Cartesian models): d-faces are numbered by FIRST ENCOUNTER looping cells in
order and, inside a cell, local d-faces in the reference-polytope order; the
stored vertex list of a face is the first cell's.
"""
import json
import numpy as np
from . import polytopes as pt


class DiscreteModel:
    """Plain container.

    node_coords   (nnodes, D)          grid nodes (geometry)
    cell_nodes    (ncells, nv)         grid node ids per cell (geometry)
    cell_vertices (ncells, nv)         topological vertex ids (== cell_nodes
                                       unless periodic)
    face_vertices[d] (nfaces_d, nvf)   vertex ids of the d-faces, d = 0..D
    cell_faces[d]    (ncells, nlf_d)   d-face ids of each cell
    face_entity[d]   (nfaces_d,)       label entity of each d-face
    tags             name -> list of entity ids
    """

    def __init__(self, poly, node_coords, cell_nodes, cell_vertices, nvertices):
        self.poly = poly
        self.D = poly.dim
        self.node_coords = np.asarray(node_coords, dtype=float)
        self.cell_nodes = np.asarray(cell_nodes, dtype=np.int64)
        self.cell_vertices = np.asarray(cell_vertices, dtype=np.int64)
        self.ncells = len(self.cell_nodes)
        self.nvertices = nvertices
        self.face_vertices = {}
        self.cell_faces = {}
        self.face_entity = {}
        self.tags = {}

    # -- topology ---------------------------------------------------------
    def build_topology(self):
        D = self.D
        cv = self.cell_vertices
        self.face_vertices[0] = np.arange(self.nvertices, dtype=np.int64)[:, None]
        self.cell_faces[0] = cv.copy()
        self.face_vertices[D] = cv.copy()
        self.cell_faces[D] = np.arange(self.ncells, dtype=np.int64)[:, None]
        for d in range(1, D):
            lfaces = self.poly.faces[d]
            seen = {}
            fverts = []
            c2f = np.empty((self.ncells, len(lfaces)), dtype=np.int64)
            for c in range(self.ncells):
                v = cv[c]
                for lf, lv in enumerate(lfaces):
                    fv = tuple(int(v[i]) for i in lv)
                    key = tuple(sorted(fv))
                    f = seen.get(key)
                    if f is None:
                        f = len(fverts)
                        seen[key] = f
                        fverts.append(fv)
                    c2f[c, lf] = f
            self.face_vertices[d] = np.array(fverts, dtype=np.int64)
            self.cell_faces[d] = c2f
        return self

    def nfaces(self, d):
        return len(self.face_vertices[d])

    def cells_around(self, d):
        """List (per d-face) of (cell, local face) pairs in ascending cell id."""
        out = [[] for _ in range(self.nfaces(d))]
        c2f = self.cell_faces[d]
        for c in range(self.ncells):
            for lf, f in enumerate(c2f[c]):
                out[f].append((c, lf))
        return out

    def face_mask(self, tag_names, d):
        if isinstance(tag_names, str):
            tag_names = [tag_names]
        ents = set()
        for t in tag_names:
            ents.update(self.tags[t])
        return np.isin(self.face_entity[d], list(ents))

    def add_tag_from_entities(self, name, entities):
        self.tags[name] = list(entities)


# ---------------------------------------------------------------------------
def cartesian_model(domain, partition, map_fn=None, isperiodic=None):
    """`CartesianDiscreteModel(domain, partition; map, isperiodic)`.

    Node coordinate along axis a: x0[a] + i*dx[a], dx = (x1-x0)/n (one multiply,
    one add: the reference relies on exact tests such as `x == H`
    (`src/Yago_freq_domain.jl:101`) and `xi[2]==0.0` (`:183`)).  `map_fn` moves
    nodes only; cells stay straight-sided Q1.
    """
    D = len(partition)
    poly = pt.ncube(D)
    n = np.array(partition, dtype=np.int64)
    x0 = np.array(domain[0::2], dtype=float)
    x1 = np.array(domain[1::2], dtype=float)
    dx = (x1 - x0) / n
    per = np.array(isperiodic if isperiodic is not None else (False,) * D)
    nn = n + 1
    idx = [g.reshape(-1, order="F") for g in np.indices(tuple(nn))]
    coords = np.stack([x0[a] + idx[a] * dx[a] for a in range(D)], axis=1)
    if map_fn is not None:
        coords = np.array([map_fn(p) for p in coords], dtype=float)

    def lin(ix, dims):
        out = np.zeros_like(ix[0])
        stride = 1
        for a in range(D):
            out = out + ix[a] * stride
            stride *= dims[a]
        return out

    cidx = [g.reshape(-1, order="F") for g in np.indices(tuple(n))]
    corners = poly.verts.astype(np.int64)
    cell_nodes = np.stack([lin([cidx[a] + c[a] for a in range(D)], nn)
                           for c in corners], axis=1)
    nv_dims = np.where(per, n, nn)
    cell_vertices = np.stack(
        [lin([(cidx[a] + c[a]) % nv_dims[a] if per[a] else cidx[a] + c[a]
              for a in range(D)], nv_dims) for c in corners], axis=1)
    m = DiscreteModel(poly, coords, cell_nodes, cell_vertices, int(np.prod(nv_dims)))
    m.partition = tuple(int(v) for v in n)
    m.isperiodic = tuple(bool(v) for v in per)
    m.build_topology()
    _cartesian_labels(m, n, per)
    return m


def _cartesian_labels(m, n, per):
    """Entity of a d-face = index (1-based, over all dims) of the n-face of the
    bounding box it lies on; box interior = last id (9 in 2-D, 27 in 3-D)."""
    D = m.D
    poly = m.poly
    # lattice position of each vertex
    nv_dims = np.where(per, n, n + 1)
    vid = np.arange(m.nvertices)
    vpos = []
    for a in range(D):
        vpos.append(vid % nv_dims[a])
        vid = vid // nv_dims[a]
    vpos = np.stack(vpos, axis=1)                      # (nvertices, D)
    offsets = np.cumsum([0] + [poly.nfaces(d) for d in range(D)])
    interior = int(offsets[-1]) + 1
    # signature of each box n-face: per axis -1 (lower), +1 (upper), 0 (spans)
    sig2ent = {}
    for d in range(D):
        for iface, fv in enumerate(poly.face_vertices(d)):
            c = poly.verts[list(fv)]
            sig = tuple(int(-1 if np.all(c[:, a] == 0) else (1 if np.all(c[:, a] == 1) else 0))
                        for a in range(D))
            sig2ent[sig] = int(offsets[d]) + iface + 1
    sig2ent[(0,) * D] = interior
    for d in range(D + 1):
        fv = m.face_vertices[d]
        ent = np.empty(len(fv), dtype=np.int64)
        for f in range(len(fv)):
            p = vpos[fv[f]]
            sig = []
            for a in range(D):
                if per[a]:
                    sig.append(0)
                elif np.all(p[:, a] == 0):
                    sig.append(-1)
                elif np.all(p[:, a] == n[a]):
                    sig.append(1)
                else:
                    sig.append(0)
            ent[f] = sig2ent[tuple(sig)]
        m.face_entity[d] = ent
    m.tags = {"interior": [interior],
              "boundary": list(range(1, interior))}


# ---------------------------------------------------------------------------
def _table(t):
    ptrs = np.asarray(t["ptrs"], dtype=np.int64) - 1
    data = np.asarray(t["data"], dtype=np.int64) - 1
    w = ptrs[1] - ptrs[0]
    assert np.all(np.diff(ptrs) == w)
    return data.reshape(-1, w)


def json_model(path):
    """Gridap `to_json_file` model (schema: SURVEY.md A5); simplices only."""
    with open(path) as f:
        j = json.load(f)
    D = j["D"]
    g = j["grid"]
    ext = g["reffes"][0]["extrusion"]
    poly = pt.simplex(D) if all(e == 2 for e in ext) else pt.ncube(D)
    coords = np.asarray(g["node_coordinates"], dtype=float).reshape(-1, g["Dp"])
    cells = _table(g["cell_node_ids"])
    vertex_node = np.asarray(j["vertex_node"], dtype=np.int64) - 1
    assert np.array_equal(vertex_node, np.arange(len(vertex_node)))
    m = DiscreteModel(poly, coords, cells, cells, len(vertex_node))
    m.face_vertices[0] = np.arange(m.nvertices, dtype=np.int64)[:, None]
    m.cell_faces[0] = cells.copy()
    for d in range(1, D + 1):
        m.face_vertices[d] = _table(j["face_vertices_%d" % d])
    m.cell_faces[D] = np.arange(m.ncells, dtype=np.int64)[:, None]
    for d in range(1, D):
        lookup = {tuple(sorted(v)): i for i, v in enumerate(m.face_vertices[d].tolist())}
        lfaces = poly.faces[d]
        c2f = np.empty((m.ncells, len(lfaces)), dtype=np.int64)
        for c in range(m.ncells):
            for lf, lv in enumerate(lfaces):
                c2f[c, lf] = lookup[tuple(sorted(int(cells[c, i]) for i in lv))]
        m.cell_faces[d] = c2f
    lab = j["labeling"]
    for d in range(D + 1):
        m.face_entity[d] = np.asarray(lab["entities_%d" % d], dtype=np.int64)
    m.tags = {name: list(ents) for name, ents in zip(lab["names"], lab["tags"])}
    return m
