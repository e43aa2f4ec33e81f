"""Reference polytopes in Gridap's extrusion ordering (oracle; test infrastructure).

Gridap (third party, pinned by `/root/reference/Project.toml:24`) numbers the
n-faces of an extrusion polytope by a recursive "anchor/extrusion" walk.  The
tables below are that walk's result (0-based here).  They are consistent with
the reference's own data: TRI/TET tables reproduce `face_vertices_1/2` of
`models/multi_geo_coarse.json`; the Cartesian entity ids implied by the n-cube
tables (2-D: corners 1-4, edges 5-8; 3-D: faces 21-26 = z-,z+,y-,y+,x-,x+) match
the tags used at `src/Khabakhpasheva_freq_domain.jl:108-112` and
`src/Yago_freq_domain.jl:112-113`.
"""
import numpy as np


class Polytope:
    def __init__(self, name, dim, verts, faces, is_simplex):
        self.name = name
        self.dim = dim
        self.verts = np.asarray(verts, dtype=float)          # (nv, dim)
        # faces[d] = list of vertex-id tuples of the d-faces, 0 <= d < dim
        self.faces = faces
        self.is_simplex = is_simplex

    def nfaces(self, d):
        if d == self.dim:
            return 1
        return len(self.faces[d])

    def face_vertices(self, d):
        if d == self.dim:
            return [tuple(range(len(self.verts)))]
        return self.faces[d]

    def facet_normals(self):
        """Outward unit normals of the (dim-1)-faces in reference space."""
        d = self.dim
        c = self.verts.mean(axis=0)
        out = []
        for fv in self.faces[d - 1]:
            p = self.verts[list(fv)]
            if d == 1:
                n = np.array([p[0][0] - c[0]])
            elif d == 2:
                t = p[1] - p[0]
                n = np.array([t[1], -t[0]])
            else:
                n = np.cross(p[1] - p[0], p[2] - p[0])
            n = n / np.linalg.norm(n)
            if np.dot(n, p.mean(axis=0) - c) < 0:
                n = -n
            out.append(n)
        return np.array(out)


def _vertex_faces(n):
    return [(i,) for i in range(n)]


VERTEX = Polytope("VERTEX", 0, np.zeros((1, 0)), {}, True)

SEGMENT = Polytope("SEGMENT", 1, [[0.0], [1.0]], {0: _vertex_faces(2)}, False)

QUAD = Polytope(
    "QUAD", 2, [[0, 0], [1, 0], [0, 1], [1, 1]],
    {0: _vertex_faces(4), 1: [(0, 1), (2, 3), (0, 2), (1, 3)]}, False)

HEX = Polytope(
    "HEX", 3,
    [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0],
     [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
    {0: _vertex_faces(8),
     1: [(0, 1), (2, 3), (0, 2), (1, 3), (4, 5), (6, 7), (4, 6), (5, 7),
         (0, 4), (1, 5), (2, 6), (3, 7)],
     2: [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 4, 5), (2, 3, 6, 7),
         (0, 2, 4, 6), (1, 3, 5, 7)]}, False)

TRI = Polytope(
    "TRI", 2, [[0, 0], [1, 0], [0, 1]],
    {0: _vertex_faces(3), 1: [(0, 1), (0, 2), (1, 2)]}, True)

TET = Polytope(
    "TET", 3, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
    {0: _vertex_faces(4),
     1: [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3), (2, 3)],
     2: [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]}, True)


def facet_polytope(p):
    return {"SEGMENT": VERTEX, "QUAD": SEGMENT, "TRI": SEGMENT,
            "HEX": QUAD, "TET": TRI}[p.name]


def ncube(dim):
    return {1: SEGMENT, 2: QUAD, 3: HEX}[dim]


def simplex(dim):
    return {1: SEGMENT, 2: TRI, 3: TET}[dim]
