"""gmsh 4.1 ASCII `.msh` -> discrete model, GridapGmsh numbering (oracle).

Restates the conversion the reference performs offline at
`models/multi_geo.jl:5-7` (`GmshDiscreteModel(...)` -> `to_json_file`).  Pinned
by the golden pair `models/multi_geo_coarse.msh` <-> `models/multi_geo_coarse.json`
(tests/test_oracle_model.py): node order = msh node order, cell order = msh
order of the D-dimensional elements, simplex vertex ids sorted ascending,
faces by first encounter, label entity = gmsh entity tag + sum of the maximum
entity tags of the lower dimensions.
"""
import numpy as np
from . import polytopes as pt
from .model import DiscreteModel

_ELEM_NODES = {15: 1, 1: 2, 2: 3, 4: 4}
_ELEM_DIM = {15: 0, 1: 1, 2: 2, 4: 3}


def read_msh(path):
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    sec = {}
    i = 0
    while i < len(lines):
        if lines[i].startswith("$") and not lines[i].startswith("$End"):
            name = lines[i][1:]
            j = lines.index("$End" + name, i)
            sec[name] = lines[i + 1:j]
            i = j
        i += 1
    out = {}
    # physical names
    phys = {}
    for ln in sec.get("PhysicalNames", [])[1:]:
        d, tag, name = ln.split(None, 2)
        phys[(int(d), int(tag))] = name.strip('"')
    out["physical_names"] = phys
    # entities: per dim, tag -> physical tags
    ent = sec["Entities"]
    counts = [int(v) for v in ent[0].split()]
    ent_phys = [dict() for _ in range(4)]
    k = 1
    for d in range(4):
        for _ in range(counts[d]):
            t = ent[k].split()
            k += 1
            tag = int(t[0])
            off = 4 if d == 0 else 7
            nph = int(t[off])
            ent_phys[d][tag] = [int(v) for v in t[off + 1:off + 1 + nph]]
    out["entity_physical"] = ent_phys
    # nodes
    nd = sec["Nodes"]
    nblocks, nnodes = int(nd[0].split()[0]), int(nd[0].split()[1])
    tags, coords, node_ent = [], [], []
    k = 1
    for _ in range(nblocks):
        edim, etag, _, nb = (int(v) for v in nd[k].split())
        k += 1
        tags.extend(int(nd[k + r]) for r in range(nb))
        node_ent.extend([(edim, etag)] * nb)
        k += nb
        coords.extend([float(v) for v in nd[k + r].split()] for r in range(nb))
        k += nb
    tags = np.array(tags)
    assert np.array_equal(tags, np.arange(1, nnodes + 1)), "non-contiguous node tags"
    out["nodes"] = np.array(coords, dtype=float)
    out["node_entity"] = np.array(node_ent, dtype=np.int64)     # (dim, tag) per node
    # elements
    el = sec["Elements"]
    nblocks = int(el[0].split()[0])
    elems = {d: [] for d in range(4)}       # d -> list of (entity tag, node ids 0-based)
    k = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in el[k].split())
        k += 1
        nn = _ELEM_NODES[etype]
        for r in range(nb):
            t = [int(v) for v in el[k + r].split()]
            elems[_ELEM_DIM[etype]].append((etag, [v - 1 for v in t[1:1 + nn]]))
        k += nb
    out["elements"] = elems
    return out


def gmsh_model(path):
    msh = read_msh(path)
    D = max(d for d in range(4) if msh["elements"][d])
    poly = pt.simplex(D)
    coords = msh["nodes"][:, :3] if D == 3 else msh["nodes"][:, :D]
    cells = np.array([sorted(n) for _, n in msh["elements"][D]], dtype=np.int64)
    m = DiscreteModel(poly, coords, cells, cells, len(coords))
    m.build_topology()
    # entity offsets per dimension
    maxtag = [max(msh["entity_physical"][d].keys(), default=0) for d in range(4)]
    offset = np.concatenate([[0], np.cumsum(maxtag)])[:4]
    UNSET = -1
    # vertices are classified by the entity block of their node in $Nodes
    ndim, ntag = msh["node_entity"][:, 0], msh["node_entity"][:, 1]
    node_label = ntag + offset[ndim]
    m.face_entity[0] = node_label.copy()
    lookups = {d: {tuple(sorted(v)): i for i, v in enumerate(m.face_vertices[d].tolist())}
               for d in range(1, D + 1)}
    for d in range(1, D + 1):
        ent = np.full(m.nfaces(d), UNSET, dtype=np.int64)
        for etag, nodes in msh["elements"][d]:
            f = lookups[d].get(tuple(sorted(nodes)))
            if f is not None:
                ent[f] = etag + offset[d]
        m.face_entity[d] = ent
    # a d-face absent from $Elements takes the smallest label found on the
    # (d+1)-faces around it (a boundary surface wins over the volume), top down
    for d in range(D - 1, 0, -1):
        ent = m.face_entity[d]
        up = m.face_entity[d + 1]
        best = np.full(len(ent), np.iinfo(np.int64).max)
        for lv in pt.simplex(d + 1).face_vertices(d):
            sub = np.sort(m.face_vertices[d + 1][:, list(lv)], axis=1)
            ids = np.array([lookups[d][tuple(r)] for r in sub.tolist()])
            np.minimum.at(best, ids, up)
        unset = ent == UNSET
        ent[unset] = best[unset]
    tags = {}
    for (d, ptag), name in sorted(msh["physical_names"].items(), key=lambda kv: kv[0][1]):
        ents = [etag + offset[d] for etag, ph in sorted(msh["entity_physical"][d].items())
                if ptag in ph]
        tags[name] = [int(e) for e in ents]
    m.tags = tags
    return m
