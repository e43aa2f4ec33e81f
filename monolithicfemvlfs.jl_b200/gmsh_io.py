"""gmsh 4.1 ASCII `.msh` -> discrete model with GridapGmsh's numbering.

Mirrors the offline conversion of the reference (`models/multi_geo.jl:5-7`:
`GmshDiscreteModel("multi_geo.msh")` -> `to_json_file`), needed because the fine
`models/multi_geo.json` of `scripts/5-5-1-MultiGeo.jl:32` is not shipped.  Rules (pinned by the
reference's golden pair multi_geo_coarse.msh <-> .json): nodes and top-dimensional cells keep the
file order; simplex node ids are sorted ascending (oriented model); lower-dimensional faces are
numbered by first encounter; the label entity of a face is its gmsh entity tag plus the sum of
the largest entity tags of the lower dimensions; a face that is not listed in $Elements
inherits the smallest label among the faces one dimension up that contain it.
"""
import numpy as np
from .geometry import DiscreteModel, simplex

_NODES_OF = {15: 1, 1: 2, 2: 3, 4: 4}
_DIM_OF = {15: 0, 1: 1, 2: 2, 4: 3}


def _sections(path):
    with open(path) as f:
        text = f.read().split("\n")
    out, name, start = {}, None, 0
    for i, ln in enumerate(text):
        s = ln.strip()
        if s.startswith("$End"):
            out[name] = text[start:i]
            name = None
        elif s.startswith("$"):
            name, start = s[1:], i + 1
    return out


def read_msh(path):
    sec = _sections(path)
    phys = {}
    for ln in sec.get("PhysicalNames", [])[1:]:
        d, tag, nm = ln.split(None, 2)
        phys[(int(d), int(tag))] = nm.strip().strip('"')
    ent = sec["Entities"]
    counts = [int(v) for v in ent[0].split()]
    ent_phys = [{} for _ in range(4)]
    k = 1
    for d in range(4):
        off = 4 if d == 0 else 7
        for _ in range(counts[d]):
            t = ent[k].split()
            k += 1
            n = int(t[off])
            ent_phys[d][int(t[0])] = [int(v) for v in t[off + 1:off + 1 + n]]
    nd = sec["Nodes"]
    nblocks, nnodes = (int(v) for v in nd[0].split()[:2])
    coords = np.empty((nnodes, 3))
    node_ent = np.empty((nnodes, 2), dtype=np.int64)
    k, filled = 1, 0
    for _ in range(nblocks):
        edim, etag, _, nb = (int(v) for v in nd[k].split())
        k += 1
        tags = np.array([int(v) for v in nd[k:k + nb]], dtype=np.int64)
        k += nb
        if nb:
            assert np.array_equal(tags, np.arange(filled + 1, filled + nb + 1)), "non-contiguous node tags"
            coords[filled:filled + nb] = np.array([ln.split() for ln in nd[k:k + nb]], dtype=np.float64)
            node_ent[filled:filled + nb] = (edim, etag)
        k += nb
        filled += nb
    el = sec["Elements"]
    nblocks = int(el[0].split()[0])
    elems = {d: ([], []) for d in range(4)}
    k = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in el[k].split())
        k += 1
        if nb:
            rows = np.array([ln.split() for ln in el[k:k + nb]], dtype=np.int64)
            elems[_DIM_OF[etype]][0].append(np.full(nb, etag, dtype=np.int64))
            elems[_DIM_OF[etype]][1].append(rows[:, 1:1 + _NODES_OF[etype]] - 1)
        k += nb
    elements = {}
    for d in range(4):
        if elems[d][0]:
            elements[d] = (np.concatenate(elems[d][0]), np.concatenate(elems[d][1]))
    return dict(physical_names=phys, entity_physical=ent_phys, nodes=coords, node_entity=node_ent,
                elements=elements)


def GmshDiscreteModel(path):
    msh = read_msh(path)
    D = max(msh["elements"])
    coords = msh["nodes"][:, :3] if D == 3 else msh["nodes"][:, :D]
    cells = np.sort(msh["elements"][D][1], axis=1)
    m = DiscreteModel(simplex(D), coords, cells, cells, len(coords))
    m.build_topology()
    maxtag = [max(msh["entity_physical"][d].keys(), default=0) for d in range(4)]
    offset = np.concatenate([[0], np.cumsum(maxtag)])[:4]
    ne = msh["node_entity"]
    m.face_entity[0] = ne[:, 1] + offset[ne[:, 0]]
    UNSET = np.iinfo(np.int64).max
    for d in range(1, D + 1):
        ent = np.full(m.nfaces(d), UNSET, dtype=np.int64)
        if d in msh["elements"]:
            etag, nodes = msh["elements"][d]
            if d == D:
                ent[:] = etag + offset[d]
            else:
                ent[m.face_ids(d, nodes)] = etag + offset[d]
        m.face_entity[d] = ent
    for d in range(D - 1, 0, -1):
        ent, up = m.face_entity[d], m.face_entity[d + 1]
        best = np.full(len(ent), UNSET, dtype=np.int64)
        lf = simplex(d + 1).lfaces(d)
        fv = m.face_vertices[d + 1]
        for lv in lf:
            np.minimum.at(best, m.face_ids(d, fv[:, lv]), up)
        unset = ent == UNSET
        ent[unset] = best[unset]
    tags = {}
    for (d, ptag), name in sorted(msh["physical_names"].items(), key=lambda kv: kv[0][1]):
        tags[name] = [int(etag + offset[d]) for etag, ph in sorted(msh["entity_physical"][d].items())
                      if ptag in ph]
    m.tags = tags
    return m
