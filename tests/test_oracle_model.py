"""Oracle pinned on what the reference holds: the golden mesh pair, entity ids used by the
scripts, and the problem sizes implied by the reference parameters."""
import os
import numpy as np
from oracle.model import json_model, cartesian_model
from oracle.gmsh import gmsh_model
from oracle.quadrature import quadrature
from oracle import polytopes as pt
from oracle.reffe import lagrangian


def test_golden_mesh_pair(golden_dir):
    j = json_model(os.path.join(golden_dir, "multi_geo_coarse.json"))
    g = gmsh_model(os.path.join(golden_dir, "multi_geo_coarse.msh"))
    assert np.array_equal(j.node_coords, g.node_coords)
    assert np.array_equal(j.cell_nodes, g.cell_nodes)
    assert j.ncells == 132 and len(j.node_coords) == 57
    for d in (1, 2):
        assert np.array_equal(j.face_vertices[d], g.face_vertices[d])
    assert j.nfaces(1) == 241 and j.nfaces(2) == 317
    for d in range(4):
        assert np.array_equal(j.face_entity[d], g.face_entity[d])
    assert j.tags == g.tags
    assert j.tags["plate"] == [61, 62, 64, 66] and j.tags["inlet"] == [71]


def test_first_encounter_numbering_reproduces_json(golden_dir):
    j = json_model(os.path.join(golden_dir, "multi_geo_coarse.json"))
    fv = {d: j.face_vertices[d].copy() for d in (1, 2)}
    j.build_topology()
    for d in (1, 2):
        assert np.array_equal(j.face_vertices[d], fv[d])


def test_cartesian_entities_2d_3d():
    m = cartesian_model((0, 1, 0, 1), (3, 2))
    # edges: 5 bottom, 6 top, 7 left, 8 right, 9 interior (src/Khabakhpasheva_freq_domain.jl:108-112)
    x = m.node_coords[m.face_vertices[1]]
    ent = m.face_entity[1]
    assert np.all(ent[np.all(x[:, :, 1] == 0, axis=1)] == 5)
    assert np.all(ent[np.all(x[:, :, 1] == 1, axis=1)] == 6)
    assert np.all(ent[np.all(x[:, :, 0] == 0, axis=1)] == 7)
    assert np.all(ent[np.all(x[:, :, 0] == 1, axis=1)] == 8)
    assert set(m.face_entity[0][[0, 3, 8, 11]]) == {1, 2, 3, 4}
    m3 = cartesian_model((0, 1, 0, 1, 0, 1), (2, 2, 2))
    x = m3.node_coords[m3.face_vertices[2]]
    ent = m3.face_entity[2]
    assert np.all(ent[np.all(x[:, :, 2] == 1, axis=1)] == 22)      # "surface" (src/Yago_freq_domain.jl:112)
    assert np.all(ent[np.all(x[:, :, 0] == 0, axis=1)] == 25)      # "inlet"   (:113)
    assert np.all(m3.face_entity[3] == 27)


def test_periodic_topology_counts():
    n = 3
    m = cartesian_model((0, 2 * np.pi, 0, 1), (2 * n, n), isperiodic=(True, False))
    assert m.nvertices == 2 * n * (n + 1)
    assert m.nfaces(1) == 2 * n * (n + 1) + 2 * n * n          # horizontal + vertical edges


def test_quadrature_exactness():
    import math
    p, w = quadrature(pt.TRI, 8)
    assert abs((w * p[:, 0] ** 2 * p[:, 1] ** 3).sum() - 2 * 6 / math.factorial(7)) < 1e-16
    p, w = quadrature(pt.TET, 8)
    assert abs((w * p[:, 0] ** 2 * p[:, 1] ** 3 * p[:, 2] ** 3).sum() - 72 / math.factorial(11)) < 1e-18
    p, w = quadrature(pt.HEX, 8)
    assert len(w) == 125 and abs(w.sum() - 1) < 1e-14


def test_lagrangian_kronecker_and_order():
    for P, o in [(pt.SEGMENT, 4), (pt.QUAD, 4), (pt.HEX, 2), (pt.TRI, 4), (pt.TET, 2)]:
        r = lagrangian(P, o)
        assert np.abs(r.values(r.nodes) - np.eye(r.ndofs)).max() < 1e-11
    r = lagrangian(pt.QUAD, 2)      # vertices, edge midpoints in edge order, centre
    assert np.allclose(r.nodes[4:], [[.5, 0], [.5, 1], [0, .5], [1, .5], [.5, .5]])


def test_yago_sizes_match_reference_parameters():
    from oracle.cases.yago import YagoCase
    y = YagoCase(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3)   # scripts/5-4-1-Yago.jl:17-30
    assert y.Om.ncells == 720 and y.X.ndofs == 11988
    assert (y.Gb.ncells, y.Gf.ncells, y.Gin.ncells, y.Lb.nfaces) == (4, 716, 36, 4)
    # graded z-levels H - H/2.5^i (src/Yago_freq_domain.jl:100-106)
    y3 = YagoCase(nx=1, ny=1, nz=3, order=2)
    z = np.unique(y3.model.node_coords[:, 2])
    assert np.allclose(z, [0, 35.1, 49.14, 58.5])
