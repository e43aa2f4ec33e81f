"""Host-side mirror (vectorised front end + descriptor marshalling) against the oracle's
independent loop-based front end and literal weak forms - CPU only."""
import os
import numpy as np
import pytest
import vlfs_b200
from vlfs_b200 import geometry as G, fe as F
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
from oracle.descriptor_eval import RecordingSystem
from oracle.cases.yago import YagoCase
from oracle import model as om, fem as ofem, reffe as orf, polytopes as opt, quadrature as oq


def test_reference_tables_match_oracle():
    rng = np.random.default_rng(0)
    for name, o in [("SEGMENT", 4), ("QUAD", 4), ("QUAD", 2), ("HEX", 2), ("TRI", 4), ("TET", 2)]:
        P = G.Polytope(name)
        a = F.get_reffe(P, o)
        b = orf.lagrangian(getattr(opt, name), o)
        assert np.allclose(a.nodes, b.nodes, atol=1e-15)
        x = rng.random((7, P.dim)) * (0.3 if P.is_simplex else 1.0)
        v, g, h = a.tabulate(x, need_hess=True)
        assert np.abs(v - b.values(x)).max() < 1e-11
        assert np.abs(g - b.gradients(x)).max() < 1e-9
        assert np.abs(h - b.hessians(x)).max() < 1e-7
    for name, deg in [("QUAD", 8), ("HEX", 8), ("TRI", 8), ("TET", 4), ("SEGMENT", 4)]:
        xa, wa = F.quadrature(G.Polytope(name), deg)
        xb, wb = oq.quadrature(getattr(opt, name), deg)
        assert np.allclose(xa, xb, atol=1e-15) and np.allclose(wa, wb, atol=1e-16)


def test_cartesian_topology_matches_oracle():
    for kw in [dict(domain=(0, 3, 0, 2), partition=(4, 3)),
               dict(domain=(0, 3, 0, 2, 0, 1), partition=(3, 2, 2)),
               dict(domain=(0, 1, 0, 1), partition=(6, 3), isperiodic=(True, False))]:
        a = G.CartesianDiscreteModel(kw["domain"], kw["partition"], isperiodic=kw.get("isperiodic"))
        b = om.cartesian_model(kw["domain"], kw["partition"], isperiodic=kw.get("isperiodic"))
        assert np.array_equal(a.cell_nodes, b.cell_nodes) and np.array_equal(a.cell_vertices, b.cell_vertices)
        for d in range(1, a.D):
            assert np.array_equal(a.face_vertices[d], b.face_vertices[d])
            assert np.array_equal(a.cell_faces[d], b.cell_faces[d])
        for d in range(a.D + 1):
            assert np.array_equal(a.face_entity[d], b.face_entity[d])


def test_json_model_matches_oracle(golden_dir):
    a = G.DiscreteModelFromFile(os.path.join(golden_dir, "multi_geo_coarse.json"))
    b = om.json_model(os.path.join(golden_dir, "multi_geo_coarse.json"))
    for d in (1, 2):
        assert np.array_equal(a.face_vertices[d], b.face_vertices[d])
        assert np.array_equal(a.cell_faces[d], b.cell_faces[d])
    ga = G.Boundary(a, tags="plate")
    gb = ofem.boundary(b, "plate")
    assert np.array_equal(ga.facets, gb.facets) and np.array_equal(ga.cell, gb.cell)
    assert np.array_equal(ga.lface, gb.lface)
    sa, sb = G.Skeleton(ga), ofem.SkeletonTriangulation(gb)
    assert np.array_equal(sa.plus, sb.plus) and np.array_equal(sa.minus, sb.minus)
    assert np.array_equal(sa.plus_lf, sb.plus_lf) and np.array_equal(sa.minus_lf, sb.minus_lf)
    va, vb = F.FESpace(ga, 4), ofem.FESpace(gb, 4)
    assert va.ndofs == vb.ndofs and np.array_equal(va.cell_dofs, vb.cell_dofs)


@pytest.mark.parametrize("kw", [dict(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3),
                                dict(nx=1, ny=1, nz=2, order=4, lfactor=0.6, dfactor=4)])
def test_yago_descriptors_reproduce_oracle_matrix(kw):
    p = YagoProblem(Yago_freq_domain_params(**kw), system_cls=RecordingSystem)
    o = YagoCase(**kw)
    for a, b in [(p.V_Om, o.V_phi), (p.V_Gk, o.V_kap), (p.V_Ge, o.V_eta)]:
        assert a.ndofs == b.ndofs and np.array_equal(a.cell_dofs, b.cell_dofs)
    mats, vecs = p.sys.evaluate()
    A, b = o.assemble()
    A = A.tocsr(); A.sort_indices()
    M = mats[0]
    assert np.array_equal(A.indptr, M.indptr) and np.array_equal(A.indices, M.indices)
    assert np.abs(A.data - M.data).max() <= 1e-12 * np.abs(A.data).max()
    assert np.abs(b - vecs[0]).max() <= 1e-12 * np.abs(b).max()


def test_yago_basis_matrices_reproduce_oracle_over_a_sweep():
    """ω-sweep by basis matrices: A(ω) = B₀ + ω B₁ + ω² B₂ + k B₃ with the B_k integrated once
    reproduces the oracle's literal form at every λfactor of scripts/5-4-1-Yago.jl:33-71 (subset)."""
    from vlfs_b200.cases.yago import MAT_B0, MAT_B1, MAT_B2, MAT_BK
    kw = dict(nx=2, ny=2, nz=1, order=2, lfactor=0.4, dfactor=3)
    p = YagoProblem(Yago_freq_domain_params(**kw), system_cls=RecordingSystem, basis=True)
    mats, _ = p.sys.evaluate()
    for lf in (0.4, 0.6, 0.8):
        p.set_frequency(lf)
        _, vecs = p.sys.evaluate()
        A, b = YagoCase(**dict(kw, lfactor=lf)).assemble()
        A = A.tocsr(); A.sort_indices()
        M = mats[MAT_B0] + p.om * mats[MAT_B1] + p.om ** 2 * mats[MAT_B2] + p.k * mats[MAT_BK]
        for B in mats:            # one union pattern, explicit zeros kept
            assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        D = (M - A).tocsr()
        assert (abs(D).max() if D.nnz else 0.0) <= 1e-12 * np.abs(A.data).max()
        assert np.abs(b - vecs[0]).max() <= 1e-12 * np.abs(b).max()


# ---- the other configurations of BASELINE.json: product descriptors vs the oracle's literal forms
def _same_pattern_and_values(M, A, tol=1e-12):
    A = A.tocsr(); A.sort_indices()
    assert np.array_equal(A.indptr, M.indptr) and np.array_equal(A.indices, M.indices)
    scale = max(np.abs(A.data).max(), 1e-300)
    assert np.abs(A.data - M.data).max() <= tol * scale


@pytest.mark.parametrize("kw", [dict(n=3, dt=0.1, tf=0.2, orderphi=2, ordereta=2, k=1),
                                dict(n=4, dt=1e-6, tf=1e-4, orderphi=2, ordereta=4, k=15)])
def test_periodic_beam_descriptors_reproduce_oracle(kw):
    from vlfs_b200.cases.periodic_beam import PeriodicBeamProblem, Periodic_Beam_params
    from oracle.cases.periodic_beam import PeriodicBeamCase
    p = PeriodicBeamProblem(Periodic_Beam_params(**kw), system_cls=RecordingSystem)
    o = PeriodicBeamCase(**kw)
    assert np.array_equal(p.V_Om.cell_dofs, o.V_phi.cell_dofs) and np.array_equal(p.V_G.cell_dofs, o.V_eta.cell_dofs)
    mats, _ = p.sys.evaluate()
    for M, A in zip(mats[:3], o.assemble()):
        _same_pattern_and_values(M, A)
    for d in range(3):
        assert np.abs(p.interpolate(0.0, d) - o.interpolate(0.0, d)).max() <= 1e-15 * max(np.abs(o.interpolate(0.0, d)).max(), 1)


@pytest.mark.parametrize("kw", [dict(nx=10, ny=1, order=2, xi=0.0),          # scripts/5-2-1-...jl:17-28
                                dict(nx=10, ny=2, order=4, xi=625.0)])
def test_khabakhpasheva_descriptors_reproduce_oracle(kw):
    from vlfs_b200.cases.khabakhpasheva import (KhabakhpashevaFreqProblem, KhabakhpashevaTimeProblem,
                                                Khabakhpasheva_freq_domain_params, Khabakhpasheva_time_domain_params)
    from oracle.cases.khabakhpasheva import KhabakhpashevaFreqCase, KhabakhpashevaTimeCase
    p = KhabakhpashevaFreqProblem(Khabakhpasheva_freq_domain_params(**kw), system_cls=RecordingSystem)
    o = KhabakhpashevaFreqCase(**kw)
    assert o.Lj.nfaces == 1
    for a, b in [(p.V_Om, o.V_phi), (p.V_Gk, o.V_kap), (p.V_Ge, o.V_eta)]:
        assert np.array_equal(a.cell_dofs, b.cell_dofs)
    mats, vecs = p.sys.evaluate()
    A, b = o.assemble()
    _same_pattern_and_values(mats[0], A)
    assert np.abs(b - vecs[0]).max() <= 1e-12 * np.abs(b).max()
    pt_ = KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(**kw), system_cls=RecordingSystem)
    ot = KhabakhpashevaTimeCase(**kw)
    mats, vecs = pt_.sys.evaluate()
    for M, A in zip(mats[:3], ot.assemble()):
        _same_pattern_and_values(M, A)
    for t in (0.3, 1.7):
        bt = np.cos(ot.c["om"] * t) * vecs[0] + np.sin(ot.c["om"] * t) * vecs[1]
        br = ot.rhs(t)
        assert np.abs(bt - br).max() <= 1e-12 * np.abs(br).max()
    assert abs(pt_.ah - ot.time_constants()["ah"]) < 1e-15 and abs(pt_.dt - ot.time_constants()["dt"]) < 1e-18


@pytest.mark.parametrize("order", [2, 4])
def test_multigeo_descriptors_reproduce_oracle(order, golden_dir):
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    from oracle.cases.multigeo import MultiGeoCase
    f = os.path.join(golden_dir, "multi_geo_coarse.json")
    p = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=order), system_cls=RecordingSystem)
    o = MultiGeoCase(f, order=order)
    for a, b in [(p.V_Om, o.V_phi), (p.V_Gk, o.V_kap), (p.V_Ge, o.V_eta)]:
        assert np.array_equal(a.cell_dofs, b.cell_dofs)
    mats, vecs = p.sys.evaluate()
    A, b = o.assemble()
    _same_pattern_and_values(mats[0], A, tol=2e-12)
    assert np.abs(b - vecs[0]).max() <= 2e-12 * np.abs(b).max()


def test_gmsh_reader_reproduces_golden_json(golden_dir):
    """The reference's golden pair multi_geo_coarse.msh <-> .json (models/multi_geo.jl:5-7)."""
    a = G.DiscreteModelFromFile(os.path.join(golden_dir, "multi_geo_coarse.msh"))
    b = G.DiscreteModelFromFile(os.path.join(golden_dir, "multi_geo_coarse.json"))
    assert np.array_equal(a.node_coords, b.node_coords) and np.array_equal(a.cell_nodes, b.cell_nodes)
    for d in (1, 2, 3):
        assert np.array_equal(a.face_vertices[d], b.face_vertices[d])
    for d in range(4):
        assert np.array_equal(a.face_entity[d], b.face_entity[d])
    assert a.tags == b.tags


def test_liu_descriptors_reproduce_oracle(golden_dir):
    """Configuration 5-3-1 on the reference's own model file (models/Liu.json, 6 016 triangles)."""
    from vlfs_b200.cases.liu import LiuProblem, Liu_params
    from oracle.cases.liu import LiuCase
    f = os.path.join(golden_dir, "Liu.json")
    p = LiuProblem(Liu_params(omega=0.4, mesh_file=f), system_cls=RecordingSystem)
    o = LiuCase(f, omega=0.4)
    assert abs(p.c["k"] - o.c["k"]) <= 1e-15
    for a, b in [(p.V_Om, o.V_phi), (p.V_Gk, o.V_kap), (p.V_Ge, o.V_eta)]:
        assert a.ndofs == b.ndofs and np.array_equal(a.cell_dofs, b.cell_dofs)
    mats, vecs = p.sys.evaluate()
    A, b = o.assemble()
    _same_pattern_and_values(mats[0], A)
    assert np.abs(b - vecs[0]).max() <= 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("kw", [dict(n=4, dt=0.01, tf=0.02, order=2, k=1),
                                dict(n=6, dt=1e-3, tf=2e-3, order=4, k=3)])
def test_periodic_beam_fs_descriptors_reproduce_oracle(kw):
    """Configuration 5-1-4 (src/Periodic_Beam_FS.jl): beam on half the surface, three fields."""
    from vlfs_b200.cases.periodic_beam_fs import PeriodicBeamFSProblem, Periodic_Beam_FS_params
    from oracle.cases.periodic_beam_fs import PeriodicBeamFSCase
    p = PeriodicBeamFSProblem(Periodic_Beam_FS_params(**kw), system_cls=RecordingSystem)
    o = PeriodicBeamFSCase(**kw)
    assert o.Gb.ncells == kw["n"] and o.Gfs.ncells == kw["n"] and o.Lb.nfaces == kw["n"] - 1
    for a, b in [(p.V_Om, o.V_phi), (p.V_Gfs, o.V_kap), (p.V_Gb, o.V_eta)]:
        assert np.array_equal(a.cell_dofs, b.cell_dofs)
    mats, _ = p.sys.evaluate()
    for M, A in zip(mats[:3], o.assemble()):
        _same_pattern_and_values(M, A)
    for d in range(3):
        assert np.abs(p.interpolate(0.0, d) - o.interpolate(0.0, d)).max() <= 1e-15 * max(np.abs(o.interpolate(0.0, d)).max(), 1)
