"""Uniform refinement of simplicial models (`refine_uniform`, the synthetic M2 meshes of SURVEY.md
§8d): geometric and topological invariants, inherited labels, and the monolithic MultiGeo / Liu
operators assembled on the refined meshes (compiled CPU restatement) - CPU only."""
import os
import numpy as np
import pytest
from vlfs_b200.geometry import DiscreteModelFromFile, refine_uniform, Boundary, Skeleton

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _measure(X):
    d = X.shape[1] - 1
    E = X[:, 1:] - X[:, :1]
    G = np.einsum("eab,ecb->eac", E, E)
    return np.sqrt(np.abs(np.linalg.det(G))) / [1, 1, 2, 6][d]


@pytest.mark.parametrize("name,tags", [("multi_geo_coarse.json", ["inlet", "plate", "freeSurface"]),
                                       ("Liu.json", ["inlet", "beam", "free_surface", "damping_in", "damping_out"])])
def test_refinement_invariants(name, tags):
    m = DiscreteModelFromFile(os.path.join(GOLD, name))
    f = refine_uniform(m)
    D = m.D
    assert f.ncells == m.ncells * 2 ** D and f.nvertices == m.nvertices + m.nfaces(1)
    vm, vf = _measure(m.node_coords[m.cell_nodes]), _measure(f.node_coords[f.cell_nodes])
    assert abs(vm.sum() - vf.sum()) <= 1e-12 * vm.sum() and vf.min() > 0
    # children of a simplex have equal measure
    assert np.abs(vf.reshape(m.ncells, -1) - (vm / 2 ** D)[:, None]).max() <= 1e-12 * vm.max()
    for tag in tags:
        a, b = Boundary(m, tags=tag), Boundary(f, tags=tag)
        assert b.ncells == a.ncells * 2 ** (D - 1)
        assert abs(_measure(a.cell_coords()).sum() - _measure(b.cell_coords()).sum()) <= 1e-12 * _measure(a.cell_coords()).sum()
    _, _, count = f.facet_incidence()
    assert count.max() == 2 and (count == 1).sum() == 2 ** (D - 1) * (m.facet_incidence()[2] == 1).sum()
    euler = lambda mm: sum((-1) ** d * mm.nfaces(d) for d in range(D)) + (-1) ** D * mm.ncells
    assert euler(f) == euler(m)
    assert np.all(np.diff(f.cell_vertices, axis=1) > 0)          # simplex vertex ids ascending
    # interior faces keep the entity of the parent cell, boundary labels only live on the boundary
    interior_entity = set(np.unique(m.face_entity[D]))
    facets_in = np.nonzero(count == 2)[0]
    assert set(np.unique(f.face_entity[D - 1][facets_in])) <= interior_entity | set(np.unique(m.face_entity[D - 1][m.facet_incidence()[2] == 2]))


def test_multigeo_operator_on_a_refined_mesh():
    from oracle.c_assembly import CSystem
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    f = os.path.join(GOLD, "multi_geo_coarse.json")
    p0 = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=2), system_cls=CSystem)
    p1 = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=2, refine=1), system_cls=CSystem)
    assert p1.Om.ncells == 8 * p0.Om.ncells and p1.Gb.ncells == 4 * p0.Gb.ncells and p1.Lb.ncells == 145
    mats, vecs = p1.sys.assemble()
    A = mats[0]
    # fluid DOFs away from every surface: pure Laplace rows, constants in the kernel
    touched = np.zeros(p1.X.ndofs, bool)
    for tri in (p1.Gf, p1.Gb, p1.Gin):
        touched[np.unique(p1.V_Om.cell_dofs[tri.cell])] = True
    o = p1.X.offsets
    inner = np.nonzero(~touched[o[0]:o[1]])[0]
    assert len(inner) > 100
    rows = A[inner]
    assert np.abs(np.asarray(rows.sum(axis=1)).ravel()).max() <= 1e-12 * np.abs(rows).sum(axis=1).max()
    # the plate block (v,η) is Hermitian up to the complex scalar: its bending + mass part is real symmetric
    eta = slice(o[2], o[3])
    B = A[eta][:, eta].toarray()
    assert np.abs(B - B.T).max() <= 1e-10 * np.abs(B).max()
