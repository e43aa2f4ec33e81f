"""`writevtk` of the host mirror (the reference's `vtk_output=true` branches): lattice geometry, sub-cell
connectivity and FE function evaluation, checked on functions the spaces represent exactly."""
import re
import numpy as np
import pytest
import vlfs_b200
from vlfs_b200.vtk import writevtk, _lattice


def _read(path):
    txt = open(path).read()
    arr = lambda name: np.array(re.search(r'Name="%s"[^>]*>\n(.*?)</DataArray>' % name, txt, re.S).group(1).split(), dtype=float)
    pts = np.array(re.search(r'<Points>\n<DataArray[^>]*>\n(.*?)</DataArray>', txt, re.S).group(1).split(), dtype=float).reshape(-1, 3)
    return pts, arr


@pytest.mark.parametrize("dims,order", [((3,), 2), ((3, 2), 3), ((2, 2, 2), 2)])
def test_cartesian_lattice_reproduces_polynomials(tmp_path, dims, order):
    dom = sum(([0.0, 1.0 + a] for a in range(len(dims))), [])
    model = vlfs_b200.CartesianDiscreteModel(tuple(dom), dims)
    trian = vlfs_b200.Interior(model) if hasattr(vlfs_b200, "Interior") else vlfs_b200.Triangulation(model)
    V = vlfs_b200.TestFESpace(trian, vlfs_b200.ReferenceFE("lagrangian", float, order))
    xd = V.cell_dof_coords()
    f = lambda x: 1.0 + x[..., 0] ** order + (0.5 * x[..., -1] if x.shape[-1] > 1 else 0.0)
    dofs = np.zeros(V.ndofs)
    dofs[V.cell_dofs] = f(xd)
    path = writevtk(trian, str(tmp_path / "t"), [("f", V, dofs)], nsubcells=3)
    pts, arr = _read(path)
    D = len(dims)
    assert len(pts) == trian.ncells * 4 ** D
    assert np.abs(arr("f") - f(pts[:, :D])).max() < 1e-12
    ncell = int(np.prod(dims)) * 3 ** D
    assert len(arr("offsets")) == ncell and len(arr("types")) == ncell


@pytest.mark.parametrize("name", ["TRI", "TET"])
def test_simplex_sub_cells_tile_the_reference_cell(name):
    from vlfs_b200.geometry import Polytope
    poly = Polytope(name)
    n = 4
    pts, cells, _ = _lattice(poly, n)
    d = poly.dim
    vol = 0.0
    for c in cells:
        M = (pts[c[1:]] - pts[c[0]])
        vol += abs(np.linalg.det(M)) / (2 if d == 2 else 6)
    assert abs(vol - (0.5 if d == 2 else 1.0 / 6.0)) < 1e-12          # no gaps, no overlaps
    assert len(cells) == n ** d


def test_boundary_triangulation_of_a_tet_mesh(tmp_path, golden_dir):
    """Surface triangulations of the gmsh model (the `_Gk` / `_Ge` files of 5-5-1): points lie on the
    boundary facets and a P2 field is reproduced on the lattice."""
    import os
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    from oracle.descriptor_eval import RecordingSystem
    f = os.path.join(golden_dir, "multi_geo_coarse.json")
    prob = MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=f, order=2), system_cls=RecordingSystem)
    V = prob.V_Ge
    xd = V.cell_dof_coords()
    g = lambda x: x[..., 0] ** 2 - 0.3 * x[..., 0] * x[..., 1] + 2.0
    dofs = np.zeros(V.ndofs)
    dofs[V.cell_dofs] = g(xd)
    path = writevtk(prob.Gb, str(tmp_path / "ge"), [("eta_re", V, dofs)], nsubcells=4)
    pts, arr = _read(path)
    assert len(pts) == prob.Gb.ncells * 15 and np.abs(arr("eta_re") - g(pts)).max() < 1e-12
