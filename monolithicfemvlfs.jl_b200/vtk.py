"""`writevtk(trian, filename, cellfields=[name => f...], nsubcells=n)` of the reference's `vtk_output=true`
branches (src/Yago_freq_domain.jl:201-205, src/Khabakhpasheva_freq_domain.jl, src/Multi_geo_freq_domain.jl):
every cell is cut into an n-per-direction lattice of linear sub-cells, the finite-element functions are
evaluated at the lattice points (discontinuous across cells, as Gridap's visualisation grid is) and one
`.vtu` file (VTK XML unstructured grid, ASCII) is written.  Host-side post-processing of solutions the
GPU computed; nothing here is on the hot path."""
import itertools
import numpy as np
from .fe import _vertex_shape

VTK_LINE, VTK_TRIANGLE, VTK_QUAD, VTK_TETRA, VTK_HEXAHEDRON = 3, 5, 9, 10, 12


def _lattice(poly, n):
    """Reference lattice points and the linear sub-cells (local point ids, VTK vertex order)."""
    d = poly.dim
    if not (poly.is_simplex and d > 1):                     # segment / square / cube: tensor lattice, axis 0 fastest
        pts = np.array([idx[::-1] for idx in itertools.product(*([range(n + 1)] * d))], dtype=np.float64) / n
        pid = lambda *ix: sum(i * (n + 1) ** a for a, i in enumerate(ix))
        cells = []
        for idx in itertools.product(*([range(n)] * d)):
            i = idx[::-1]
            if d == 1:
                cells.append([pid(i[0]), pid(i[0] + 1)])
            elif d == 2:
                cells.append([pid(i[0], i[1]), pid(i[0] + 1, i[1]), pid(i[0] + 1, i[1] + 1), pid(i[0], i[1] + 1)])
            else:
                c = lambda a, b, e: pid(i[0] + a, i[1] + b, i[2] + e)
                cells.append([c(0, 0, 0), c(1, 0, 0), c(1, 1, 0), c(0, 1, 0), c(0, 0, 1), c(1, 0, 1), c(1, 1, 1), c(0, 1, 1)])
        return pts, np.array(cells), {1: VTK_LINE, 2: VTK_QUAD, 3: VTK_HEXAHEDRON}[d]
    # simplices: Freudenthal (edgewise) subdivision.  In the coordinates y_k = x_k + ... + x_d the simplex is
    # the ordered region n >= y_1 >= ... >= y_d >= 0, which Kuhn's triangulation of the lattice cubes respects:
    # keep the Kuhn simplices whose vertices are all ordered and map them back.
    ids, pts = {}, []
    for idx in itertools.product(*([range(n + 1)] * d)):
        i = idx[::-1]
        if sum(i) <= n:
            ids[i] = len(pts)
            pts.append([v / n for v in i])
    to_x = lambda y: tuple(y[k] - (y[k + 1] if k + 1 < d else 0) for k in range(d))
    ordered = lambda y: all(y[k] >= y[k + 1] for k in range(d - 1)) and y[0] <= n and y[-1] >= 0
    cells = []
    for base in itertools.product(*([range(n)] * d)):
        for perm in itertools.permutations(range(d)):
            verts, cur = [tuple(base)], list(base)
            for ax in perm:
                cur = list(cur); cur[ax] += 1
                verts.append(tuple(cur))
            if all(ordered(v) for v in verts):
                cells.append([ids[to_x(v)] for v in verts])
    return np.array(pts), np.array(cells), (VTK_TRIANGLE if d == 2 else VTK_TETRA)


def evaluate_on_lattice(space, x, ref_pts):
    """(ncells, npts) values of the FE function with DOF vector x at reference points of every cell."""
    N = space.reffe.tabulate(ref_pts)[0]                     # (npts, ndofs)
    return np.einsum("pl,el->ep", N, np.asarray(x)[space.cell_dofs])


def writevtk(trian, filename, cellfields=(), nsubcells=1):
    """cellfields: iterable of (name, space, dof_vector) - the FE function `FEFunction(space, dof_vector)`.
    Complex vectors are not split here: pass `x.real` / `x.imag` as the reference does."""
    poly = trian.poly
    ref_pts, sub, vtk_type = _lattice(poly, int(nsubcells))
    X = trian.cell_coords()                                   # (ncells, nv, D)
    D = X.shape[2]
    P = np.einsum("pv,evd->epd", _vertex_shape(poly, ref_pts), X).reshape(-1, D)
    if D < 3:
        P = np.concatenate([P, np.zeros((len(P), 3 - D))], axis=1)
    npc = len(ref_pts)
    conn = (sub[None, :, :] + (np.arange(trian.ncells) * npc)[:, None, None]).reshape(-1, sub.shape[1])
    data = [(name, evaluate_on_lattice(space, np.real(x), ref_pts).reshape(-1)) for name, space, x in cellfields]
    path = filename if filename.endswith(".vtu") else filename + ".vtu"
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1" byte_order="LittleEndian">\n')
        f.write('<UnstructuredGrid>\n<Piece NumberOfPoints="%d" NumberOfCells="%d">\n' % (len(P), len(conn)))
        f.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">\n')
        np.savetxt(f, P, fmt="%.17g")
        f.write('</DataArray>\n</Points>\n<Cells>\n<DataArray type="Int64" Name="connectivity" format="ascii">\n')
        np.savetxt(f, conn, fmt="%d")
        f.write('</DataArray>\n<DataArray type="Int64" Name="offsets" format="ascii">\n')
        np.savetxt(f, (np.arange(len(conn)) + 1) * conn.shape[1], fmt="%d")
        f.write('</DataArray>\n<DataArray type="UInt8" Name="types" format="ascii">\n')
        np.savetxt(f, np.full(len(conn), vtk_type), fmt="%d")
        f.write('</DataArray>\n</Cells>\n<PointData>\n')
        for name, v in data:
            f.write('<DataArray type="Float64" Name="%s" format="ascii">\n' % name)
            np.savetxt(f, v, fmt="%.17g")
            f.write('</DataArray>\n')
        f.write('</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
    return path
