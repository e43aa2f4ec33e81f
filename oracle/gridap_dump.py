"""Reader / writer of the table dumps of `julia/dump_gridap_tables.jl` and the comparison of an
oracle case with such a dump.  Test infrastructure only (tests/test_gridap_dump.py): this is what
turns the oracle's "parity unpinned" into pinned once somebody with a Julia install has run the
dump script on the reference (SURVEY.md §7 hard part 2d, §8c)."""
import os
import numpy as np
import scipy.sparse as sp


def write_dump(dirname, name, A_csc, b, x, spaces, dof_coords):
    """Same files the Julia script writes (1-based indices, column-major nothing: flat arrays)."""
    d = os.path.join(dirname, name)
    os.makedirs(d, exist_ok=True)
    A = sp.csc_matrix(A_csc)
    cplx = np.iscomplexobj(A.data)
    ext = "c128" if cplx else "f64"
    dt = np.complex128 if cplx else np.float64
    (A.indptr.astype(np.int64) + 1).tofile(os.path.join(d, "colptr.i64"))
    (A.indices.astype(np.int64) + 1).tofile(os.path.join(d, "rowval.i64"))
    A.data.astype(dt).tofile(os.path.join(d, "nzval." + ext))
    np.asarray(b, dtype=dt).tofile(os.path.join(d, "b." + ext))
    np.asarray(x, dtype=dt).tofile(os.path.join(d, "x." + ext))
    lines = ["case %s" % name, "ndofs %d" % A.shape[0], "nnz %d" % A.nnz, "scalar %s" % ext,
             "nfields %d" % len(spaces)]
    off = 0
    for f, V in enumerate(spaces, start=1):
        (np.asarray(V.cell_dofs, dtype=np.int64) + 1).tofile(os.path.join(d, "cell_dofs_%d.i64" % f))
        lines.append("field %d ncells %d ndofs_per_cell %d free_dofs %d offset %d"
                     % (f, V.cell_dofs.shape[0], V.cell_dofs.shape[1], V.ndofs, off))
        off += V.ndofs
    np.asarray(dof_coords, dtype=np.float64).tofile(os.path.join(d, "dof_coords.f64"))
    lines.append("dim %d" % np.asarray(dof_coords).shape[1])
    with open(os.path.join(d, "manifest.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


def read_dump(d):
    man = {}
    fields = []
    for line in open(os.path.join(d, "manifest.txt")):
        t = line.split()
        if not t:
            continue
        if t[0] == "field":
            fields.append(dict(ncells=int(t[3]), nd=int(t[5]), free=int(t[7]), offset=int(t[9])))
        else:
            man[t[0]] = t[1]
    n, nnz, ext = int(man["ndofs"]), int(man["nnz"]), man["scalar"]
    dt = np.complex128 if ext == "c128" else np.float64
    colptr = np.fromfile(os.path.join(d, "colptr.i64"), dtype=np.int64) - 1
    rowval = np.fromfile(os.path.join(d, "rowval.i64"), dtype=np.int64) - 1
    nz = np.fromfile(os.path.join(d, "nzval." + ext), dtype=dt)
    assert len(colptr) == n + 1 and len(rowval) == nnz and len(nz) == nnz
    out = dict(n=n, A=sp.csc_matrix((nz, rowval, colptr), shape=(n, n)),
               b=np.fromfile(os.path.join(d, "b." + ext), dtype=dt),
               x=np.fromfile(os.path.join(d, "x." + ext), dtype=dt), fields=[],
               dim=int(man["dim"]))
    for f, info in enumerate(fields, start=1):
        tab = np.fromfile(os.path.join(d, "cell_dofs_%d.i64" % f), dtype=np.int64).reshape(info["ncells"], info["nd"]) - 1
        out["fields"].append(dict(info, cell_dofs=tab))
    out["dof_coords"] = np.fromfile(os.path.join(d, "dof_coords.f64"), dtype=np.float64).reshape(n, out["dim"])
    return out


def compare(case, dump, entry_tol=1e-12, sol_tol=1e-8):
    """Oracle case (has .X.spaces, .assemble()) against a dump.  Returns a dict of findings; raises
    AssertionError with the first mismatch."""
    A, b = case.assemble()
    A = sp.csc_matrix(A)
    A.sort_indices()
    G = dump["A"].copy()
    G.sort_indices()
    spaces = case.X.spaces
    assert dump["n"] == A.shape[0], "number of DOFs: oracle %d, Gridap %d" % (A.shape[0], dump["n"])
    assert len(dump["fields"]) == len(spaces)
    for f, (V, fd) in enumerate(zip(spaces, dump["fields"])):
        assert fd["free"] == V.ndofs, "field %d: free DOFs differ" % f
        assert fd["cell_dofs"].shape == V.cell_dofs.shape, "field %d: cell->DOF table shape" % f
        assert np.array_equal(fd["cell_dofs"], V.cell_dofs), "field %d: cell->DOF numbering differs from Gridap" % f
    assert np.array_equal(G.indptr, A.indptr) and np.array_equal(G.indices, A.indices), "sparsity pattern differs"
    scale = np.abs(G.data).max()
    err = np.abs(G.data - A.data).max() / scale
    assert err <= entry_tol, "matrix entries differ: %.3e" % err
    berr = np.abs(dump["b"] - b).max() / max(np.abs(dump["b"]).max(), 1e-300)
    assert berr <= entry_tol, "right-hand side differs: %.3e" % berr
    import scipy.sparse.linalg as spla
    x = spla.splu(A).solve(b)
    serr = np.linalg.norm(x - dump["x"]) / np.linalg.norm(dump["x"])
    assert serr <= sol_tol, "solution differs: %.3e" % serr
    return dict(entry_err=err, rhs_err=berr, solution_err=serr, nnz=A.nnz)
