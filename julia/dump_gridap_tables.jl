# dump_gridap_tables.jl - the pinning tool (SURVEY.md §7 hard part 2d, §8c).
#
# STATUS: UNTESTED here (no Julia in the build container / on the GPU box).  Run it once on any
# machine with Julia and the reference checked out:
#
#     julia --project=/path/to/MonolithicFEMVLFS.jl julia/dump_gridap_tables.jl out_dir
#
# It runs the five warm-up configurations of the reference through UNMODIFIED Gridap and writes,
# per case, a directory `out_dir/<case>/` of flat little-endian binaries + `manifest.txt`:
#
#   cell_dofs_<field>.i64   cell -> DOF table of every field on its own triangulation (1-based, Gridap order)
#   dof_coords.f64          [ndofs][D] DOF node coordinates in Gridap DOF order
#   colptr.i64, rowval.i64  pattern of get_matrix(op)::SparseMatrixCSC (1-based)
#   nzval.f64 / nzval.c128  its values (explicit zeros kept), b.f64 / b.c128 = get_vector(op)
#   x.f64 / x.c128          solve(op) with LUSolver() (UMFPACK)
#
# Copy `out_dir` to `tests/golden/gridap_dump/` of this repository: `tests/test_gridap_dump.py` then
# compares the oracle's DOF numbering, sparsity pattern (bit-exact), entries (1e-12) and solution
# (1e-8) with what Gridap itself produced, which turns the oracle's "parity unpinned" into pinned.
using Gridap
using SparseArrays
using LinearAlgebra

const REF = get(ENV, "VLFS_REFERENCE", joinpath(@__DIR__, "..", "..", "MonolithicFEMVLFS.jl"))

function write_array(path, a)
  open(path, "w") do io
    write(io, a)
  end
end

function dump_operator(dir, name, op, X, spaces::Vector, trians::Vector)
  d = joinpath(dir, name)
  mkpath(d)
  A = get_matrix(op)
  b = get_vector(op)
  x = A \ b                                   # UMFPACK, what LUSolver() calls
  T = eltype(A)
  ext = T <: Complex ? "c128" : "f64"
  write_array(joinpath(d, "colptr.i64"), Int64.(A.colptr))
  write_array(joinpath(d, "rowval.i64"), Int64.(A.rowval))
  write_array(joinpath(d, "nzval.$ext"), A.nzval)
  write_array(joinpath(d, "b.$ext"), Vector{T}(b))
  write_array(joinpath(d, "x.$ext"), Vector{T}(x))
  lines = String["case $name", "ndofs $(size(A, 1))", "nnz $(nnz(A))", "scalar $ext", "nfields $(length(spaces))"]
  off = 0
  for (f, (V, trian)) in enumerate(zip(spaces, trians))
    ids = Gridap.FESpaces.get_cell_dof_ids(V)
    nc = length(ids); nd = length(ids[1])
    tab = Array{Int64}(undef, nd, nc)          # column = cell  ->  row-major [ncells][ndofs] on disk
    for (e, row) in enumerate(ids), i in 1:nd
      tab[i, e] = row[i]
    end
    write_array(joinpath(d, "cell_dofs_$f.i64"), tab)
    push!(lines, "field $f ncells $nc ndofs_per_cell $nd free_dofs $(num_free_dofs(V)) offset $off")
    off += num_free_dofs(V)
  end
  # DOF coordinates through the cell-wise DOF nodes
  D = num_point_dims(trians[1])
  xyz = zeros(Float64, D, size(A, 1))
  off = 0
  for V in spaces
    nodes = Gridap.CellData.get_array(Gridap.CellData.get_cell_points(Gridap.FESpaces.get_fe_dof_basis(V)))
    ids = Gridap.FESpaces.get_cell_dof_ids(V)
    for (e, xs) in enumerate(nodes), (i, p) in enumerate(xs)
      xyz[:, off + ids[e][i]] .= Tuple(p)
    end
    off += num_free_dofs(V)
  end
  write_array(joinpath(d, "dof_coords.f64"), xyz)
  push!(lines, "dim $D")
  open(joinpath(d, "manifest.txt"), "w") do io
    foreach(l -> println(io, l), lines)
  end
  println("wrote $d: N=$(size(A,1)) nnz=$(nnz(A))")
end

# --- the warm-up cases: the reference's own modules with their warm-up parameters ----------------
# Each src/*.jl keeps its operator in local variables; the smallest change that exposes them is a
# `return (op, X, (V_Ω, V_Γκ, V_Γη), (Ω, Γκ, Γη))` appended after the `AffineFEOperator` line
# (src/Yago_freq_domain.jl:167, src/Khabakhpasheva_freq_domain.jl:240, src/Multi_geo_freq_domain.jl:131)
# behind an optional keyword `dump=true`.  With that patch applied:
function main(out)
  include(joinpath(REF, "src", "MonolithicFEMVLFS.jl"))
  M = Main.MonolithicFEMVLFS
  cases = [
    ("yago_warmup", () -> M.Yago_freq_domain.run_Yago_freq_domain(
        M.Yago_freq_domain.Yago_freq_domain_params(nx=2, ny=2, nz=1, order=2, λfactor=0.4, dfactor=3); dump=true)),
    ("khabakhpasheva_freq_warmup", () -> M.Khabakhpasheva_freq_domain.run_Khabakhpasheva_freq_domain(
        M.Khabakhpasheva_freq_domain.Khabakhpasheva_freq_domain_params(nx=2, ny=1, order=2, ξ=0.0); dump=true)),
    ("multigeo_coarse", () -> M.MultiGeo_freq_domain.run_MultiGeo_freq_domain(
        M.MultiGeo_freq_domain.MultiGeo_freq_domain_params(mesh_file="multi_geo_coarse.json", order=2); dump=true)),
    ("liu_warmup", () -> M.Liu.run_Liu(M.Liu.Liu_params(ω=0.4, order=2); dump=true)),
  ]
  for (name, f) in cases
    try
      op, X, spaces, trians = f()
      dump_operator(out, name, op, X, collect(spaces), collect(trians))
    catch err
      @warn "case $name failed (is the `dump=true` return patch applied?)" err
    end
  end
end

main(length(ARGS) >= 1 ? ARGS[1] : "gridap_dump")
