# VLFSB200Gridap.jl - the Gridap-side half of the drop-in boundary (SURVEY.md §8b).
#
# STATUS: UNTESTED.  No Julia toolchain exists in the build container or on the GPU box
# (SURVEY.md §0.3); nothing in this file has ever been executed.  It is written against
# Gridap 0.17 (Project.toml:24 of the reference) and is the twin of the Python host mirror
# (monolithicfemvlfs.jl_b200/fe.py `Measure.*_slot`, assembler.py), which IS what the parity tests
# run; the C entry points it binds are exercised from C++ exactly the way `ccall` would call them
# (tests/cabi_harness.cpp).  A maintainer with Julia should start with `julia/dump_gridap_tables.jl`,
# which also pins the oracle's ◇ assumptions about Gridap's numbering.
#
# Three levels of integration, smallest change first:
#
#  (1) `B200LUSolver()` - a `Gridap.Algebra.LinearSolver`.  One-token change at
#      src/Periodic_Beam.jl:107 and src/Khabakhpasheva_time_domain.jl:258
#      (`ls = LUSolver()` -> `ls = B200LUSolver()`), and `solve(LinearFESolver(B200LUSolver()), op)`
#      at src/Yago_freq_domain.jl:171 / src/Multi_geo_freq_domain.jl:135 /
#      src/Khabakhpasheva_freq_domain.jl:241.  Gridap still assembles on the CPU; the matrix it hands
#      to `numerical_setup` is imported with `vlfs_import_csr` and factorised on the GPU.
#  (2) `B200Assembler(X, Y; terms...)` + `AffineFEOperator(a, l, X, Y, assem)` is NOT possible
#      without changing what Gridap integrates (at `assemble_matrix(assem, data)` the cell matrices
#      already exist on the CPU, SURVEY.md §8b), therefore:
#  (3) `b200_affine_operator(terms, X, Y)` - integration, assembly AND solve on the GPU: the weak form
#      is given as a list of `B200Term`s (one per product of the reference's `a`, `l`), the tables
#      come out of Gridap through `own_slot` / `trace_slot` / `surface_slot` / `skeleton_slot` below,
#      and the result is an ordinary `AffineFEOperator(trial, test, A, b)`.
#
# `MultiGPUSystem(ndofs; ngpu=8)` requests all GPUs of the box through `vlfs_multi_create`; the
# collective calls (`vlfs_multi_*`) run one host thread per GPU inside the library.
module VLFSB200Gridap

using SparseArrays
using LinearAlgebra
using Gridap
using Gridap.Algebra
using Gridap.Arrays
using Gridap.CellData
using Gridap.FESpaces
using Gridap.Fields
using Gridap.Geometry
using Gridap.ReferenceFEs
using Gridap.TensorValues

include("VLFSB200.jl")
using .VLFSB200
const LIB = VLFSB200.LIB

# ---------------------------------------------------------------------------------------------
# (1) LinearSolver on a matrix Gridap assembled
# ---------------------------------------------------------------------------------------------
struct B200LUSolver <: Gridap.Algebra.LinearSolver
  device::Int
  rtol::Float64
  coords::Union{Nothing,Matrix{Float64}}     # optional DOF coordinates [dim, ndofs] for the nested dissection
end
B200LUSolver(; device=0, rtol=1e-12, coords=nothing) = B200LUSolver(device, rtol, coords)

struct B200SymbolicSetup <: Gridap.Algebra.SymbolicSetup
  solver::B200LUSolver
end

mutable struct B200NumericalSetup <: Gridap.Algebra.NumericalSetup
  ctx::Ptr{Cvoid}
  perm::Vector{Int}          # CSR slot -> position in the CSC `nzval`
  n::Int
  iscomplex::Bool
end

function _check(ctx, rc; allow=(0,))
  rc in allow && return rc
  msg = unsafe_string(ccall((:vlfs_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
  error("libvlfs_b200 status $rc: $msg")
end

# CSR view (0-based Int32) of a SparseMatrixCSC and the permutation of its values
function csr_of(A::SparseMatrixCSC)
  P = SparseMatrixCSC(size(A, 1), size(A, 2), A.colptr, A.rowval, collect(1:nnz(A)))
  Pt = sparse(transpose(P))                   # columns of Pt = rows of A, row indices ascending
  Int32.(Pt.colptr .- 1), Int32.(Pt.rowval .- 1), Pt.nzval
end

Gridap.Algebra.symbolic_setup(ls::B200LUSolver, A::AbstractMatrix) = B200SymbolicSetup(ls)

function Gridap.Algebra.numerical_setup(ss::B200SymbolicSetup, A::SparseMatrixCSC{T}) where {T}
  ls = ss.solver
  ref = Ref{Ptr{Cvoid}}(C_NULL)
  rc = ccall((:vlfs_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, ls.device)
  rc == 0 || error("vlfs_create failed (status $rc): no usable sm_100 GPU; there is no CPU fallback")
  ctx = ref[]
  rowptr, colind, perm = csr_of(A)
  cplx = T <: Complex
  GC.@preserve rowptr colind begin
    _check(ctx, ccall((:vlfs_import_csr, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Ptr{Int32}, Ptr{Int32}),
                      ctx, size(A, 1), cplx, 1, rowptr, colind))
  end
  ns = B200NumericalSetup(ctx, perm, size(A, 1), cplx)
  finalizer(x -> ccall((:vlfs_destroy, LIB), Cint, (Ptr{Cvoid},), x.ctx), ns)
  if ls.coords !== nothing
    xyz = Matrix{Float64}(ls.coords)          # [dim, ndofs] column-major = row-major [ndofs][dim]
    GC.@preserve xyz _check(ctx, ccall((:vlfs_set_dof_coords, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), ctx, size(xyz, 1), xyz))
  end
  _upload_values!(ns, A)
  o = VLFSB200.SolverOpts(VLFSB200.SOLVER_GMRES, VLFSB200.PRECOND_SPARSE_LU, 20, 40, ls.rtol, 0, 0)
  # 4 = VLFS_PIVOT_PERTURBED: usable factorisation, refinement decides
  _check(ctx, ccall((:vlfs_solver_setup, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{VLFSB200.SolverOpts}), ctx, 0, o); allow=(0, 4))
  ns
end

function _upload_values!(ns::B200NumericalSetup, A::SparseMatrixCSC{T}) where {T}
  vals = ns.iscomplex ? ComplexF64.(A.nzval[ns.perm]) : Float64.(A.nzval[ns.perm])
  GC.@preserve vals _check(ns.ctx, ccall((:vlfs_set_values, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), ns.ctx, 0, vals))
end

function Gridap.Algebra.numerical_setup!(ns::B200NumericalSetup, A::SparseMatrixCSC)
  _upload_values!(ns, A)                       # same pattern: keep the symbolic analysis
  _check(ns.ctx, ccall((:vlfs_solver_refactor, LIB), Cint, (Ptr{Cvoid}, Cint), ns.ctx, 0); allow=(0, 4))
  ns
end

function Gridap.Algebra.solve!(x::AbstractVector, ns::B200NumericalSetup, b::AbstractVector)
  T = ns.iscomplex ? ComplexF64 : Float64
  bb = Vector{T}(b)
  xx = zeros(T, ns.n)
  st = VLFSB200.SolveStats()
  GC.@preserve bb xx begin
    rc = ccall((:vlfs_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{VLFSB200.SolveStats}), ns.ctx, bb, xx, st)
  end
  _check(ns.ctx, rc; allow=(0, 3))              # 3 = VLFS_NOT_CONVERGED: statistics, not an exception
  st.converged != 0 || @warn "B200LUSolver: refinement stopped at relative residual $(st.rel_residual), backward error $(st.backward_error)"
  copyto!(x, xx)
  x
end

# ---------------------------------------------------------------------------------------------
# (3) tables out of Gridap  (twins of fe.py: Measure.own_slot / trace_slot / surface_slot / skeleton_slot)
# ---------------------------------------------------------------------------------------------
# components of a Gridap tensor value as plain numbers
_comps(v::Number) = (v,)
_comps(v::MultiValue) = Tuple(v.data)

# [nvar][nq][ndofs] values, [..][dref] gradients, [..][dref][dref] Hessians of the scalar Lagrangian
# shape functions of `reffe` at reference points `pts`; arrays are returned in the library's
# row-major layouts (Julia arrays with reversed dimension order).
function tabulate(reffe, pts::Vector{<:Point}, need_hess::Bool)
  sf = get_shapefuns(reffe)
  nq, nd, dref = length(pts), num_dofs(reffe), length(pts[1])
  V = evaluate(sf, pts)                                  # nq x nd
  G = evaluate(Broadcasting(∇)(sf), pts)                 # nq x nd of VectorValue
  val = Array{Float64}(undef, nd, nq)
  grad = Array{Float64}(undef, dref, nd, nq)
  for q in 1:nq, i in 1:nd
    val[i, q] = V[q, i]
    for a in 1:dref
      grad[a, i, q] = G[q, i][a]
    end
  end
  hess = nothing
  if need_hess
    H = evaluate(Broadcasting(∇∇)(sf), pts)
    hess = Array{Float64}(undef, dref, dref, nd, nq)
    for q in 1:nq, i in 1:nd, a in 1:dref, b in 1:dref
      hess[b, a, i, q] = H[q, i][a, b]
    end
  end
  val, grad, hess
end

# reference gradients of the geometry map's vertex shape functions: [nq][nv][dref]
function geo_gradients(poly, pts::Vector{<:Point})
  r1 = LagrangianRefFE(Float64, poly, 1)
  G = evaluate(Broadcasting(∇)(get_shapefuns(r1)), pts)
  nq, nv, dref = length(pts), num_dofs(r1), length(pts[1])
  out = Array{Float64}(undef, dref, nv, nq)
  for q in 1:nq, v in 1:nv, a in 1:dref
    out[a, v, q] = G[q, v][a]
  end
  out
end

# cell coordinates [ncells][nv][D] (row-major) of a triangulation
function coords_table(trian)
  cc = get_cell_coordinates(trian)
  nc = length(cc); nv = length(cc[1]); D = length(cc[1][1])
  out = Array{Float64}(undef, D, nv, nc)
  for (e, xs) in enumerate(cc), v in 1:nv, d in 1:D
    out[d, v, e] = xs[v][d]
  end
  out
end

# cell -> DOF table of a single-field space on `trian`, 0-based, multi-field offset applied
function dof_table(V, trian, offset::Integer)
  ids = get_cell_dof_ids(V, trian)
  nc = length(ids); nd = length(ids[1])
  out = Array{Int32}(undef, nd, nc)
  for (e, row) in enumerate(ids), i in 1:nd
    row[i] > 0 || error("Dirichlet DOFs are not supported (the reference has none)")
    out[i, e] = Int32(row[i] - 1 + offset)
  end
  out
end

_quad_points(dΩ) = collect(dΩ.quad.cell_point[1])          # reference points of the first cell (one cell type)
_quad_weights(dΩ) = collect(Float64, dΩ.quad.cell_weight[1])
_poly(trian) = get_polytope(first(get_reffes(trian)))

"A field integrated on its own cells (`dΩ`, or a surface field on `dΓ`)."
function own_slot(V, dΩ, offset::Integer, order::Integer; need_hess=false, trian=get_triangulation(dΩ.quad))
  poly = _poly(trian)
  pts = _quad_points(dΩ)
  reffe = LagrangianRefFE(Float64, poly, order)
  val, grad, hess = tabulate(reffe, pts, need_hess)
  (dref=Int32(num_dims(poly)), nv=Int32(num_vertices(poly)), nvar=Int32(1), ndofs=Int32(num_dofs(reffe)),
   coords=coords_table(trian), variant=nothing, geo_grad=geo_gradients(poly, pts), val=val, grad=grad, hess=hess,
   dofs=dof_table(V, trian, offset), ref_normal=nothing, ref_tangent=nothing)
end
surface_slot(V, dΓ, offset, order; need_hess=false) = own_slot(V, dΓ, offset, order; need_hess=need_hess)

"A volume field seen from a boundary triangulation through the facet -> cell glue (`∇ₙ(ϕ)` on Γf)."
function trace_slot(V_Ω, dΓ, offset::Integer, order::Integer)
  Γ = get_triangulation(dΓ.quad)
  Ω = get_triangulation(V_Ω)
  D = num_cell_dims(Ω)
  glue = get_glue(Γ, Val(D))
  cellpoly = _poly(Ω)
  pts = _quad_points(dΓ)
  # reference points of every facet inside its cell (Gridap's own facet -> cell reference map, which
  # carries the facet's local id and its vertex permutation)
  refmaps = glue.tface_to_mface_map
  facet_pts = [evaluate(refmaps[f], pts) for f in 1:num_cells(Γ)]
  # table variants = distinct point sets (one per local face on oriented meshes)
  keyof(p) = Tuple(round.(collect(Iterators.flatten(_comps.(p))); digits=12))
  variants = Dict{Any,Int}(); reps = Int[]; variant = Vector{Int32}(undef, num_cells(Γ))
  for (f, p) in enumerate(facet_pts)
    k = keyof(p)
    if !haskey(variants, k)
      variants[k] = length(reps); push!(reps, f)
    end
    variant[f] = variants[k]
  end
  reffe = LagrangianRefFE(Float64, cellpoly, order)
  tabs = [tabulate(reffe, facet_pts[f], false) for f in reps]
  ggs = [geo_gradients(cellpoly, facet_pts[f]) for f in reps]
  cat_last(xs) = cat(xs...; dims=ndims(xs[1]) + 1)
  cells = glue.tface_to_mface
  cc = get_cell_coordinates(Ω)
  nv = num_vertices(cellpoly)
  coords = Array{Float64}(undef, D, nv, num_cells(Γ))
  for (f, c) in enumerate(cells), v in 1:nv, d in 1:D
    coords[d, v, f] = cc[c][v][d]
  end
  ids = get_cell_dof_ids(V_Ω)
  nd = num_dofs(reffe)
  dofs = Array{Int32}(undef, nd, num_cells(Γ))
  for (f, c) in enumerate(cells), i in 1:nd
    dofs[i, f] = Int32(ids[c][i] - 1 + offset)
  end
  (dref=Int32(D), nv=Int32(nv), nvar=Int32(length(reps)), ndofs=Int32(nd), coords=coords, variant=variant,
   geo_grad=cat_last(ggs), val=cat_last([t[1] for t in tabs]), grad=cat_last([t[2] for t in tabs]), hess=nothing,
   dofs=dofs, ref_normal=nothing, ref_tangent=nothing)
end

"One side (`:plus` / `:minus`) of a skeleton of a surface triangulation: jumps and means of ∇η, ∇∇η on Λb."
function skeleton_slot(V_Γ, dΛ, side::Symbol, offset::Integer, order::Integer)
  Λ = get_triangulation(dΛ.quad)
  S = side == :plus ? Λ.plus : Λ.minus
  Γ = get_triangulation(V_Γ)
  d = num_cell_dims(Γ)
  s = trace_slot(V_Γ, (quad=(cell_point=dΛ.quad.cell_point, cell_weight=dΛ.quad.cell_weight, trian=S),), offset, order)
  # second derivatives and the reference normals / tangents of the surface cell's facets
  poly = _poly(Γ)
  glue = get_glue(S, Val(d))
  pts = _quad_points(dΛ)
  reffe = LagrangianRefFE(Float64, poly, order)
  reps = [findfirst(==(v), s.variant) for v in 0:(s.nvar - 1)]
  hess = cat([tabulate(reffe, evaluate(glue.tface_to_mface_map[f], pts), true)[3] for f in reps]...; dims=5)
  nrm = get_facet_normal(poly)                        # outward reference normals per local facet
  lface = [glue.tface_to_mface_map[f] for f in reps]  # (kept for the maintainer: variant -> local facet)
  ref_normal = Array{Float64}(undef, d, s.nvar)
  for (k, f) in enumerate(reps)
    lf = S.glue.face_to_lface[f]
    for a in 1:d
      ref_normal[a, k] = nrm[lf][a]
    end
  end
  merge(s, (hess=hess, ref_normal=ref_normal, ref_tangent=nothing))
end

# ---------------------------------------------------------------------------------------------
# weak form as a list of products (what src/Yago_freq_domain.jl:162-166 writes with ∫( )dΩ)
# ---------------------------------------------------------------------------------------------
struct B200Term
  domain::Int                # index into the registered domains
  coef::ComplexF64
  test_op::Int32; test_slots
  trial_op::Int32; trial_slots
  weight::Int
  test_dir::NTuple{3,Float64}; trial_dir::NTuple{3,Float64}
end
B200Term(domain, coef, top, ts, uop, us; weight=-1, test_dir=(0.0, 0.0, 0.0), trial_dir=(0.0, 0.0, 0.0)) =
  B200Term(domain, ComplexF64(coef), top, ts, uop, us, weight, test_dir, trial_dir)

"""
    b200_affine_operator(sys, X, Y) -> AffineFEOperator

After the domains / terms were registered on `sys` (VLFSB200.add_domain!, add_term!, add_rhs!) and
`build_pattern!` / `assemble!` ran: Gridap's standard operator around the GPU-assembled matrix and
vector (inner constructor `AffineFEOperator(trial,test,A,b)`), so `solve`, `get_matrix`,
`get_vector`, `FEFunction` downstream (src/Yago_freq_domain.jl:171-199) are unchanged.
"""
function b200_affine_operator(sys::VLFSB200.System, X, Y)
  A = VLFSB200.get_matrix(sys, 0)
  b = VLFSB200.get_vector(sys, 0)
  AffineFEOperator(X, Y, A, b)
end

"DOF node coordinates of a multi-field space in Gridap's DOF order, row-major [ndofs][D] (for `vlfs_set_dof_coords`)."
function dof_coordinates(X::MultiFieldFESpace)
  spaces = [X[i] for i in 1:num_fields(X)]
  D = num_point_dims(get_triangulation(spaces[1]))
  out = zeros(Float64, D, num_free_dofs(X))
  off = 0
  for V in spaces
    nodes = get_cell_points(get_fe_dof_basis(V))        # physical DOF nodes, cell by cell
    ids = get_cell_dof_ids(V)
    for (e, xs) in enumerate(get_array(nodes)), (i, x) in enumerate(xs)
      g = ids[e][i]
      g > 0 && (out[:, off + g] .= Tuple(x))
    end
    off += num_free_dofs(V)
  end
  out
end

# ---------------------------------------------------------------------------------------------
# all GPUs of the box from one Julia process (vlfs_multi_create: one host thread per GPU inside
# the library, NCCL communicator owned by the library)
# ---------------------------------------------------------------------------------------------
mutable struct MultiGPUSystem
  ctxs::Vector{Ptr{Cvoid}}
  ngpu::Int
  function MultiGPUSystem(; ngpu::Integer, devices::Vector{<:Integer}=collect(0:ngpu-1))
    ctxs = fill(Ptr{Cvoid}(C_NULL), ngpu)
    devs = Cint.(devices)
    rc = ccall((:vlfs_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Cint}), ctxs, ngpu, devs)
    rc == 0 || error("vlfs_multi_create failed (status $rc): needs $ngpu sm_100 GPUs and NCCL")
    m = new(ctxs, ngpu)
    finalizer(x -> foreach(c -> ccall((:vlfs_destroy, LIB), Cint, (Ptr{Cvoid},), c), x.ctxs), m)
    m
  end
end
# Each rank's context takes the cells that touch its owned DOFs with LOCAL ids [owned | ghosts]
# (partition + halo plan: monolithicfemvlfs.jl_b200/dist.py `cell_partition_owner`,
# `local_halo_arrays` - 120 lines of index bookkeeping to transliterate), then:
multi_build_pattern!(m::MultiGPUSystem) = ccall((:vlfs_multi_build_pattern, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Ptr{Int64}, Ptr{Int64}), m.ctxs, m.ngpu, C_NULL, C_NULL)
multi_assemble!(m::MultiGPUSystem) = ccall((:vlfs_multi_assemble, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), m.ctxs, m.ngpu)
function multi_solver_setup!(m::MultiGPUSystem; rtol=1e-10)
  o = VLFSB200.SolverOpts(VLFSB200.SOLVER_GMRES, VLFSB200.PRECOND_SPARSE_LU, 20, 40, rtol, 0, 0)
  ccall((:vlfs_multi_solver_setup, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Ref{VLFSB200.SolverOpts}), m.ctxs, m.ngpu, 0, o)
end
function multi_solve_vec!(m::MultiGPUSystem, xs::Vector{<:Vector})      # xs[r]: n_local(r) entries
  ptrs = [Ptr{Cvoid}(pointer(x)) for x in xs]
  GC.@preserve xs ccall((:vlfs_multi_solve_vec, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}), m.ctxs, m.ngpu, 0, ptrs, C_NULL)
end

end # module
