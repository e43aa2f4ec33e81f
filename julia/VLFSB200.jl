# VLFSB200.jl - thin Julia shim over libvlfs_b200.so (include/vlfs_b200.h).
#
# STATUS: UNTESTED.  Neither this container nor the GPU box has a Julia toolchain (SURVEY.md
# §0.3), so this file has never been executed.  It is the mechanical `ccall` twin of the ctypes
# binding `monolithicfemvlfs.jl_b200/_lib.py`, which IS exercised by the test-suite; the struct
# layouts below mirror that file field by field.
#
# What it plugs into (Gridap 0.17 extension points, SURVEY.md §8b):
#   * `B200LUSolver <: Gridap.Algebra.LinearSolver` - drop-in for `LUSolver()` in
#     `solve(LinearFESolver(ls),op)` and `Newmark(ls,dt,γ,β)`
#     (src/Periodic_Beam.jl:107, src/Khabakhpasheva_time_domain.jl:258, src/Yago_freq_domain.jl:171).
#   * `b200_affine_operator(sys,X,Y)` - returns an ordinary `AffineFEOperator(trial,test,A,b)` whose
#     matrix and vector were integrated and assembled on the GPU (src/Yago_freq_domain.jl:167).
module VLFSB200

using SparseArrays

const LIB = get(ENV, "VLFS_B200_LIB", "libvlfs_b200.so")
const MAX_SLOTS = 2

# --- operator / enum codes (vlfs_b200.h) ---------------------------------------------------
const OP_VAL, OP_GRAD, OP_DDIR, OP_LAPL, OP_HESS, OP_CHESS, OP_GRAD_NOWN, OP_CHESS_NPLUS = Int32.(0:7)
const MEASURE_CELL, MEASURE_FACET = Int32(0), Int32(1)
const SOLVER_GMRES, SOLVER_BICGSTAB, SOLVER_CG = Int32.(0:2)
const PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_ILU0, PRECOND_SPARSE_LU = Int32.(0:3)
const VLFS_OK, VLFS_NOT_CONVERGED = Int32(0), Int32(3)

# --- C structs (isbits, same field order as the header) ------------------------------------
struct SlotDesc
  dref::Int32; nv::Int32; nvar::Int32; ndofs::Int32
  coords::Ptr{Float64}; variant::Ptr{Int32}; geo_grad::Ptr{Float64}; val::Ptr{Float64}
  grad::Ptr{Float64}; hess::Ptr{Float64}; dofs::Ptr{Int32}
  ref_normal::Ptr{Float64}; ref_tangent::Ptr{Float64}
end
const NULL_SLOT = SlotDesc(0, 0, 0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL)

struct DomainDesc
  D::Int32; nq::Int32; nslots::Int32; measure_slot::Int32; measure_kind::Int32
  nweights::Int32; affine::Int32
  ncells::Int64
  qweights::Ptr{Float64}; weights::Ptr{Float64}; C::Ptr{Float64}
  slot::NTuple{MAX_SLOTS,SlotDesc}
end

struct TermDesc
  mat::Int32; weight::Int32
  coef_re::Float64; coef_im::Float64
  test_op::Int32; trial_op::Int32
  test_scale::NTuple{MAX_SLOTS,Float64}; trial_scale::NTuple{MAX_SLOTS,Float64}
  test_dir::NTuple{3,Float64}; trial_dir::NTuple{3,Float64}
end

struct RhsDesc
  vec::Int32; weight_re::Int32; weight_im::Int32
  coef_re::Float64; coef_im::Float64
  op::Int32
  scale::NTuple{MAX_SLOTS,Float64}
  dir::NTuple{3,Float64}
end

struct SolverOpts
  method::Int32; precond::Int32; restart::Int32; maxit::Int32
  rtol::Float64
  block_size_hint::Int32; sweeps::Int32
end

mutable struct SolveStats
  iterations::Int32; converged::Int32
  rel_residual::Float64; setup_ms::Float64; solve_ms::Float64; backward_error::Float64
  SolveStats() = new(0, 0, 0.0, 0.0, 0.0, 0.0)
end

# --- context ---------------------------------------------------------------------------------
mutable struct System
  ctx::Ptr{Cvoid}
  ndofs::Int
  iscomplex::Bool
  nnz::Int
  function System(ndofs::Integer; iscomplex::Bool, nmat::Integer=1, nvec::Integer=1, device::Integer=0)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:vlfs_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
    rc == 0 || error("vlfs_create failed (status $rc): no usable sm_100 GPU; there is no CPU fallback")
    s = new(ref[], ndofs, iscomplex, 0)
    finalizer(x -> ccall((:vlfs_destroy, LIB), Cint, (Ptr{Cvoid},), x.ctx), s)
    check(s, ccall((:vlfs_system_init, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Cint, Cint),
                   s.ctx, ndofs, iscomplex, nmat, nvec))
    s
  end
end

function check(s::System, rc; allow=(VLFS_OK,))
  rc in allow && return rc
  msg = unsafe_string(ccall((:vlfs_last_error, LIB), Cstring, (Ptr{Cvoid},), s.ctx))
  error("libvlfs_b200 status $rc: $msg")
end

# `slots`: vector of NamedTuples with the fields of SlotDesc as Julia arrays (0-based DOF ids,
# multi-field offset applied).  All arrays stay rooted by GC.@preserve during the call; the
# library copies and never keeps a host pointer.
function add_domain!(s::System; D, nq, qweights::Vector{Float64}, slots, measure_slot=0,
                     measure_kind=MEASURE_CELL, weights::Matrix{Float64}=zeros(0, 0),
                     Ctensor::Union{Nothing,Vector{Float64}}=nothing, affine::Bool=false, ncells::Integer)
  p(a) = a === nothing ? Ptr{eltype(Float64)}(C_NULL) : pointer(a)
  GC.@preserve qweights weights Ctensor slots begin
    sd = ntuple(MAX_SLOTS) do k
      k > length(slots) && return NULL_SLOT
      t = slots[k]
      SlotDesc(t.dref, t.nv, t.nvar, t.ndofs, pointer(t.coords),
               t.variant === nothing ? Ptr{Int32}(C_NULL) : pointer(t.variant),
               pointer(t.geo_grad), pointer(t.val), pointer(t.grad),
               t.hess === nothing ? Ptr{Float64}(C_NULL) : pointer(t.hess), pointer(t.dofs),
               t.ref_normal === nothing ? Ptr{Float64}(C_NULL) : pointer(t.ref_normal),
               t.ref_tangent === nothing ? Ptr{Float64}(C_NULL) : pointer(t.ref_tangent))
    end
    d = DomainDesc(D, nq, length(slots), measure_slot, measure_kind, size(weights, 2) == 0 ? 0 : size(weights, 1),
                   affine, ncells, pointer(qweights), isempty(weights) ? Ptr{Float64}(C_NULL) : pointer(weights),
                   Ctensor === nothing ? Ptr{Float64}(C_NULL) : pointer(Ctensor), sd)
    id = Ref{Cint}(-1)
    check(s, ccall((:vlfs_add_domain, LIB), Cint, (Ptr{Cvoid}, Ref{DomainDesc}, Ref{Cint}), s.ctx, d, id))
    return Int(id[])
  end
end

scales(spec::Integer) = ntuple(k -> k == spec + 1 ? 1.0 : 0.0, MAX_SLOTS)
scales(spec::Dict) = ntuple(k -> Float64(get(spec, k - 1, 0.0)), MAX_SLOTS)

function add_term!(s::System, dom::Integer; mat=0, coef, test_op, test_slots, trial_op, trial_slots,
                   weight=-1, test_dir=(0.0, 0.0, 0.0), trial_dir=(0.0, 0.0, 0.0))
  t = TermDesc(mat, weight, real(coef), imag(coef), test_op, trial_op, scales(test_slots), scales(trial_slots),
               Float64.(test_dir), Float64.(trial_dir))
  id = Ref{Cint}(-1)
  check(s, ccall((:vlfs_add_term, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{TermDesc}, Ref{Cint}), s.ctx, dom, t, id))
  Int(id[])
end

function add_rhs!(s::System, dom::Integer; vec=0, coef, op, slots, weight_re=-1, weight_im=-1, dir=(0.0, 0.0, 0.0))
  r = RhsDesc(vec, weight_re, weight_im, real(coef), imag(coef), op, scales(slots), Float64.(dir))
  id = Ref{Cint}(-1)
  check(s, ccall((:vlfs_add_rhs_term, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{RhsDesc}, Ref{Cint}), s.ctx, dom, r, id))
  Int(id[])
end

set_coef!(s, dom, term, c) = check(s, ccall((:vlfs_set_term_coef, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble), s.ctx, dom, term, real(c), imag(c)))
set_weights!(s, dom, w, v::Vector{Float64}) = check(s, ccall((:vlfs_set_weights, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), s.ctx, dom, w, v))

function build_pattern!(s::System)
  nnz, ncoo = Ref{Int64}(0), Ref{Int64}(0)
  check(s, ccall((:vlfs_build_pattern, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), s.ctx, nnz, ncoo))
  s.nnz = nnz[]
end

assemble!(s::System) = check(s, ccall((:vlfs_assemble, LIB), Cint, (Ptr{Cvoid},), s.ctx))
# only the matrices / only the right-hand sides (a sweep by basis matrices re-integrates the vectors alone)
assemble_matrices!(s::System) = check(s, ccall((:vlfs_assemble_matrices, LIB), Cint, (Ptr{Cvoid},), s.ctx))
assemble_vectors!(s::System) = check(s, ccall((:vlfs_assemble_vectors, LIB), Cint, (Ptr{Cvoid},), s.ctx))

# The matrix as Gridap's default type: SparseMatrixCSC{T,Int64}, 1-based, explicit zeros kept.
function get_matrix(s::System, mat::Integer=0)
  T = s.iscomplex ? ComplexF64 : Float64
  colptr = Vector{Int32}(undef, s.ndofs + 1); rowval = Vector{Int32}(undef, s.nnz); slot = Vector{Int32}(undef, s.nnz)
  check(s, ccall((:vlfs_get_pattern_csc, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}), s.ctx, colptr, rowval, slot))
  vals = Vector{T}(undef, s.nnz)
  check(s, ccall((:vlfs_get_values, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.ctx, mat, vals))
  SparseMatrixCSC{T,Int64}(s.ndofs, s.ndofs, Int64.(colptr) .+ 1, Int64.(rowval) .+ 1, vals[slot .+ 1])
end

function get_vector(s::System, vec::Integer=0)
  b = Vector{s.iscomplex ? ComplexF64 : Float64}(undef, s.ndofs)
  check(s, ccall((:vlfs_get_vector, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), s.ctx, vec, b))
  b
end

# block_rows > 0 with PRECOND_BLOCK_ILU0: block Jacobi of ILU(0) blocks of that many consecutive rows
function solver_setup!(s::System, mat=0; method=SOLVER_GMRES, precond=PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12,
                       block_rows=0)
  o = SolverOpts(method, precond, restart, maxit, rtol, block_rows, 0)
  check(s, ccall((:vlfs_solver_setup, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{SolverOpts}), s.ctx, mat, o); allow=(VLFS_OK, 4))
end

# z = M⁻¹ r with the preconditioner of the last solver_setup! (Jacobi, ILU(0), sparse LU)
function precond_apply(s::System, r::Vector)
  z = similar(r)
  check(s, ccall((:vlfs_precond_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), s.ctx, r, z))
  z
end

function solve!(x::Vector, s::System, b::Vector)
  st = SolveStats()
  rc = ccall((:vlfs_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{SolveStats}), s.ctx, b, x, st)
  check(s, rc; allow=(VLFS_OK, VLFS_NOT_CONVERGED))
  st
end

set_rhs_coef!(s, dom, term, c) = check(s, ccall((:vlfs_set_rhs_coef, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble), s.ctx, dom, term, real(c), imag(c)))

# DOF coordinates (row-major [ndofs][dim]) for the geometric nested dissection of the sparse LU
set_dof_coords!(s::System, dim::Integer, xyz::Vector{Float64}) =
  check(s, ccall((:vlfs_set_dof_coords, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), s.ctx, dim, xyz))

# A = Σ c_k M_k on the union pattern (the Newmark matrix K + γ/(βΔt) C + 1/(βΔt²) M)
function combine!(s::System, target::Integer, src::Vector{<:Integer}, coef::Vector{<:Number})
  c = collect(Iterators.flatten((Float64(real(z)), Float64(imag(z))) for z in coef))
  check(s, ccall((:vlfs_combine, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Float64}),
                 s.ctx, target, length(src), Cint.(src), c))
end

# a value array no term targets: A(ω) = B₀ + ω B₁ + ω² B₂ + k B₃ of a sweep by basis matrices
function add_work_matrix!(s::System)
  m = Ref{Cint}(-1)
  check(s, ccall((:vlfs_add_work_matrix, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), s.ctx, m))
  Int(m[])
end

function spmv(s::System, x::Vector, mat::Integer=0)
  y = similar(x)
  check(s, ccall((:vlfs_spmv, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}), s.ctx, mat, x, y))
  y
end

# keep the symbolic analysis, refactorise the current values (next ω of a sweep)
refactor!(s::System, mat::Integer=0) = check(s, ccall((:vlfs_solver_refactor, LIB), Cint, (Ptr{Cvoid}, Cint), s.ctx, mat); allow=(VLFS_OK, 4))

function solve_vec!(x::Vector, s::System, vec::Integer=0)
  st = SolveStats()
  rc = ccall((:vlfs_solve_vec, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ref{SolveStats}), s.ctx, vec, x, st)
  check(s, rc; allow=(VLFS_OK, VLFS_NOT_CONVERGED))
  st
end

# `solve(Newmark(ls,dt,γ,β),op,(x₀,v₀,a₀),t₀,tf)` on the device: u0, v0, a0 are advanced in place,
# `out` receives the DOFs out_lo:out_hi-1 (0-based) of every out_stride-th step (row-major);
# tcoef[nb, nsteps] are the time factors of the assembled right-hand-side vectors 0..nb-1
function newmark_run!(s::System, u0::Vector{Float64}, v0::Vector{Float64}, a0::Vector{Float64};
                      matK=0, matC=1, matM=2, nsteps::Integer, dt::Real, γ::Real=0.5, β::Real=0.25,
                      tcoef::Matrix{Float64}=zeros(0, nsteps), out_lo::Integer=0, out_hi::Integer=0, out_stride::Integer=1)
  nb = size(tcoef, 1)
  nout = out_hi > out_lo ? cld(nsteps, out_stride) : 0
  out = Matrix{Float64}(undef, out_hi - out_lo, nout)         # column = one recorded step
  st = SolveStats()
  rc = ccall((:vlfs_newmark_run, LIB), Cint,
             (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cdouble, Cdouble, Cdouble, Cint, Ptr{Float64},
              Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Cint, Ptr{Float64}, Ref{SolveStats}),
             s.ctx, matK, matC, matM, nsteps, dt, γ, β, nb, nb == 0 ? C_NULL : tcoef, u0, v0, a0,
             out_lo, out_hi, out_stride, nout == 0 ? C_NULL : out, st)
  check(s, rc; allow=(VLFS_OK, VLFS_NOT_CONVERGED))
  out, st
end

# ∑∫ |Σ_s scale_s OP(u_h) − f|² dμ on one domain (`∑(∫(x⋅x)dΩ)` of src/Periodic_Beam.jl:119-120)
struct FunctionalDesc
  op::Int32; weight_f::Int32
  scale::NTuple{MAX_SLOTS,Float64}
  dir::NTuple{3,Float64}
end
function eval_functional(s::System, dom::Integer, x::Vector{Float64}; op, slots, weight_f=-1, dir=(0.0, 0.0, 0.0))
  f = FunctionalDesc(op, weight_f, scales(slots), Float64.(dir))
  r = Ref{Float64}(0.0)
  check(s, ccall((:vlfs_eval_functional, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{FunctionalDesc}, Ptr{Float64}, Ref{Float64}),
                 s.ctx, dom, f, x, r))
  r[]
end

# ---------------------------------------------------------------------------------------------
# Gridap glue (needs `using Gridap`; kept in a function so that the module loads without it)
# ---------------------------------------------------------------------------------------------
"""
    b200_affine_operator(sys, X, Y) -> AffineFEOperator

After the terms were registered and `build_pattern!`/`assemble!` ran, wrap the GPU-assembled
matrix and vector in Gridap's standard operator (inner constructor
`AffineFEOperator(trial,test,A,b)`), so `solve`, `get_matrix`, `get_vector`, `FEFunction`
downstream are unchanged.
"""
function b200_affine_operator(Gridap, sys::System, X, Y)
  A = get_matrix(sys, 0)
  b = get_vector(sys, 0)
  Gridap.FESpaces.AffineFEOperator(X, Y, Gridap.Algebra.AffineOperator(A, b))
end

end # module
