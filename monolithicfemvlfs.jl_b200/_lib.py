"""ctypes binding of libvlfs_b200.so (the C ABI declared in include/vlfs_b200.h).

This is the Python twin of the Julia `ccall` shim shown in INTEGRATION.md.  There is no
CPU fallback: if the shared library is missing, or no sm_100 GPU is usable, every compute
entry point raises.
"""
import ctypes as C
import os
import numpy as np

MAX_SLOTS = 2
OP_VAL, OP_GRAD, OP_DDIR, OP_LAPL, OP_HESS, OP_CHESS, OP_GRAD_NOWN, OP_CHESS_NPLUS = range(8)
MEASURE_CELL, MEASURE_FACET = 0, 1
SOLVER_GMRES, SOLVER_BICGSTAB, SOLVER_CG = 0, 1, 2
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_ILU0, PRECOND_SPARSE_LU = 0, 1, 2, 3
OK, NOT_CONVERGED, PIVOT_PERTURBED, ERR_SINGULAR = 0, 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class SlotDesc(C.Structure):
    _fields_ = [("dref", C.c_int32), ("nv", C.c_int32), ("nvar", C.c_int32), ("ndofs", C.c_int32),
                ("coords", _dp), ("variant", _ip), ("geo_grad", _dp), ("val", _dp), ("grad", _dp),
                ("hess", _dp), ("dofs", _ip), ("ref_normal", _dp), ("ref_tangent", _dp)]


class DomainDesc(C.Structure):
    _fields_ = [("D", C.c_int32), ("nq", C.c_int32), ("nslots", C.c_int32),
                ("measure_slot", C.c_int32), ("measure_kind", C.c_int32), ("nweights", C.c_int32),
                ("affine", C.c_int32), ("ncells", C.c_int64), ("qweights", _dp), ("weights", _dp),
                ("C", _dp), ("slot", SlotDesc * MAX_SLOTS)]


class TermDesc(C.Structure):
    _fields_ = [("mat", C.c_int32), ("weight", C.c_int32), ("coef_re", C.c_double),
                ("coef_im", C.c_double), ("test_op", C.c_int32), ("trial_op", C.c_int32),
                ("test_scale", C.c_double * MAX_SLOTS), ("trial_scale", C.c_double * MAX_SLOTS),
                ("test_dir", C.c_double * 3), ("trial_dir", C.c_double * 3)]


class RhsDesc(C.Structure):
    _fields_ = [("vec", C.c_int32), ("weight_re", C.c_int32), ("weight_im", C.c_int32),
                ("coef_re", C.c_double), ("coef_im", C.c_double), ("op", C.c_int32),
                ("scale", C.c_double * MAX_SLOTS), ("dir", C.c_double * 3)]


class FunctionalDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("weight_f", C.c_int32), ("scale", C.c_double * MAX_SLOTS),
                ("dir", C.c_double * 3)]


class SolverOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("precond", C.c_int32), ("restart", C.c_int32),
                ("maxit", C.c_int32), ("rtol", C.c_double), ("block_size_hint", C.c_int32),
                ("sweeps", C.c_int32)]


class SolveStats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("rel_residual", C.c_double),
                ("setup_ms", C.c_double), ("solve_ms", C.c_double), ("backward_error", C.c_double)]


LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvlfs_b200.so")

# every symbol include/vlfs_b200.h declares
EXPORTS = [
    "vlfs_version", "vlfs_device_count", "vlfs_create", "vlfs_destroy", "vlfs_last_error",
    "vlfs_system_init", "vlfs_add_domain", "vlfs_add_term", "vlfs_add_rhs_term",
    "vlfs_set_term_coef", "vlfs_set_rhs_coef", "vlfs_set_weights", "vlfs_build_pattern",
    "vlfs_get_pattern", "vlfs_get_pattern_csc", "vlfs_assemble", "vlfs_assemble_matrices",
    "vlfs_assemble_vectors", "vlfs_get_values", "vlfs_get_vector", "vlfs_combine", "vlfs_spmv",
    "vlfs_set_dof_coords", "vlfs_solver_setup", "vlfs_solver_refactor", "vlfs_direct_info", "vlfs_solve", "vlfs_solve_vec", "vlfs_newmark_run", "vlfs_spmv_device",
    "vlfs_assemble_timed", "vlfs_assemble_profile", "vlfs_last_kernel_launches", "vlfs_device_ptrs",
    "vlfs_set_stream", "vlfs_set_owned_rows", "vlfs_dev_spmv", "vlfs_dev_precond", "vlfs_dev_multidot",
    "vlfs_dev_gemv_sub", "vlfs_dev_axpby", "vlfs_dev_gather", "vlfs_get_vector_device",
    "vlfs_eval_functional", "vlfs_set_roles", "vlfs_schur_size", "vlfs_schur_get", "vlfs_schur_forward",
    "vlfs_schur_backward", "vlfs_dense_factor", "vlfs_chain_factor", "vlfs_dense_solve",
    "vlfs_direct_status", "vlfs_import_csr", "vlfs_set_values", "vlfs_set_vector",
    "vlfs_dist_unique_id", "vlfs_dist_init", "vlfs_multi_create", "vlfs_dist_set_halo", "vlfs_dist_info",
    "vlfs_multi_build_pattern", "vlfs_multi_assemble", "vlfs_multi_solver_setup", "vlfs_multi_solver_refactor",
    "vlfs_multi_solve_vec", "vlfs_multi_spmv", "vlfs_multi_spmv_device", "vlfs_measure_fp64_peak",
    "vlfs_add_work_matrix", "vlfs_precond_apply", "vlfs_precond_info",
]

_lib = None


class VlfsError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VlfsError("libvlfs_b200.so not built: run `python -c 'import __graft_entry__ as g; "
                        "g.build()'` (needs nvcc); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.vlfs_version.restype = C.c_int
    lib.vlfs_last_error.restype = C.c_char_p
    lib.vlfs_last_error.argtypes = [vp]
    lib.vlfs_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.vlfs_create.argtypes = [C.POINTER(vp), C.c_int]
    lib.vlfs_destroy.argtypes = [vp]
    lib.vlfs_system_init.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int]
    lib.vlfs_add_domain.argtypes = [vp, C.POINTER(DomainDesc), C.POINTER(C.c_int)]
    lib.vlfs_add_term.argtypes = [vp, C.c_int, C.POINTER(TermDesc), C.POINTER(C.c_int)]
    lib.vlfs_add_rhs_term.argtypes = [vp, C.c_int, C.POINTER(RhsDesc), C.POINTER(C.c_int)]
    lib.vlfs_set_term_coef.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.vlfs_set_rhs_coef.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
    lib.vlfs_set_weights.argtypes = [vp, C.c_int, C.c_int, _dp]
    lib.vlfs_build_pattern.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.vlfs_get_pattern.argtypes = [vp, _ip, _ip]
    lib.vlfs_get_pattern_csc.argtypes = [vp, _ip, _ip, _ip]
    lib.vlfs_assemble.argtypes = [vp]
    lib.vlfs_assemble_matrices.argtypes = [vp]
    lib.vlfs_assemble_vectors.argtypes = [vp]
    lib.vlfs_get_values.argtypes = [vp, C.c_int, vp]
    lib.vlfs_get_vector.argtypes = [vp, C.c_int, vp]
    lib.vlfs_combine.argtypes = [vp, C.c_int, C.c_int, _ip, _dp]
    lib.vlfs_add_work_matrix.argtypes = [vp, C.POINTER(C.c_int)]
    lib.vlfs_precond_apply.argtypes = [vp, vp, vp]
    lib.vlfs_precond_info.argtypes = [vp, _dp]
    lib.vlfs_spmv.argtypes = [vp, C.c_int, vp, vp]
    lib.vlfs_solver_setup.argtypes = [vp, C.c_int, C.POINTER(SolverOpts)]
    lib.vlfs_set_dof_coords.argtypes = [vp, C.c_int, _dp]
    lib.vlfs_solver_refactor.argtypes = [vp, C.c_int]
    lib.vlfs_direct_info.argtypes = [vp, _dp]
    lib.vlfs_solve.argtypes = [vp, vp, vp, C.POINTER(SolveStats)]
    lib.vlfs_solve_vec.argtypes = [vp, C.c_int, vp, C.POINTER(SolveStats)]
    lib.vlfs_newmark_run.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, _dp,
                                     C.c_int64, C.c_int64, C.c_int, _dp, C.POINTER(SolveStats)]
    lib.vlfs_spmv_device.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.vlfs_assemble_timed.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.vlfs_assemble_profile.argtypes = [vp, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    lib.vlfs_last_kernel_launches.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.vlfs_device_ptrs.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.vlfs_set_stream.argtypes = [vp, vp]
    lib.vlfs_set_owned_rows.argtypes = [vp, C.c_int64]
    lib.vlfs_dev_spmv.argtypes = [vp, C.c_int, vp, vp]
    lib.vlfs_dev_precond.argtypes = [vp, vp, vp]
    lib.vlfs_dev_multidot.argtypes = [vp, vp, C.c_int64, C.c_int, vp, C.c_int64, vp]
    lib.vlfs_dev_gemv_sub.argtypes = [vp, vp, C.c_int64, C.c_int, vp, vp, C.c_int64]
    lib.vlfs_dev_axpby.argtypes = [vp, C.c_double, C.c_double, vp, C.c_double, C.c_double, vp, C.c_int64]
    lib.vlfs_dev_gather.argtypes = [vp, C.c_int64, vp, vp, vp]
    lib.vlfs_get_vector_device.argtypes = [vp, C.c_int, C.POINTER(vp)]
    lib.vlfs_eval_functional.argtypes = [vp, C.c_int, C.POINTER(FunctionalDesc), vp, _dp]
    lib.vlfs_set_roles.argtypes = [vp, C.POINTER(C.c_int8)]
    lib.vlfs_schur_size.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.vlfs_schur_get.argtypes = [vp, vp]
    lib.vlfs_schur_forward.argtypes = [vp, vp, vp]
    lib.vlfs_schur_backward.argtypes = [vp, vp, vp]
    lib.vlfs_dense_factor.argtypes = [vp, C.c_int64, vp]
    lib.vlfs_chain_factor.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), vp]
    lib.vlfs_dense_solve.argtypes = [vp, vp, vp]
    i64p = C.POINTER(C.c_int64)
    lib.vlfs_dist_unique_id.argtypes = [vp]
    lib.vlfs_dist_init.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.vlfs_multi_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(C.c_int)]
    lib.vlfs_dist_set_halo.argtypes = [vp, C.c_int64, C.c_int, _ip, i64p, _ip, i64p]
    lib.vlfs_dist_info.argtypes = [vp, _dp]
    lib.vlfs_multi_build_pattern.argtypes = [C.POINTER(vp), C.c_int, i64p, i64p]
    lib.vlfs_multi_assemble.argtypes = [C.POINTER(vp), C.c_int]
    lib.vlfs_multi_solver_setup.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(SolverOpts)]
    lib.vlfs_multi_solver_refactor.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.vlfs_multi_solve_vec.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp), C.POINTER(SolveStats)]
    lib.vlfs_multi_spmv.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp)]
    lib.vlfs_multi_spmv_device.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.vlfs_measure_fp64_peak.argtypes = [vp, _dp]
    lib.vlfs_direct_status.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.vlfs_import_csr.argtypes = [vp, C.c_int64, C.c_int, C.c_int, _ip, _ip]
    lib.vlfs_set_values.argtypes = [vp, C.c_int, vp]
    lib.vlfs_set_vector.argtypes = [vp, C.c_int, vp]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def iptr(a):
    return a.ctypes.data_as(_ip) if a is not None else None
