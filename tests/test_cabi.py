"""The C-ABI library loads and exports every symbol include/vlfs_b200.h declares; without a
GPU the context constructor fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import pytest
import vlfs_b200
from vlfs_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vlfs_b200.h")).read()
    return sorted(set(re.findall(r"\b(vlfs_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = L.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.vlfs_version() >= 100


def test_struct_layouts_match_header_sizes():
    # 4 int32 + 9 pointers ; see vlfs_slot_desc
    assert C.sizeof(L.SlotDesc) == 16 + 9 * 8
    assert C.sizeof(L.TermDesc) == 8 + 16 + 8 + 16 + 16 + 24 + 24
    assert C.sizeof(L.SolveStats) == 8 + 8 + 16 + 8


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.VlfsError):
        vlfs_b200.B200System(10, False)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every declared entry point to the reference interface it replaces."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [name for name in _declared() if name not in doc]
    assert not missing, missing


def test_julia_shim_binds_the_reference_facing_calls():
    """The (untested) ccall shim covers the calls a maintainer of the reference needs; the
    device-pointer building blocks of the multi-GPU path and the benchmark hooks are Python-only."""
    jl = open(os.path.join(ROOT, "julia", "VLFSB200.jl")).read()
    needed = ["vlfs_create", "vlfs_destroy", "vlfs_last_error", "vlfs_system_init", "vlfs_add_domain", "vlfs_add_term",
              "vlfs_add_rhs_term", "vlfs_set_term_coef", "vlfs_set_rhs_coef", "vlfs_set_weights", "vlfs_build_pattern",
              "vlfs_get_pattern_csc", "vlfs_assemble", "vlfs_get_values", "vlfs_get_vector", "vlfs_combine", "vlfs_spmv",
              "vlfs_set_dof_coords", "vlfs_solver_setup", "vlfs_solver_refactor", "vlfs_solve", "vlfs_solve_vec",
              "vlfs_newmark_run", "vlfs_eval_functional", "vlfs_add_work_matrix", "vlfs_precond_apply",
              "vlfs_assemble_vectors", "vlfs_assemble_matrices"]
    assert [n for n in needed if ":" + n not in jl] == []
    assert "UNTESTED" in jl


def _harness(tmp_path_factory):
    import subprocess
    exe = str(tmp_path_factory.mktemp("cabi") / "cabi_harness")
    pkg = os.path.join(ROOT, "monolithicfemvlfs.jl_b200")
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "cabi_harness.cpp"), "-o", exe,
                    "-L" + pkg, "-lvlfs_b200", "-Wl,-rpath," + pkg], check=True)
    return exe


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    return _harness(tmp_path_factory)


def test_struct_layouts_of_the_header_match_ctypes_and_julia(harness):
    """A C++ program compiled against include/vlfs_b200.h reports sizeof/offsetof of every struct; the
    ctypes mirror (what the tests call through) must agree field for field, and the Julia shim declares
    the same fields in the same order."""
    import subprocess
    out = subprocess.run([harness, "--layout"], capture_output=True, text=True, check=True).stdout
    got = {}
    for line in out.strip().splitlines():
        t = line.split()
        got[t[0]] = dict(size=int(t[1]), **{t[k]: int(t[k + 1]) for k in range(2, len(t), 2)})
    pairs = {"vlfs_slot_desc": L.SlotDesc, "vlfs_domain_desc": L.DomainDesc, "vlfs_term_desc": L.TermDesc,
             "vlfs_rhs_desc": L.RhsDesc, "vlfs_solver_opts": L.SolverOpts, "vlfs_solve_stats": L.SolveStats,
             "vlfs_functional_desc": L.FunctionalDesc}
    for name, cls in pairs.items():
        assert C.sizeof(cls) == got[name]["size"], name
        for field, off in got[name].items():
            if field != "size":
                assert getattr(cls, field).offset == off, (name, field)
    jl = open(os.path.join(ROOT, "julia", "VLFSB200.jl")).read()
    for cls, jname in ((L.SlotDesc, "SlotDesc"), (L.TermDesc, "TermDesc"), (L.RhsDesc, "RhsDesc"), (L.SolverOpts, "SolverOpts")):
        body = jl[jl.index("struct " + jname):]
        body = body[:body.index("\nend")]
        order = [f for f, _ in cls._fields_]
        pos = [re.search(r"\b" + f + r"::", body).start() for f in order]
        assert pos == sorted(pos), jname


def test_harness_refuses_to_run_without_a_gpu(harness):
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert subprocess.run([harness, "--run"]).returncode == 77


@pytest.mark.gpu
def test_cabi_harness_on_the_gpu(tmp_path_factory):
    """The ABI exercised from C++ exactly as `ccall` would: assembly from hand-written tables against
    the analytic matrix, SpMV, solve, the imported-matrix LinearSolver path and a two-rank system."""
    import subprocess
    exe = _harness(tmp_path_factory)
    out = subprocess.run([exe, "--run"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all checks passed" in out.stdout
