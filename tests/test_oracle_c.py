"""Compiled CPU restatement of the assembly (oracle/c/asm_cpu.cpp, the CPU baseline of bench.py)
against the NumPy descriptor interpreter, which tests/test_host_tables.py ties to the oracle's
literal weak forms: same pattern, entries and right-hand sides <= 1e-12, bitwise independent of
the number of host threads - CPU only."""
import os
import numpy as np
import pytest
from oracle.c_assembly import CSystem
from oracle.descriptor_eval import RecordingSystem

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _yago(sc):
    from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params
    return YagoProblem(Yago_freq_domain_params(nx=2, ny=1, nz=2, order=4, lfactor=0.6, dfactor=4), system_cls=sc)


def _khab_freq(sc):
    from vlfs_b200.cases.khabakhpasheva import KhabakhpashevaFreqProblem, Khabakhpasheva_freq_domain_params
    return KhabakhpashevaFreqProblem(Khabakhpasheva_freq_domain_params(nx=10, ny=2, order=4, xi=625.0), system_cls=sc)


def _khab_time(sc):
    from vlfs_b200.cases.khabakhpasheva import KhabakhpashevaTimeProblem, Khabakhpasheva_time_domain_params
    return KhabakhpashevaTimeProblem(Khabakhpasheva_time_domain_params(nx=10, ny=1, order=2, xi=0.0), system_cls=sc)


def _multigeo(sc):
    from vlfs_b200.cases.multigeo import MultiGeoProblem, MultiGeo_freq_domain_params
    return MultiGeoProblem(MultiGeo_freq_domain_params(mesh_file=os.path.join(GOLD, "multi_geo_coarse.json"), order=4),
                           system_cls=sc)


def _pbeam_fs(sc):
    from vlfs_b200.cases.periodic_beam_fs import PeriodicBeamFSProblem, Periodic_Beam_FS_params
    return PeriodicBeamFSProblem(Periodic_Beam_FS_params(n=4, dt=1e-3, tf=2e-3, order=3, k=2), system_cls=sc)


@pytest.mark.parametrize("build", [_yago, _khab_freq, _khab_time, _multigeo, _pbeam_fs],
                         ids=["yago", "khab_freq", "khab_time", "multigeo", "periodic_beam_fs"])
def test_compiled_restatement_equals_numpy_interpreter(build):
    c = build(CSystem)
    mats, vecs = c.sys.assemble()
    r = build(RecordingSystem)
    rmats, rvecs = r.sys.evaluate()
    for A, B in zip(mats, rmats):
        B = B.tocsr(); B.sort_indices()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert np.abs(A.data - B.data).max() <= 1e-12 * max(np.abs(B.data).max(), 1e-300)
    for a, b in zip(vecs, rvecs):
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1e-300)


def test_result_does_not_depend_on_the_number_of_threads():
    out = []
    for threads in (1, 3):
        c = _yago(CSystem)
        c.sys.threads = threads
        mats, vecs = c.sys.assemble()
        out.append((mats[0].indptr.copy(), mats[0].indices.copy(), mats[0].data.copy(), vecs[0].copy()))
    for a, b in zip(*out):
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_threaded_spmv_equals_scipy():
    from oracle.c_assembly import spmv_ms
    c = _yago(CSystem)
    mats, _ = c.sys.assemble()
    A = mats[0]
    n = A.shape[0]
    x = np.sin(0.1 * np.arange(n)) + 1j * np.cos(0.1 * np.arange(n))
    ms, y = spmv_ms(A, x, reps=2, threads=3)
    ref = A @ x
    assert ms > 0 and np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()
    ms, y = spmv_ms(A.real.tocsr(), x.real, reps=1, threads=2)
    assert np.abs(y - A.real @ x.real).max() <= 1e-13 * np.abs(ref).max()
