"""ctypes front end of the compiled CPU restatement of the assembly (oracle/c/asm_cpu.cpp; TEST
INFRASTRUCTURE / CPU BASELINE).  `CSystem` has the surface of `RecordingSystem` (and therefore of
the product's `B200System` as far as the case modules use it); `assemble()` marshals the recorded
domains into the descriptor structs of include/vlfs_b200.h and integrates + compresses them on the
host cores.  Checked against the NumPy interpreter `RecordingSystem.evaluate` in
tests/test_oracle_c.py; timed by bench.py's CPU legs.  Never imported by the product."""
import ctypes as C
import os
import subprocess
import numpy as np
import scipy.sparse as sp
from .descriptor_eval import RecordingSystem

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _structs():
    # layouts of the public header, shared with the product's binding
    from vlfs_b200 import _lib as L
    return L


def load(build=True):
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "c", "liboracle_asm.so")
    if not os.path.exists(path):
        if not build:
            raise RuntimeError("oracle/c/liboracle_asm.so is not built (make -C oracle/c)")
        subprocess.check_call(["make", "-C", os.path.join(HERE, "c")])
    L = _structs()

    class OracleDomain(C.Structure):
        _fields_ = [("desc", C.POINTER(L.DomainDesc)), ("nterms", C.c_int32), ("terms", C.POINTER(L.TermDesc)),
                    ("nrhs", C.c_int32), ("rhs", C.POINTER(L.RhsDesc))]

    lib = C.CDLL(path)
    lib.OracleDomain = OracleDomain
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.oracle_assemble.restype = C.c_int
    lib.oracle_assemble.argtypes = [C.c_int32, C.POINTER(OracleDomain), C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                    C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(ip),
                                    C.POINTER(ip), C.POINTER(dp), dp, dp]
    lib.oracle_spmv.restype = C.c_double
    lib.oracle_spmv.argtypes = [C.c_int64, ip, ip, dp, C.c_int32, dp, dp, C.c_int32, C.c_int32]
    lib.oracle_free.argtypes = [C.c_void_p]
    lib.oracle_free.restype = None
    _lib = lib
    return lib


class CSystem(RecordingSystem):
    """RecordingSystem whose `assemble` runs the compiled restatement on `threads` host cores
    (0 = all)."""

    threads = 0

    def assemble(self):
        L = _structs()
        lib = load()
        keep, odoms = [], (lib.OracleDomain * len(self.domains))()
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None
        iptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None
        for k, d in enumerate(self.domains):
            dd = L.DomainDesc()
            dd.D, dd.nq, dd.nslots = d.D, d.nq, len(d.slots)
            dd.measure_slot, dd.measure_kind = d.measure_slot, d.measure_kind
            dd.nweights, dd.affine, dd.ncells = len(d.weights), 0, d.ncells
            qw = f64(d.qweights); keep.append(qw); dd.qweights = dptr(qw)
            if len(d.weights):
                w = f64(np.stack(d.weights)); keep.append(w); dd.weights = dptr(w)
            if d.C is not None:
                ct = f64(d.C); keep.append(ct); dd.C = dptr(ct)
            for si, s in enumerate(d.slots):
                sd = dd.slot[si]
                sd.dref, sd.nv, sd.nvar, sd.ndofs = s["dref"], s["nv"], s["nvar"], s["ndofs"]
                arrs = {n: (f64(s[n]) if s.get(n) is not None else None)
                        for n in ("coords", "geo_grad", "val", "grad", "hess", "ref_normal", "ref_tangent")}
                arrs.update({n: (i32(s[n]) if s.get(n) is not None else None) for n in ("variant", "dofs")})
                keep.append(arrs)
                sd.coords, sd.geo_grad, sd.val = dptr(arrs["coords"]), dptr(arrs["geo_grad"]), dptr(arrs["val"])
                sd.grad, sd.hess = dptr(arrs["grad"]), dptr(arrs["hess"])
                sd.ref_normal, sd.ref_tangent = dptr(arrs["ref_normal"]), dptr(arrs["ref_tangent"])
                sd.variant, sd.dofs = iptr(arrs["variant"]), iptr(arrs["dofs"])
            terms = (L.TermDesc * max(len(d.terms), 1))()
            for ti, t in enumerate(d.terms):
                T = terms[ti]
                T.mat, T.weight = t["mat"], t["weight"]
                T.coef_re, T.coef_im = t["coef"].real, t["coef"].imag
                T.test_op, T.trial_op = t["test_op"], t["trial_op"]
                for q in range(len(d.slots)):
                    T.test_scale[q], T.trial_scale[q] = t["st"][q], t["su"][q]
                for q in range(3):
                    T.test_dir[q], T.trial_dir[q] = t["test_dir"][q], t["trial_dir"][q]
            rhs = (L.RhsDesc * max(len(d.rhs), 1))()
            for ri, r in enumerate(d.rhs):
                R = rhs[ri]
                R.vec, R.weight_re, R.weight_im = r["vec"], r["wre"], r["wim"]
                R.coef_re, R.coef_im, R.op = r["coef"].real, r["coef"].imag, r["op"]
                for q in range(len(d.slots)):
                    R.scale[q] = r["sc"][q]
                for q in range(3):
                    R.dir[q] = r["dir"][q]
            keep += [dd, terms, rhs]
            odoms[k].desc = C.pointer(dd)
            odoms[k].nterms, odoms[k].terms = len(d.terms), terms
            odoms[k].nrhs, odoms[k].rhs = len(d.rhs), rhs
        sc = 2 if self.is_complex else 1
        nnz, ncoo = C.c_int64(0), C.c_int64(0)
        rp, ci = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
        vals = (C.POINTER(C.c_double) * self.nmat)()
        vecs = np.zeros(self.nvec * self.ndofs * sc)
        ms = np.zeros(3)
        rc = lib.oracle_assemble(len(self.domains), odoms, self.ndofs, int(self.is_complex), self.nmat, self.nvec,
                                 int(self.threads), C.byref(nnz), C.byref(ncoo), C.byref(rp), C.byref(ci), vals,
                                 vecs.ctypes.data_as(C.POINTER(C.c_double)), ms.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise RuntimeError("oracle_assemble failed (%d)" % rc)
        n = nnz.value
        rowptr = np.ctypeslib.as_array(rp, shape=(self.ndofs + 1,)).copy()
        colind = np.ctypeslib.as_array(ci, shape=(max(n, 1),))[:n].copy()
        self.mats = []
        for m in range(self.nmat):
            v = np.ctypeslib.as_array(vals[m], shape=(max(n, 1) * sc,))[:n * sc].copy()
            v = v.view(np.complex128) if self.is_complex else v
            self.mats.append(sp.csr_matrix((v, colind, rowptr), shape=(self.ndofs, self.ndofs)))
            lib.oracle_free(vals[m])
        lib.oracle_free(rp); lib.oracle_free(ci)
        v = vecs.reshape(self.nvec, self.ndofs * sc)
        self.vecs = [(v[k].view(np.complex128) if self.is_complex else v[k]).copy() for k in range(self.nvec)]
        self.nnz, self.ncoo = n, ncoo.value
        self.ms = dict(integrate_and_fill=ms[0], compress=ms[1], total=ms[2])
        return self.mats, self.vecs


def spmv_ms(A, x, reps=10, threads=0):
    """Mean milliseconds of y = A x (SciPy CSR, real or complex) on the host threads, and y."""
    lib = load()
    A = A.tocsr()
    cplx = np.iscomplexobj(A.data)
    dt = np.complex128 if cplx else np.float64
    vals = np.ascontiguousarray(A.data, dtype=dt)
    xv = np.ascontiguousarray(x, dtype=dt)
    y = np.zeros(A.shape[0], dtype=dt)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    ms = lib.oracle_spmv(A.shape[0], rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), vals.view(np.float64).ctypes.data_as(dp),
                         int(cplx), xv.view(np.float64).ctypes.data_as(dp), y.view(np.float64).ctypes.data_as(dp),
                         int(reps), int(threads))
    return float(ms), y
