"""Host-side assembler/operator objects over the C ABI.

`B200System` plays the role of Gridap's `SparseMatrixAssembler(X,Y)` + `AffineFEOperator`
+ `LinearFESolver` for the reference's hot path (src/Yago_freq_domain.jl:167-171,
src/Khabakhpasheva_time_domain.jl:257-266): it registers integration domains and terms,
builds the CSR pattern on the GPU, assembles, multiplies and solves.  All arithmetic is in
libvlfs_b200.so; this file only marshals tables.
"""
import ctypes as C
import numpy as np
import scipy.sparse as sp
from . import _lib as L


def is_affine(coords, poly):
    """True when every cell has a constant Jacobian (simplices; parallelotopes)."""
    if poly.is_simplex or poly.dim <= 1:
        return True
    X = coords
    x0 = X[:, 0]
    edges = [X[:, 1 << a] - x0 for a in range(poly.dim)]
    scale = max(float(np.abs(X - x0[:, None, :]).max()), 1e-300)
    for v, c in enumerate(poly.verts):
        pred = x0 + sum(c[a] * edges[a] for a in range(poly.dim))
        if np.abs(pred - X[:, v]).max() > 1e-12 * scale:
            return False
    return True


def slot_is_affine(coords, dref, nv):
    """Constant Jacobian on every cell of a slot: simplices always; n-cubes (lexicographic
    corners) when every vertex is x0 + sum_a c_a (x_{2^a} - x0)."""
    X = np.asarray(coords, dtype=np.float64)
    if dref <= 1 or nv == dref + 1 or len(X) == 0:
        return True
    if nv != 1 << dref:
        return False
    x0 = X[:, 0]
    edges = [X[:, 1 << a] - x0 for a in range(dref)]
    scale = max(float(np.abs(X - x0[:, None, :]).max()), 1e-300)
    for v in range(nv):
        pred = x0 + sum(edges[a] for a in range(dref) if (v >> a) & 1)
        if np.abs(pred - X[:, v]).max() > 1e-12 * scale:
            return False
    return True


class Domain:
    def __init__(self, system, did, nslots, D):
        self.system, self.id, self.nslots, self.D = system, did, nslots, D
        self.nterms = 0
        self.nrhs = 0
        self.slot_ndofs = []          # bookkeeping for the byte / flop counts of bench.py
        self.table_bytes = 0
        self.term_flops = 0           # flops per (cell, quadrature point) of the bilinear contractions

    def _count_term(self, coef, test_op, test_slots, trial_slots):
        ncomp = {L.OP_VAL: 1, L.OP_DDIR: 1, L.OP_LAPL: 1, L.OP_GRAD_NOWN: 1, L.OP_GRAD: self.D,
                 L.OP_CHESS_NPLUS: self.D}.get(test_op, self.D * (self.D + 1) // 2)
        used = lambda spec: [int(k) for k, v in spec.items() if v != 0] if isinstance(spec, dict) else [int(spec)]
        if not self.slot_ndofs:
            return
        nr = sum(self.slot_ndofs[k] for k in used(test_slots))
        nc = sum(self.slot_ndofs[k] for k in used(trial_slots))
        parts = (np.real(coef) != 0) + (np.imag(coef) != 0)
        self.term_flops += 2 * ncomp * nr * nc * max(int(parts), 1)

    def _scales(self, spec):
        out = (C.c_double * L.MAX_SLOTS)()
        if isinstance(spec, dict):
            for k, v in spec.items():
                out[k] = float(v)
        else:
            out[int(spec)] = 1.0
        return out

    def add_term(self, mat, coef, test_op, test_slots, trial_op, trial_slots, weight=-1,
                 test_dir=(0, 0, 0), trial_dir=(0, 0, 0)):
        """mat += coef ∫ w (Σ_s scale_s OP_test)(Σ_s scale_s OP_trial) dμ; *_slots is a slot
        index or a {slot: scale} dict (skeleton jumps/means)."""
        self._count_term(coef, test_op, test_slots, trial_slots)
        t = L.TermDesc()
        t.mat, t.weight = mat, weight
        t.coef_re, t.coef_im = float(np.real(coef)), float(np.imag(coef))
        t.test_op, t.trial_op = test_op, trial_op
        t.test_scale, t.trial_scale = self._scales(test_slots), self._scales(trial_slots)
        t.test_dir = (C.c_double * 3)(*[float(v) for v in test_dir])
        t.trial_dir = (C.c_double * 3)(*[float(v) for v in trial_dir])
        tid = C.c_int(-1)
        self.system._check(self.system.lib.vlfs_add_term(self.system.ctx, self.id, C.byref(t), C.byref(tid)))
        self.nterms += 1
        return tid.value

    def add_rhs(self, vec, coef, op, slots, weight_re=-1, weight_im=-1, direction=(0, 0, 0)):
        r = L.RhsDesc()
        r.vec, r.weight_re, r.weight_im = vec, weight_re, weight_im
        r.coef_re, r.coef_im = float(np.real(coef)), float(np.imag(coef))
        r.op = op
        r.scale = self._scales(slots)
        r.dir = (C.c_double * 3)(*[float(v) for v in direction])
        tid = C.c_int(-1)
        self.system._check(self.system.lib.vlfs_add_rhs_term(self.system.ctx, self.id, C.byref(r), C.byref(tid)))
        self.nrhs += 1
        return tid.value

    def functional(self, x, op, slots, weight_f=-1, direction=(0, 0, 0)):
        """∑∫ |Σ_s scale_s OP(u_h) − f|² dμ on this domain, u_h from the DOF vector x
        (`∑(∫(x⋅x)dΩ)` of src/Periodic_Beam.jl:119-120 evaluated on the GPU)."""
        f = L.FunctionalDesc()
        f.op, f.weight_f = op, weight_f
        f.scale = self._scales(slots)
        f.dir = (C.c_double * 3)(*[float(v) for v in direction])
        xv = np.ascontiguousarray(x, dtype=np.float64)
        out = C.c_double(0.0)
        self.system._check(self.system.lib.vlfs_eval_functional(
            self.system.ctx, self.id, C.byref(f), xv.ctypes.data_as(C.c_void_p), C.byref(out)))
        return out.value

    def set_coef(self, term, coef):
        self.system._check(self.system.lib.vlfs_set_term_coef(
            self.system.ctx, self.id, term, float(np.real(coef)), float(np.imag(coef))))

    def set_rhs_coef(self, term, coef):
        self.system._check(self.system.lib.vlfs_set_rhs_coef(
            self.system.ctx, self.id, term, float(np.real(coef)), float(np.imag(coef))))

    def set_weights(self, index, values):
        v = L.f64(values)
        self.system._check(self.system.lib.vlfs_set_weights(self.system.ctx, self.id, index, L.dptr(v)))


class B200System:
    def __init__(self, ndofs, is_complex, nmat=1, nvec=1, device=0):
        self.lib = L.load()
        self.ctx = C.c_void_p()
        rc = self.lib.vlfs_create(C.byref(self.ctx), device)
        if rc != 0:
            raise L.VlfsError("vlfs_create failed (status %d): no usable sm_100 GPU; the B200 "
                              "path has no CPU fallback" % rc)
        self.ndofs, self.is_complex, self.nmat, self.nvec = int(ndofs), bool(is_complex), nmat, nvec
        self.dtype = np.complex128 if is_complex else np.float64
        self._check(self.lib.vlfs_system_init(self.ctx, self.ndofs, int(self.is_complex), nmat, nvec))
        self.nnz = None
        self.ncoo = None
        self.domains = []
        self._pattern = None

    @classmethod
    def adopt(cls, ctx, ndofs, is_complex, nmat=1, nvec=1):
        """A system on a context somebody else created (the ranks of `vlfs_multi_create`)."""
        self = cls.__new__(cls)
        self.lib = L.load()
        self.ctx = ctx
        self.owns_ctx = False
        self.ndofs, self.is_complex, self.nmat, self.nvec = int(ndofs), bool(is_complex), nmat, nvec
        self.dtype = np.complex128 if is_complex else np.float64
        self._check(self.lib.vlfs_system_init(self.ctx, self.ndofs, int(self.is_complex), nmat, nvec))
        self.nnz = self.ncoo = None
        self.domains, self._pattern = [], None
        return self

    def close(self):
        if not getattr(self, "owns_ctx", True):
            self.ctx = C.c_void_p()
            return
        if self.ctx:
            self.lib.vlfs_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=(0,)):
        if rc not in allow:
            msg = self.lib.vlfs_last_error(self.ctx)
            raise L.VlfsError("libvlfs_b200 status %d: %s" % (rc, msg.decode() if msg else ""))
        return rc

    # ------------------------------------------------------------------------------
    def add_domain(self, D, nq, qweights, slots, measure_slot=0, measure_kind=L.MEASURE_CELL,
                   weights=(), Ctensor=None, affine=None):
        """slots: list of dicts produced by fe.Measure.*_slot.  affine=None: detect whether
        every slot's cells have a constant Jacobian (enables the cheaper geometry stage and
        the reference-matrix fast path)."""
        if affine is None:
            affine = all(slot_is_affine(s["coords"], s["dref"], s["nv"]) for s in slots)
        d = L.DomainDesc()
        keep = []
        ncells = len(slots[0]["dofs"])
        d.D, d.nq, d.nslots = D, nq, len(slots)
        d.measure_slot, d.measure_kind = measure_slot, measure_kind
        d.nweights, d.affine, d.ncells = len(weights), int(bool(affine)), ncells
        qw = L.f64(qweights); keep.append(qw)
        d.qweights = L.dptr(qw)
        if len(weights):
            w = L.f64(np.stack([np.asarray(x, dtype=np.float64).reshape(ncells, nq) for x in weights]))
            keep.append(w)
            d.weights = L.dptr(w)
        if Ctensor is not None:
            ct = L.f64(Ctensor); keep.append(ct)
            assert ct.size == D ** 4
            d.C = L.dptr(ct)
        for k, s in enumerate(slots):
            sd = d.slot[k]
            sd.dref, sd.nv, sd.nvar, sd.ndofs = s["dref"], s["nv"], s["nvar"], s["ndofs"]
            arrs = {}
            for name in ("coords", "geo_grad", "val", "grad", "hess", "ref_normal", "ref_tangent"):
                arrs[name] = L.f64(s[name]) if s.get(name) is not None else None
            for name in ("variant", "dofs"):
                arrs[name] = L.i32(s[name]) if s.get(name) is not None else None
            assert arrs["coords"].shape == (ncells, sd.nv, D), (arrs["coords"].shape, ncells, sd.nv, D)
            assert arrs["dofs"].shape == (ncells, sd.ndofs)
            assert arrs["val"].shape == (sd.nvar, nq, sd.ndofs)
            assert arrs["grad"].shape == (sd.nvar, nq, sd.ndofs, sd.dref)
            assert arrs["geo_grad"].shape == (sd.nvar, nq, sd.nv, sd.dref)
            if arrs["hess"] is not None:
                assert arrs["hess"].shape == (sd.nvar, nq, sd.ndofs, sd.dref, sd.dref)
            keep.append(arrs)
            sd.coords, sd.geo_grad, sd.val = L.dptr(arrs["coords"]), L.dptr(arrs["geo_grad"]), L.dptr(arrs["val"])
            sd.grad, sd.hess = L.dptr(arrs["grad"]), L.dptr(arrs["hess"])
            sd.ref_normal, sd.ref_tangent = L.dptr(arrs["ref_normal"]), L.dptr(arrs["ref_tangent"])
            sd.variant, sd.dofs = L.iptr(arrs["variant"]), L.iptr(arrs["dofs"])
        did = C.c_int(-1)
        self._check(self.lib.vlfs_add_domain(self.ctx, C.byref(d), C.byref(did)))
        dom = Domain(self, did.value, len(slots), D)
        dom.ncells, dom.nq = ncells, nq
        dom.slot_ndofs = [int(s["ndofs"]) for s in slots]
        dom.table_bytes = ncells * (sum(4 * int(s["ndofs"]) + 8 * D * int(s["nv"]) for s in slots) + 8 * nq * len(weights))
        self.domains.append(dom)
        return dom

    # ------------------------------------------------------------------------------
    def build_pattern(self):
        import time
        nnz, ncoo = C.c_int64(0), C.c_int64(0)
        t0 = time.perf_counter()
        self._check(self.lib.vlfs_build_pattern(self.ctx, C.byref(nnz), C.byref(ncoo)))
        self.pattern_build_ms = (time.perf_counter() - t0) * 1e3     # symbolic phase, once per mesh
        self.nnz, self.ncoo = nnz.value, ncoo.value
        return self.nnz

    def get_pattern(self):
        if self._pattern is None:
            rp = np.empty(self.ndofs + 1, dtype=np.int32)
            ci = np.empty(self.nnz, dtype=np.int32)
            self._check(self.lib.vlfs_get_pattern(self.ctx, L.iptr(rp), L.iptr(ci)))
            self._pattern = (rp, ci)
        return self._pattern

    def get_pattern_csc(self):
        cp = np.empty(self.ndofs + 1, dtype=np.int32)
        rv = np.empty(self.nnz, dtype=np.int32)
        sl = np.empty(self.nnz, dtype=np.int32)
        self._check(self.lib.vlfs_get_pattern_csc(self.ctx, L.iptr(cp), L.iptr(rv), L.iptr(sl)))
        return cp, rv, sl

    def assemble(self, matrices=True, vectors=True):
        if matrices and vectors:
            self._check(self.lib.vlfs_assemble(self.ctx))
        elif matrices:
            self._check(self.lib.vlfs_assemble_matrices(self.ctx))
        elif vectors:
            self._check(self.lib.vlfs_assemble_vectors(self.ctx))

    def get_values(self, mat=0):
        v = np.empty(self.nnz, dtype=self.dtype)
        self._check(self.lib.vlfs_get_values(self.ctx, mat, v.ctypes.data_as(C.c_void_p)))
        return v

    def get_matrix(self, mat=0):
        """CSR matrix (explicit zeros kept), the `get_matrix(op)` of the reference."""
        rp, ci = self.get_pattern()
        return sp.csr_matrix((self.get_values(mat), ci.copy(), rp.copy()), shape=(self.ndofs, self.ndofs))

    def get_vector(self, vec=0):
        v = np.empty(self.ndofs, dtype=self.dtype)
        self._check(self.lib.vlfs_get_vector(self.ctx, vec, v.ctypes.data_as(C.c_void_p)))
        return v

    def combine(self, target, sources, coefs):
        src = L.i32(sources)
        cf = np.zeros(2 * len(sources))
        cf[0::2] = np.real(coefs)
        cf[1::2] = np.imag(coefs)
        self._check(self.lib.vlfs_combine(self.ctx, target, len(sources), L.iptr(src), L.dptr(cf)))

    def precond_apply(self, r):
        """z = M⁻¹ r with the preconditioner of the last `solver_setup` (`vlfs_precond_apply`)."""
        r = np.ascontiguousarray(r, dtype=self.dtype)
        z = np.zeros_like(r)
        self._check(self.lib.vlfs_precond_apply(self.ctx, r.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p)))
        return z

    def precond_info(self):
        out = np.zeros(8)
        self._check(self.lib.vlfs_precond_info(self.ctx, L.dptr(out)))
        d = dict(zip(("levels_lower", "levels_upper", "launches_lower", "launches_upper"), out[:4].astype(int).tolist()))
        d.update(symbolic_ms=out[4], numeric_ms=out[5], apply_ms=out[6], nnz=int(out[7]))
        return d

    def add_work_matrix(self):
        """One more value array on the pattern that no term targets (`vlfs_add_work_matrix`)."""
        m = C.c_int(-1)
        self._check(self.lib.vlfs_add_work_matrix(self.ctx, C.byref(m)))
        return m.value

    def spmv(self, x, mat=0):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        y = np.empty_like(x)
        self._check(self.lib.vlfs_spmv(self.ctx, mat, x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)))
        return y

    def set_dof_coords(self, coords):
        """DOF node coordinates (ndofs, dim) - ordering input of the sparse LU."""
        c = L.f64(coords)
        assert c.shape[0] == self.ndofs
        self._check(self.lib.vlfs_set_dof_coords(self.ctx, c.shape[1], L.dptr(c)))

    def set_owned_rows(self, n_owned):
        """Distributed systems: rows [0,n_owned) are owned, the rest are ghosts (dist.py)."""
        self._check(self.lib.vlfs_set_owned_rows(self.ctx, int(n_owned)))
        self.n_owned = int(n_owned)

    def refactor(self, mat=0):
        self.factor_status = self._check(self.lib.vlfs_solver_refactor(self.ctx, mat), allow=(0, L.PIVOT_PERTURBED))

    def direct_status(self):
        """(pivots replaced by the static perturbation, non-finite pivots) of the last factorisation."""
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.vlfs_direct_status(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    @classmethod
    def from_csr(cls, A, nmat=1, device=0):
        """The `LUSolver()` drop-in on a matrix the host already assembled (Gridap's
        `numerical_setup(ss, A)`): imports the CSR pattern and the values of a SciPy matrix."""
        A = sp.csr_matrix(A)
        A.sort_indices()
        is_complex = np.iscomplexobj(A.data)
        self = cls.__new__(cls)
        self.lib = L.load()
        self.ctx = C.c_void_p()
        rc = self.lib.vlfs_create(C.byref(self.ctx), device)
        if rc != 0:
            raise L.VlfsError("vlfs_create failed (status %d): no usable sm_100 GPU; the B200 "
                              "path has no CPU fallback" % rc)
        self.ndofs, self.is_complex, self.nmat, self.nvec = A.shape[0], bool(is_complex), nmat, 1
        self.dtype = np.complex128 if is_complex else np.float64
        rp, ci = L.i32(A.indptr), L.i32(A.indices)
        self._check(self.lib.vlfs_import_csr(self.ctx, self.ndofs, int(is_complex), nmat, L.iptr(rp), L.iptr(ci)))
        self.nnz = self.ncoo = int(rp[-1])
        self.domains, self._pattern = [], (rp, ci)
        self.set_values(A.data, 0)
        return self

    def set_values(self, values, mat=0):
        v = np.ascontiguousarray(values, dtype=self.dtype)
        assert v.shape == (self.nnz,)
        self._check(self.lib.vlfs_set_values(self.ctx, mat, v.ctypes.data_as(C.c_void_p)))

    def set_vector(self, values, vec=0):
        v = np.ascontiguousarray(values, dtype=self.dtype)
        assert v.shape == (self.ndofs,)
        self._check(self.lib.vlfs_set_vector(self.ctx, vec, v.ctypes.data_as(C.c_void_p)))

    def direct_info(self):
        out = np.zeros(8)
        self._check(self.lib.vlfs_direct_info(self.ctx, L.dptr(out)))
        keys = ["nfronts", "nlevels", "maxfront", "fill", "flops", "pool_bytes", "symbolic_ms", "numeric_ms"]
        return dict(zip(keys, out.tolist()))

    def solver_setup(self, mat=0, method=L.SOLVER_GMRES, precond=L.PRECOND_JACOBI, restart=100,
                     maxit=5000, rtol=1e-10, block_size_hint=0, sweeps=0):
        o = L.SolverOpts(method, precond, restart, maxit, rtol, block_size_hint, sweeps)
        # VLFS_PIVOT_PERTURBED is a status: the factorisation is usable, refinement decides
        self.factor_status = self._check(self.lib.vlfs_solver_setup(self.ctx, mat, C.byref(o)),
                                         allow=(0, L.PIVOT_PERTURBED))

    def solve(self, b=None, vec=0, x0=None):
        x = np.zeros(self.ndofs, dtype=self.dtype) if x0 is None else np.ascontiguousarray(x0, dtype=self.dtype).copy()
        st = L.SolveStats()
        if b is None:
            rc = self.lib.vlfs_solve_vec(self.ctx, vec, x.ctypes.data_as(C.c_void_p), C.byref(st))
        else:
            b = np.ascontiguousarray(b, dtype=self.dtype)
            rc = self.lib.vlfs_solve(self.ctx, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), C.byref(st))
        self._check(rc, allow=(0, L.NOT_CONVERGED))
        return x, dict(iterations=st.iterations, converged=bool(st.converged), rel_residual=st.rel_residual,
                       setup_ms=st.setup_ms, solve_ms=st.solve_ms, backward_error=st.backward_error,
                       at_attainable_accuracy=st.converged == 2)

    def newmark_run(self, matK, matC, matM, nsteps, dt, gamma, beta, tcoef, u0, v0, a0,
                    out_range=(0, 0), out_stride=1):
        tc = L.f64(tcoef).reshape(nsteps, -1) if tcoef is not None and np.size(tcoef) else np.zeros((nsteps, 0))
        nb = tc.shape[1]
        u, v, a = (np.ascontiguousarray(z, dtype=np.float64).copy() for z in (u0, v0, a0))
        lo, hi = out_range
        nout = nsteps // out_stride
        out = np.zeros((nout, hi - lo))
        st = L.SolveStats()
        rc = self.lib.vlfs_newmark_run(self.ctx, matK, matC, matM, nsteps, dt, gamma, beta, nb,
                                       L.dptr(tc) if nb else None, L.dptr(u), L.dptr(v), L.dptr(a),
                                       lo, hi, out_stride, L.dptr(out) if hi > lo else None, C.byref(st))
        self._check(rc, allow=(0, L.NOT_CONVERGED))
        return u, v, a, out, dict(iterations=st.iterations, converged=bool(st.converged),
                                  rel_residual=st.rel_residual, solve_ms=st.solve_ms)

    # device-resident timing hooks
    def spmv_device_ms(self, mat=0, reps=10):
        ms = C.c_float(0)
        self._check(self.lib.vlfs_spmv_device(self.ctx, mat, reps, C.byref(ms)))
        return ms.value / reps

    def assemble_device_ms(self, reps=1):
        ms = C.c_float(0)
        self._check(self.lib.vlfs_assemble_timed(self.ctx, reps, C.byref(ms)))
        return ms.value / reps

    def assemble_profile(self, reps=3):
        """[(kernel name, avg ms)] of one assembly pass, CUDA events on the launching stream."""
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 64)()
        n = C.c_int(0)
        self._check(self.lib.vlfs_assemble_profile(self.ctx, reps, names, 4096, ms, 64, C.byref(n)))
        nm = names.value.decode().split(";")[:n.value]
        return [(nm[k], float(ms[k])) for k in range(n.value)]

    def fp64_peak_tflops(self):
        """Measured FP64 FMA rate of this device (roofline denominator of the FP64-bound kernels)."""
        v = C.c_double(0.0)
        self._check(self.lib.vlfs_measure_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def launches(self):
        n = C.c_int64(0)
        self.lib.vlfs_last_kernel_launches(self.ctx, C.byref(n))
        return n.value
