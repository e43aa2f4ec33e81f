"""NumPy interpreter of the C-ABI descriptors (test infrastructure).

`RecordingSystem` has the same surface as the product's `B200System` but only records the
domains/terms a case module registers; `evaluate` then integrates them on the CPU with the
semantics documented in include/vlfs_b200.h.  It lets the `-m "not gpu"` tests check the
host-side marshalling (tables, slots, operators, scales, coefficient fields) against the
oracle's literal weak forms without a GPU.  It is NOT a fallback of the product: nothing
in `monolithicfemvlfs.jl_b200/` imports it.
"""
import numpy as np
from .assembly import COO

OP_VAL, OP_GRAD, OP_DDIR, OP_LAPL, OP_HESS, OP_CHESS, OP_GRAD_NOWN, OP_CHESS_NPLUS = range(8)
MEASURE_CELL, MEASURE_FACET = 0, 1


class _Dom:
    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.terms, self.rhs = [], []

    @staticmethod
    def _sc(spec, n):
        out = np.zeros(n)
        if isinstance(spec, dict):
            for k, v in spec.items():
                out[k] = v
        else:
            out[int(spec)] = 1.0
        return out

    def add_term(self, mat, coef, test_op, test_slots, trial_op, trial_slots, weight=-1,
                 test_dir=(0, 0, 0), trial_dir=(0, 0, 0)):
        n = len(self.slots)
        self.terms.append(dict(mat=mat, coef=complex(coef), test_op=test_op, trial_op=trial_op,
                               st=self._sc(test_slots, n), su=self._sc(trial_slots, n), weight=weight,
                               test_dir=np.array(test_dir, float), trial_dir=np.array(trial_dir, float)))
        return len(self.terms) - 1

    def add_rhs(self, vec, coef, op, slots, weight_re=-1, weight_im=-1, direction=(0, 0, 0)):
        self.rhs.append(dict(vec=vec, coef=complex(coef), op=op, sc=self._sc(slots, len(self.slots)),
                             wre=weight_re, wim=weight_im, dir=np.array(direction, float)))
        return len(self.rhs) - 1

    def set_coef(self, term, coef):
        self.terms[term]["coef"] = complex(coef)

    def set_rhs_coef(self, term, coef):
        self.rhs[term]["coef"] = complex(coef)

    def set_weights(self, index, values):
        self.weights[index] = np.asarray(values, dtype=float).reshape(self.ncells, self.nq)


class RecordingSystem:
    def __init__(self, ndofs, is_complex, nmat=1, nvec=1, device=0):
        self.ndofs, self.is_complex, self.nmat, self.nvec = ndofs, is_complex, nmat, nvec
        self.dtype = np.complex128 if is_complex else np.float64
        self.domains = []

    def add_domain(self, D, nq, qweights, slots, measure_slot=0, measure_kind=MEASURE_CELL,
                   weights=(), Ctensor=None, affine=None):
        ncells = len(slots[0]["dofs"])
        d = _Dom(D=D, nq=nq, qweights=np.asarray(qweights, float), slots=slots, measure_slot=measure_slot,
                 measure_kind=measure_kind, ncells=ncells, affine=affine,
                 weights=[np.asarray(w, float).reshape(ncells, nq).copy() for w in weights],
                 C=None if Ctensor is None else np.asarray(Ctensor, float).reshape((D,) * 4))
        self.domains.append(d)
        return d

    def build_pattern(self):
        return None

    # ---- evaluation -------------------------------------------------------------
    def evaluate(self):
        mats = [COO(self.ndofs, self.dtype) for _ in range(self.nmat)]
        vecs = [np.zeros(self.ndofs, dtype=self.dtype) for _ in range(self.nvec)]
        for d in self.domains:
            _eval_domain(d, mats, vecs, self.is_complex)
        return [m.tocsr() for m in mats], vecs


def _slot_geometry(d, s):
    D, nq = d.D, d.nq
    X = np.asarray(s["coords"], float)
    var = np.zeros(d.ncells, dtype=np.int64) if s["variant"] is None else np.asarray(s["variant"])
    gg = np.asarray(s["geo_grad"], float)[var]                      # (e,q,v,a)
    Jt = np.einsum("eqva,evb->eqab", gg, X)
    dref = s["dref"]
    if dref == D:
        pinv = np.linalg.inv(Jt)
        meas = np.abs(np.linalg.det(Jt))
    elif dref == 0:
        pinv = np.zeros((d.ncells, nq, D, 0))
        meas = np.ones((d.ncells, nq))
    else:
        G = Jt @ np.swapaxes(Jt, -1, -2)
        pinv = np.swapaxes(Jt, -1, -2) @ np.linalg.inv(G)
        meas = np.sqrt(np.linalg.det(G))
    normal = None
    if s.get("ref_normal") is not None:
        n = np.einsum("eqba,ea->eqb", pinv, np.asarray(s["ref_normal"], float)[var])
        normal = n / np.linalg.norm(n, axis=-1, keepdims=True)
    fmeas = None
    if s.get("ref_tangent") is not None:
        t = np.einsum("ea,eqab->eqb", np.asarray(s["ref_tangent"], float)[var], Jt)
        fmeas = np.linalg.norm(t, axis=-1)
    return dict(pinv=pinv, meas=meas, normal=normal, fmeas=fmeas, var=var)


def _sym_comps(D):
    return [(a, a) for a in range(D)] + [(a, b) for a in range(D) for b in range(a + 1, D)]


def _op(d, geos, k, op, direction):
    """(e, q, ncomp, ndofs) operator rows of slot k."""
    s, g = d.slots[k], geos[k]
    var = g["var"]
    if op == OP_VAL:
        return np.asarray(s["val"], float)[var][:, :, None, :]
    grad = np.einsum("eqba,eqia->eqbi", g["pinv"], np.asarray(s["grad"], float)[var])
    if op == OP_GRAD:
        return grad
    if op == OP_DDIR:
        return np.einsum("eqbi,b->eqi", grad, direction[:d.D])[:, :, None, :]
    if op == OP_GRAD_NOWN:
        return np.einsum("eqbi,eqb->eqi", grad, g["normal"])[:, :, None, :]
    H = np.einsum("eqba,eqiac,eqdc->eqibd", g["pinv"], np.asarray(s["hess"], float)[var], g["pinv"])
    if op == OP_LAPL:
        return np.einsum("eqibb->eqi", H)[:, :, None, :]
    comps = _sym_comps(d.D)
    if op == OP_HESS:
        return np.stack([H[:, :, :, a, b] for a, b in comps], axis=2)
    CH = np.einsum("abcd,eqicd->eqiab", d.C, H)
    if op == OP_CHESS:
        return np.stack([(1.0 if a == b else 2.0) * CH[:, :, :, a, b] for a, b in comps], axis=2)
    if op == OP_CHESS_NPLUS:
        return np.einsum("eqiab,eqb->eqai", CH, geos[0]["normal"])
    raise ValueError(op)


def _eval_domain(d, mats, vecs, is_complex):
    if d.ncells == 0:
        return
    geos = [_slot_geometry(d, s) for s in d.slots]
    gm = geos[d.measure_slot]
    if d.measure_kind == MEASURE_FACET:
        meas = gm["fmeas"] if gm["fmeas"] is not None else np.ones((d.ncells, d.nq))
    else:
        meas = gm["meas"]
    dV = meas * d.qweights[None, :]
    ns = len(d.slots)
    touched = np.zeros((ns, ns), dtype=bool)
    for t in d.terms:
        touched |= np.outer(t["st"] != 0, t["su"] != 0)
    for a in range(ns):
        for b in range(ns):
            if not touched[a, b]:
                continue
            rows, cols = np.asarray(d.slots[a]["dofs"]), np.asarray(d.slots[b]["dofs"])
            blk = [np.zeros((d.ncells, rows.shape[1], cols.shape[1]), dtype=complex) for _ in mats]
            for t in d.terms:
                if t["st"][a] == 0 or t["su"][b] == 0:
                    continue
                A = _op(d, geos, a, t["test_op"], t["test_dir"])
                B = _op(d, geos, b, t["trial_op"], t["trial_dir"])
                w = dV if t["weight"] < 0 else dV * d.weights[t["weight"]]
                blk[t["mat"]] += t["coef"] * t["st"][a] * t["su"][b] * np.einsum("eq,eqci,eqcj->eij", w, A, B)
            for m, M in enumerate(mats):
                # every matrix shares the union pattern: touched blocks are stored in all of them
                M.add_block(rows, cols, blk[m] if is_complex else blk[m].real)
    for r in d.rhs:
        for a in range(ns):
            if r["sc"][a] == 0:
                continue
            A = _op(d, geos, a, r["op"], r["dir"])[:, :, 0, :]
            w = dV * ((d.weights[r["wre"]] if r["wre"] >= 0 else 1.0)
                      + 1j * (d.weights[r["wim"]] if r["wim"] >= 0 else 0.0))
            v = r["coef"] * r["sc"][a] * np.einsum("eq,eqi->ei", w, A)
            np.add.at(vecs[r["vec"]], np.asarray(d.slots[a]["dofs"]).reshape(-1),
                      (v if is_complex else v.real).reshape(-1))
