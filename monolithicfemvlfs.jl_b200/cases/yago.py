"""B200 mirror of `run_Yago_freq_domain` (src/Yago_freq_domain.jl:22-209): same parameters,
same weak form, assembled and solved by libvlfs_b200 on the GPU."""
from dataclasses import dataclass
import numpy as np
from .. import _lib as L
from ..geometry import (CartesianDiscreteModel, add_tag_from_tags, Interior, Boundary,
                        Triangulation, Skeleton)
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System


@dataclass
class Yago_freq_domain_params:          # src/Yago_freq_domain.jl:11-20
    name: str = "YagoFreq"
    nx: int = 32
    ny: int = 4
    nz: int = 4
    order: int = 4
    lfactor: float = 0.4                # λfactor
    dfactor: float = 3
    vtk_output: bool = False


MAT_B0, MAT_B1, MAT_B2, MAT_BK = 0, 1, 2, 3     # basis matrices of the ω sweep; A(ω) is a work matrix (prob.MAT_A)


def wave_constants(L_, H, g, lfactor):
    k = 2 * np.pi / (L_ * lfactor)
    om = np.sqrt(g * k * np.tanh(k * H))
    return k, om


class YagoProblem:
    """Everything up to and including `op = AffineFEOperator(a,b,X,Y)`."""

    def __init__(self, params, device=0, system_cls=B200System, affine=None, basis=False):
        """affine=None: detect constant-Jacobian domains; False: force the general per-point
        geometry path everywhere (parity check of the fast paths).

        basis=True: ω-sweep by basis matrices.  ω enters the bilinear form only through
        the scalars ω, ω² and k (src/Yago_freq_domain.jl:69-94,162-165), so
        A(ω) = B₀ + ω B₁ + ω² B₂ + k B₃ with ω-independent B_k: the terms are registered on the matrices
        MAT_B0..MAT_BK of ONE union pattern with their ω-independent factors, `assemble_basis` integrates
        them once, and `combine_frequency` forms the work matrix `self.MAT_A` with one `vlfs_combine` per
        frequency."""
        p = self.params = params
        self.basis = bool(basis)
        nx, ny, nz, order = p.nx, p.ny, p.nz, p.order
        # fixed parameters (:29-50)
        Lb, B, H, hb = 300, 60, 58.5, 2.0
        nLO, nBO = 10, 18
        LO, BO = nLO * Lb, nBO * B
        xb0 = 4.5 * Lb; xb1 = xb0 + Lb; yb0 = -B / 2; yb1 = yb0 + B
        g, rho, E, rhob, nu = 9.81, 1025, 11.9e9, 256.25, 0.13
        I = hb ** 3 / 12
        D_ = E * I / (1.0 - nu ** 2)
        d0 = rhob * hb / rho
        mu_ = E / (2 * (1 + nu)); lam = nu * E / (1 - nu ** 2)
        Ct = np.zeros((3, 3, 3, 3))          # :52-66 (in-plane components only)
        dl = np.eye(2)
        for i in range(2):
            for j in range(2):
                for k in range(2):
                    for l in range(2):
                        Ct[i, j, k, l] = I / rho * (mu_ * (dl[i, k] * dl[j, l] + dl[i, l] * dl[j, k])
                                                    + lam * dl[i, j] * dl[k, l])
        self.c = dict(L=Lb, H=H, LO=LO, xb0=xb0, g=g, rho=rho, D=D_, d0=d0, eta0=0.01,
                      h=LO / (nLO * nx), bh=0.5, mu0=2.5, Ld=p.dfactor * Lb, Ld0=p.dfactor * Lb,
                      xd=LO - p.dfactor * Lb, xd0=p.dfactor * Lb)
        c = self.c

        def zmap(X):                         # f_z (:100-106), vectorised
            z = X[:, 2]
            i = z / (H / nz)
            zz = np.where(z == H, H, H - H / (2.5 ** i))
            return np.stack([X[:, 0], X[:, 1], zz], axis=1)

        model = CartesianDiscreteModel((0.0, LO, -BO / 2, BO / 2, 0.0, H), (nLO * nx, nBO * ny, nz), map=zmap)
        add_tag_from_tags(model, "surface", [22])
        add_tag_from_tags(model, "inlet", [25])
        self.model = model
        Om = Interior(model)
        G = Boundary(model, tags="surface")
        Gin = Boundary(model, tags="inlet")
        xG = G.cell_coords()
        is_plate = np.all((xb0 <= xG[:, :, 0]) & (xG[:, :, 0] <= xb1) & (yb0 <= xG[:, :, 1]) & (xG[:, :, 1] <= yb1), axis=1)
        Gb = Triangulation(G, np.nonzero(is_plate)[0])
        Gf = Triangulation(G, np.nonzero(~is_plate)[0])
        Lb_ = Skeleton(Gb)
        self.Om, self.G, self.Gin, self.Gb, self.Gf, self.Lb = Om, G, Gin, Gb, Gf, Lb_
        degree = 2 * order
        dOm, dGb, dGf, dGin, dLb = (Measure(t, degree) for t in (Om, Gb, Gf, Gin, Lb_))
        cplx = np.complex128
        V_Om = TestFESpace(Om, ReferenceFE(lagrangian, float, 2), vector_type=cplx)
        V_Gk = TestFESpace(Gf, ReferenceFE(lagrangian, float, order), vector_type=cplx)
        V_Ge = TestFESpace(Gb, ReferenceFE(lagrangian, float, order), vector_type=cplx)
        self.X = X = MultiFieldFESpace([TrialFESpace(V_Om), TrialFESpace(V_Gk), TrialFESpace(V_Ge)])
        self.V_Om, self.V_Gk, self.V_Ge = V_Om, V_Gk, V_Ge

        sys_ = self.sys = system_cls(X.ndofs, True, nmat=4 if self.basis else 1, nvec=1, device=device)
        ez = (0.0, 0.0, 1.0)
        # target matrix of a term by its ω-class: everything on matrix 0 unless `basis`
        M0, M1, M2, MK = (MAT_B0, MAT_B1, MAT_B2, MAT_BK) if self.basis else (0, 0, 0, 0)
        bh = c["bh"]
        fb = self._basis_factors = self.basis_factors(bh, g, c["d0"])
        fac = (lambda name: fb[name]) if self.basis else (lambda name: 0)
        # ---- Ω : ∫ ∇(ϕ)⋅∇(w) (:162)
        s_phi = dOm.own_slot(V_Om)
        self.dom_Om = sys_.add_domain(3, dOm.nq, dOm.wq, [s_phi], affine=affine)
        self.dom_Om.add_term(M0, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)
        # ---- Γf (:163): slot 0 = ϕ trace, slot 1 = κ ; weights: 0 = μ₁, 1 = Re/Im of rhs fields
        xq = dGf.points()
        self._xq_Gf = xq[..., 0].copy()
        self.dom_Gf = sys_.add_domain(3, dGf.nq, dGf.wq, [dGf.trace_slot(V_Om), dGf.surface_slot(V_Gk)],
                                      measure_slot=1, weights=[np.zeros_like(self._xq_Gf)] * 5, affine=affine)
        d = self.dom_Gf
        PHI, KAP = 0, 1
        self.t_Gf = dict(
            uk=d.add_term(M0, fac("uk"), L.OP_VAL, KAP, L.OP_VAL, KAP),                # βₕ g κ u
            up=d.add_term(M1, fac("up"), L.OP_VAL, KAP, L.OP_VAL, PHI),                # -iωβₕ ϕ u
            upn=d.add_term(M0, 1.0, L.OP_VAL, KAP, L.OP_DDIR, PHI, weight=0, trial_dir=ez),   # μ₁ ∇ₙϕ u
            wk=d.add_term(M1, fac("wk"), L.OP_VAL, PHI, L.OP_VAL, KAP),                # (βₕ g αₕ + iω) κ w
            wkm=d.add_term(MK, fac("wkm"), L.OP_VAL, PHI, L.OP_VAL, KAP, weight=0),    # -k μ₁ κ w
            wp=d.add_term(M2, fac("wp"), L.OP_VAL, PHI, L.OP_VAL, PHI),                # -iωβₕαₕ ϕ w
            wpn=d.add_term(M1, fac("wpn"), L.OP_VAL, PHI, L.OP_DDIR, PHI, weight=0, trial_dir=ez))   # αₕ μ₁ ∇ₙϕ w
        # rhs (:166): -∫(ηd w - ∇ₙϕd (u + αₕ w)); complex fields sampled at the points
        self.r_Gf = dict(
            w1=d.add_rhs(0, -1.0, L.OP_VAL, PHI, weight_re=1, weight_im=2),           # -ηd w
            w2=d.add_rhs(0, 0, L.OP_VAL, PHI, weight_re=3, weight_im=4),              # αₕ ∇ₙϕd w
            u=d.add_rhs(0, 1.0, L.OP_VAL, KAP, weight_re=3, weight_im=4))             # ∇ₙϕd u
        # ---- Γb (:164): slot 0 = ϕ trace, slot 1 = η
        self.dom_Gb = d = sys_.add_domain(3, dGb.nq, dGb.wq,
                                          [dGb.trace_slot(V_Om), dGb.surface_slot(V_Ge, need_hess=True)],
                                          measure_slot=1, Ctensor=Ct, affine=affine)
        ETA = 1
        self.t_Gb = dict(
            vv=d.add_term(M0, fac("vv0"), L.OP_VAL, ETA, L.OP_VAL, ETA),               # (g-ω²d₀) η v
            bend=d.add_term(M0, 1.0, L.OP_HESS, ETA, L.OP_CHESS, ETA),                # ∇∇v⊙(C⊙∇∇η)
            vp=d.add_term(M1, fac("vp"), L.OP_VAL, ETA, L.OP_VAL, PHI),                # -iω ϕ v
            we=d.add_term(M1, fac("we"), L.OP_VAL, PHI, L.OP_VAL, ETA))                # iω η w
        if self.basis:
            d.add_term(M2, fb["vv2"], L.OP_VAL, ETA, L.OP_VAL, ETA)                   # -ω²d₀ η v
        # ---- Λb (:165): slot 0 = η⁺, slot 1 = η⁻
        self.dom_Lb = d = sys_.add_domain(3, dLb.nq, dLb.wq,
                                          [dLb.skeleton_slot(V_Ge, "plus"), dLb.skeleton_slot(V_Ge, "minus")],
                                          measure_slot=0, measure_kind=L.MEASURE_FACET, Ctensor=Ct, affine=affine)
        jump, mean = {0: 1.0, 1: -1.0}, {0: 0.5, 1: 0.5}
        gamma = 1.0 * order * (order + 1) / c["h"]
        d.add_term(M0, -1.0, L.OP_GRAD, jump, L.OP_CHESS_NPLUS, mean)
        d.add_term(M0, -1.0, L.OP_CHESS_NPLUS, mean, L.OP_GRAD, jump)
        d.add_term(M0, D_ / rho * gamma, L.OP_GRAD, jump, L.OP_GRAD, jump)
        # ---- Γin rhs (:166): ∫ vin w ; slot 0 = ϕ trace, slot 1 = Q1 face (measure only)
        V_aux = TestFESpace(Gin, ReferenceFE(lagrangian, float, 1))
        V_aux.offset = 0
        self._xq_Gin = dGin.points()
        self.dom_Gin = d = sys_.add_domain(3, dGin.nq, dGin.wq, [dGin.trace_slot(V_Om), dGin.surface_slot(V_aux)],
                                           measure_slot=1, weights=[np.zeros(self._xq_Gin.shape[:2])] * 2, affine=affine)
        d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=0, weight_im=1)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords()[:, :2])       # quasi-2-D: dissect in the (x,y)
        self.nnz = sys_.build_pattern()
        self.MAT_A = sys_.add_work_matrix() if self.basis and hasattr(sys_, "add_work_matrix") else (4 if self.basis else 0)
        self.set_frequency(p.lfactor)

    @staticmethod
    def basis_factors(bh, g, d0):
        """ω-independent factors of the ω-dependent terms: with αₕ = -iω/g (1-βₕ)/βₕ
        (src/Yago_freq_domain.jl:89-94) every coefficient of `frequency_inputs` is one of
        {1, ω, ω², k} times the number below."""
        a1 = -1j / g * (1 - bh) / bh                       # αₕ / ω
        return dict(uk=bh * g, up=-1j * bh, wk=bh * g * a1 + 1j, wkm=-1.0, wp=-1j * bh * a1,
                    wpn=a1, vv0=g, vv2=-d0, vp=-1j, we=1j)

    def assemble_basis(self):
        """Integrates B₀..B₃ (and the right-hand side of the current frequency) once."""
        assert self.basis
        self.sys.assemble()

    def combine_frequency(self, lfactor=None):
        """A(ω) = B₀ + ω B₁ + ω² B₂ + k B₃ into the work matrix `MAT_A` (`vlfs_combine`); the right-hand side still
        comes from `set_frequency` + `assemble(matrices=False)`: it holds ω inside exp(ikx) fields."""
        assert self.basis
        if lfactor is not None:
            self.set_frequency(lfactor)
        self.sys.combine(self.MAT_A, [MAT_B0, MAT_B1, MAT_B2, MAT_BK], [1.0, self.om, self.om ** 2, self.k])

    # ω-dependent scalars and coefficient fields only (the sweep of scripts/5-4-1-Yago.jl:33-71)
    def frequency_inputs(self, lfactor):
        """Host evaluation of everything that depends on ω (scripts/5-4-1-Yago.jl:33-71): the
        coefficient fields at the quadrature points and the scalar coefficients of the terms."""
        c = self.c
        g, H = c["g"], c["H"]
        k, om = wave_constants(c["L"], H, g, lfactor)
        bh = c["bh"]
        ah = -1j * om / g * (1 - bh) / bh
        x1 = self._xq_Gf
        mu1 = (c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - c["xd"]) / c["Ld"])) * (x1 > c["xd"])
               + c["mu0"] * (1 - np.cos(np.pi / 2 * (c["Ld0"] - x1) / c["Ld0"])) * (x1 < c["xd0"]))
        eta_in = c["eta0"] * np.exp(1j * k * x1)
        vz_in = -1j * om * c["eta0"] * np.exp(1j * k * x1)
        eta_d = mu1 * k * eta_in * (x1 < c["xd0"])
        dphi_d = mu1 * vz_in * (x1 < c["xd0"])
        x = self._xq_Gin
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 2]) / np.sinh(k * H)) * np.exp(1j * k * x[..., 0])
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        t, tb = self.t_Gf, self.t_Gb
        return dict(
            k=k, om=om,
            weights_Gf=[f64(mu1), f64(eta_d.real), f64(eta_d.imag), f64(dphi_d.real), f64(dphi_d.imag)],
            weights_Gin=[f64(vin.real), f64(vin.imag)],
            coef_Gf={t["uk"]: bh * g, t["up"]: -1j * om * bh, t["wk"]: bh * g * ah + 1j * om, t["wkm"]: -k,
                     t["wp"]: -1j * om * bh * ah, t["wpn"]: ah},
            rhs_Gf={self.r_Gf["w2"]: ah},
            coef_Gb={tb["vv"]: g - om ** 2 * c["d0"], tb["vp"]: -1j * om, tb["we"]: 1j * om})

    def apply_frequency_inputs(self, inp):
        """Host → device: `vlfs_set_weights` / `vlfs_set_term_coef` / `vlfs_set_rhs_coef`."""
        self.k, self.om = inp["k"], inp["om"]
        for i, w in enumerate(inp["weights_Gf"]):
            if not (self.basis and i == 0 and getattr(self, "_mu1_set", False)):
                self.dom_Gf.set_weights(i, w)       # μ₁ (field 0) does not depend on ω
        self._mu1_set = True
        if not self.basis:                          # basis mode: the factors of B_k are fixed
            for tid, v in inp["coef_Gf"].items():
                self.dom_Gf.set_coef(tid, v)
            for tid, v in inp["coef_Gb"].items():
                self.dom_Gb.set_coef(tid, v)
        for rid, v in inp["rhs_Gf"].items():
            self.dom_Gf.set_rhs_coef(rid, v)
        for i, w in enumerate(inp["weights_Gin"]):
            self.dom_Gin.set_weights(i, w)

    def set_frequency(self, lfactor):
        self.apply_frequency_inputs(self.frequency_inputs(lfactor))

    def assemble(self):
        self.sys.assemble()

    def rao(self, x):
        """|η|/η₀ at plate-cell vertices on y = 0 (:178-199)."""
        c = self.c
        eta = self.X.split(x)[2]
        xe = self.Gb.cell_coords()
        s = -2 * (xe[:, :, 0] - (c["xb0"] + c["L"] / 2)) / c["L"]
        sel = (-1 <= s) & (s <= 1) & (xe[:, :, 1] == 0.0)
        ie, iv = np.nonzero(sel)
        return s[ie, iv], np.abs(eta[self.V_Ge.cell_dofs[ie, iv]]) / c["eta0"]


def run_Yago_freq_domain(params, device=0, solver=None):
    prob = YagoProblem(params, device=device)
    prob.assemble()
    opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    opts.update(solver or {})
    prob.sys.solver_setup(0, **opts)
    x, stats = prob.sys.solve(vec=0)
    prob.stats = stats
    xps, eta = prob.rao(x)
    if params.vtk_output:                                   # src/Yago_freq_domain.jl:201-205
        from ..vtk import writevtk
        phi, kap, et = prob.X.split(x)
        name = getattr(params, "name", "Yago")
        writevtk(prob.Om, name + "_O", [("phi_re", prob.V_Om, phi.real), ("phi_im", prob.V_Om, phi.imag)], nsubcells=10)
        writevtk(prob.Gf, name + "_Gk", [("eta_re", prob.V_Gk, kap.real), ("eta_im", prob.V_Gk, kap.imag)], nsubcells=10)
        writevtk(prob.Gb, name + "_Ge", [("eta_re", prob.V_Ge, et.real), ("eta_im", prob.V_Ge, et.imag)], nsubcells=10)
    return xps, eta
