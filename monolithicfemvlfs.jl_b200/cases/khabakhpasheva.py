"""B200 mirrors of `run_Khabakhpasheva_freq_domain` (src/Khabakhpasheva_freq_domain.jl:22-261)
and `run_Khabakhpasheva_time_domain` (src/Khabakhpasheva_time_domain.jl:23-306): same
parameters and weak forms, assembled / solved / time-stepped by libvlfs_b200 on the GPU."""
from dataclasses import dataclass
import numpy as np
from .. import _lib as L
from ..geometry import (CartesianDiscreteModel, add_tag_from_tags, Interior, Boundary,
                        Triangulation, Skeleton)
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System
from .periodic_beam import nsteps_for

MAT_K, MAT_C, MAT_M, MAT_A = 0, 1, 2, 3


@dataclass
class Khabakhpasheva_freq_domain_params:      # src/Khabakhpasheva_freq_domain.jl:13-20
    name: str = "KhabakhpashevaFreqDomain"
    nx: int = 20
    ny: int = 5
    order: int = 4
    xi: float = 0.0                           # ξ
    vtk_output: bool = False


@dataclass
class Khabakhpasheva_time_domain_params:      # src/Khabakhpasheva_time_domain.jl:14-21
    name: str = "KhabakhpashevaTimeDomain"
    nx: int = 20
    ny: int = 5
    order: int = 4
    xi: float = 0.0
    vtk_output: bool = False


def isapprox(a, b):
    """Julia `a ≈ b` for Float64."""
    return np.abs(a - b) <= np.sqrt(np.finfo(float).eps) * np.maximum(np.abs(a), np.abs(b))


class _Tank:
    """Mesh, triangulations, FE spaces and measures common to both drivers (:28-224)."""

    def __init__(self, p):
        nx, ny, order, xi = p.nx, p.ny, p.order, p.xi
        Lb, m, EI1, EI2, beta, H, alpha = 12.5, 8.36, 47100.0, 471.0, 0.2, 1.1, 0.249
        Ld = Lb
        LO = 2 * Ld + 2 * Lb
        x0 = 0.0
        xd_in = x0 + Ld
        xb0 = xd_in + Lb / 2
        xbj = xb0 + beta * Lb
        xb1 = xb0 + Lb
        xd_out = LO - Ld
        g, rho = 9.81, 1025
        lam = alpha * Lb
        k = 2 * np.pi / lam
        om = np.sqrt(g * k * np.tanh(k * H))
        nx_total = int(np.ceil(nx / beta) * np.ceil(LO / Lb))
        h = LO / nx_total
        self.c = c = dict(Lb=Lb, H=H, Ld=Ld, LO=LO, xb0=xb0, xbj=xbj, xb1=xb1, xd_in=xd_in, xd_out=xd_out,
                          g=g, d0=m / rho, a1=EI1 / rho, a2=EI2 / rho, kr=xi * (EI1 / rho) / Lb, k=k, om=om,
                          T=2 * np.pi / om, eta0=0.01, h=h, gamma=1.0 * order * (order - 1) / h, bh=0.5,
                          mu0=2.5)

        def fmap(X):                               # f_y (:96-102), vectorised
            y = X[:, 1]
            i = y / (H / ny)
            return np.stack([X[:, 0], np.where(y == H, H, H - H / (2.5 ** i))], axis=1)

        model = CartesianDiscreteModel((x0, LO, 0.0, H), (nx_total, ny), map=fmap)
        add_tag_from_tags(model, "surface", [3, 4, 6])
        add_tag_from_tags(model, "bottom", [1, 2, 5])
        add_tag_from_tags(model, "inlet", [7])
        add_tag_from_tags(model, "outlet", [8])
        add_tag_from_tags(model, "water", [9])
        self.model = model
        self.Om = Interior(model)
        self.G = G = Boundary(model, tags="surface")
        self.Gin = Boundary(model, tags="inlet")
        xc = G.cell_coords().mean(axis=1)
        top = isapprox(xc[:, 1], H)
        b1 = (xb0 <= xc[:, 0]) & (xc[:, 0] <= xbj) & top
        b2 = (xbj <= xc[:, 0]) & (xc[:, 0] <= xb1) & top
        d1 = (x0 <= xc[:, 0]) & (xc[:, 0] <= xd_in) & top
        d2 = (xd_out <= xc[:, 0]) & top
        ids = lambda mask: np.nonzero(mask)[0]
        self.Gb1, self.Gb2 = Triangulation(G, ids(b1)), Triangulation(G, ids(b2))
        self.Gd1, self.Gd2 = Triangulation(G, ids(d1)), Triangulation(G, ids(d2))
        self.Gfs = Triangulation(G, ids(~(b1 | b2 | d1 | d2)))
        self.Geta = Triangulation(G, ids(b1 | b2))
        self.Gkap = Triangulation(G, ids(~(b1 | b2)))
        self.Lb1, self.Lb2 = Skeleton(self.Gb1), Skeleton(self.Gb2)
        act = G.active_model()
        xv = model.node_coords[act["parent"][0]]
        self.Lj = Skeleton(G, isapprox(xv[:, 0], xbj) & isapprox(xv[:, 1], H))
        self.degree = degree = 2 * order
        self.dOm = Measure(self.Om, degree)
        self.dG = {n: Measure(getattr(self, n), degree) for n in ("Gb1", "Gb2", "Gd1", "Gd2", "Gfs", "Gin")}
        self.dL = {n: Measure(getattr(self, n), degree) for n in ("Lb1", "Lb2", "Lj")}

    def spaces(self, vector_type):
        reffe = ReferenceFE(lagrangian, float, self.order)
        self.V_Om = TestFESpace(self.Om, reffe, conformity="H1", vector_type=vector_type)
        self.V_Gk = TestFESpace(self.Gkap, reffe, conformity="H1", vector_type=vector_type)
        self.V_Ge = TestFESpace(self.Geta, reffe, conformity="H1", vector_type=vector_type)
        self.X = MultiFieldFESpace([TrialFESpace(self.V_Om), TrialFESpace(self.V_Gk), TrialFESpace(self.V_Ge)])

    def mu1_in(self, x1):
        return self.c["mu0"] * (1.0 - np.sin(np.pi / 2 * x1 / self.c["Ld"]))

    def mu1_out(self, x1):
        return self.c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - self.c["xd_out"]) / self.c["Ld"]))

    def eta_nodes(self):
        """(xs, cell dof table, permutation) of the post-processing (:250-257)."""
        xc = self.V_Ge.cell_dof_coords()[:, :, 0]
        p = np.argsort(xc[0], kind="stable")
        return xc[:, p], self.V_Ge.cell_dofs[:, p]

    def _skeleton_domains(self, sys_, mat):
        """Λb1, Λb2 (:234-235) and Λj (:236); slot 0 = η⁺, slot 1 = η⁻."""
        c = self.c
        jump_n, mean = {0: 1.0, 1: 1.0}, {0: 0.5, 1: 0.5}
        for name, ai in (("Lb1", c["a1"]), ("Lb2", c["a2"])):
            dL = self.dL[name]
            d = sys_.add_domain(2, dL.nq, dL.wq, [dL.skeleton_slot(self.V_Ge, "plus"),
                                                  dL.skeleton_slot(self.V_Ge, "minus")],
                                measure_slot=0, measure_kind=L.MEASURE_FACET)
            d.add_term(mat, -ai, L.OP_GRAD_NOWN, jump_n, L.OP_LAPL, mean)
            d.add_term(mat, -ai, L.OP_LAPL, mean, L.OP_GRAD_NOWN, jump_n)
            d.add_term(mat, ai * c["gamma"], L.OP_GRAD_NOWN, jump_n, L.OP_GRAD_NOWN, jump_n)
        dL = self.dL["Lj"]
        d = sys_.add_domain(2, dL.nq, dL.wq, [dL.skeleton_slot(self.V_Ge, "plus", need_hess=False),
                                              dL.skeleton_slot(self.V_Ge, "minus", need_hess=False)],
                            measure_slot=0, measure_kind=L.MEASURE_FACET)
        d.add_term(mat, c["kr"], L.OP_GRAD_NOWN, jump_n, L.OP_GRAD_NOWN, jump_n)


class KhabakhpashevaFreqProblem(_Tank):
    """Everything up to `op = AffineFEOperator(a,l,X,Y)` (:226-240)."""

    def __init__(self, params, device=0, system_cls=B200System):
        self.params, self.order = params, params.order
        super().__init__(params)
        self.spaces(np.complex128)
        c = self.c
        g, om, bh, k = c["g"], c["om"], c["bh"], c["k"]
        ah = -1j * om / g * (1 - bh) / bh
        X = self.X
        sys_ = self.sys = system_cls(X.ndofs, True, nmat=1, nvec=1, device=device)
        ey = (0.0, 1.0, 0.0)
        d = sys_.add_domain(2, self.dOm.nq, self.dOm.wq, [self.dOm.own_slot(self.V_Om)])
        d.add_term(0, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)                       # ∇(w)⋅∇(ϕ)
        PHI, KAP = 0, 1
        for name, mu1f in (("Gfs", None), ("Gd1", self.mu1_in), ("Gd2", self.mu1_out)):
            dG = self.dG[name]
            if dG.ncells == 0:
                continue
            x1 = dG.points()[..., 0]
            fields = [mu1f(x1)] if mu1f is not None else []
            if name == "Gd1":                                                 # rhs fields (:237)
                m1 = fields[0]
                eta_d = m1 * k * c["eta0"] * np.exp(1j * k * x1)
                dphi_d = m1 * (-1j * om * c["eta0"] * np.exp(1j * k * x1))
                fields += [eta_d.real, eta_d.imag, dphi_d.real, dphi_d.imag]
            d = sys_.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(self.V_Om), dG.surface_slot(self.V_Gk)],
                                measure_slot=1, weights=fields)
            d.add_term(0, bh * g, L.OP_VAL, KAP, L.OP_VAL, KAP)               # βₕ g κ u
            d.add_term(0, -1j * om * bh, L.OP_VAL, KAP, L.OP_VAL, PHI)        # -iω βₕ ϕ u
            d.add_term(0, bh * g * ah + 1j * om, L.OP_VAL, PHI, L.OP_VAL, KAP)   # (βₕ g αₕ + iω) κ w
            d.add_term(0, -1j * om * bh * ah, L.OP_VAL, PHI, L.OP_VAL, PHI)   # -iω βₕ αₕ ϕ w
            if mu1f is not None:
                d.add_term(0, -k, L.OP_VAL, PHI, L.OP_VAL, KAP, weight=0)     # -μ₂ κ w
                d.add_term(0, 1.0, L.OP_VAL, KAP, L.OP_DDIR, PHI, weight=0, trial_dir=ey)   # μ₁ ∇ₙϕ u
                d.add_term(0, ah, L.OP_VAL, PHI, L.OP_DDIR, PHI, weight=0, trial_dir=ey)    # αₕ μ₁ ∇ₙϕ w
            if name == "Gd1":
                d.add_rhs(0, -1.0, L.OP_VAL, PHI, weight_re=1, weight_im=2)   # -ηd w
                d.add_rhs(0, ah, L.OP_VAL, PHI, weight_re=3, weight_im=4)     # αₕ ∇ₙϕd w
                d.add_rhs(0, 1.0, L.OP_VAL, KAP, weight_re=3, weight_im=4)    # ∇ₙϕd u
        ETA = 1
        for name, ai in (("Gb1", c["a1"]), ("Gb2", c["a2"])):
            dG = self.dG[name]
            d = sys_.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(self.V_Om),
                                                  dG.surface_slot(self.V_Ge, need_hess=True)], measure_slot=1)
            d.add_term(0, -om ** 2 * c["d0"] + g, L.OP_VAL, ETA, L.OP_VAL, ETA)
            d.add_term(0, ai, L.OP_LAPL, ETA, L.OP_LAPL, ETA)
            d.add_term(0, -1j * om, L.OP_VAL, ETA, L.OP_VAL, PHI)
            d.add_term(0, 1j * om, L.OP_VAL, PHI, L.OP_VAL, ETA)
        self._skeleton_domains(sys_, 0)
        dGin = self.dG["Gin"]
        V_aux = TestFESpace(self.Gin, ReferenceFE(lagrangian, float, 1))
        V_aux.offset = 0
        x = dGin.points()
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 1]) / np.sinh(k * c["H"])) * np.exp(1j * k * x[..., 0])
        d = sys_.add_domain(2, dGin.nq, dGin.wq, [dGin.trace_slot(self.V_Om), dGin.surface_slot(V_aux)],
                            measure_slot=1, weights=[vin.real, vin.imag])
        d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=0, weight_im=1)            # w vᵢₙ
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    def assemble(self):
        self.sys.assemble()

    def postprocess(self, x):
        xs, cd = self.eta_nodes()
        eta = self.X.split(x)[2]
        return ((xs - self.c["xb0"]) / self.c["Lb"]).reshape(-1), (np.abs(eta[cd]) / self.c["eta0"]).reshape(-1)


def run_Khabakhpasheva_freq_domain(params, device=0, solver=None):
    prob = KhabakhpashevaFreqProblem(params, device=device)
    prob.assemble()
    opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    opts.update(solver or {})
    prob.sys.solver_setup(0, **opts)
    x, prob.stats = prob.sys.solve(vec=0)
    if params.vtk_output:                                   # src/Khabakhpasheva_freq_domain.jl:244-246
        from ..vtk import writevtk
        phi, kap, et = prob.X.split(x)
        writevtk(prob.Om, params.name + "_O_solution", [("phi_re", prob.V_Om, phi.real), ("phi_im", prob.V_Om, phi.imag)], nsubcells=10)
        writevtk(prob.Gkap, params.name + "_Gk_solution", [("eta_re", prob.V_Gk, kap.real), ("eta_im", prob.V_Gk, kap.imag)], nsubcells=10)
        writevtk(prob.Geta, params.name + "_Ge_solution", [("eta_re", prob.V_Ge, et.real), ("eta_im", prob.V_Ge, et.imag)], nsubcells=10)
    return prob.postprocess(x)


class KhabakhpashevaTimeProblem(_Tank):
    """Everything up to `op = TransientConstantMatrixFEOperator(m,c,a,b,X,Y)` (:237-257); the
    time-dependent right-hand side b(t) = cos(ωt) b_c + sin(ωt) b_s is two assembled vectors."""

    def __init__(self, params, device=0, system_cls=B200System):
        self.params, self.order = params, params.order
        super().__init__(params)
        self.spaces(np.float64)
        c = self.c
        g, om, bh, k = c["g"], c["om"], c["bh"], c["k"]
        self.gamma_t, self.beta_t = 0.5, 0.25
        self.dt = c["T"] / 40
        self.tf = 50 * c["T"]
        cu = self.gamma_t / (self.beta_t * self.dt)
        ah = self.ah = cu / g * (1 - bh) / bh
        X = self.X
        sys_ = self.sys = system_cls(X.ndofs, False, nmat=4, nvec=2, device=device)
        ey = (0.0, 1.0, 0.0)
        d = sys_.add_domain(2, self.dOm.nq, self.dOm.wq, [self.dOm.own_slot(self.V_Om)])
        d.add_term(MAT_K, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)
        PHI, KAP = 0, 1
        for name, mu1f in (("Gfs", None), ("Gd1", self.mu1_in), ("Gd2", self.mu1_out)):
            dG = self.dG[name]
            if dG.ncells == 0:
                continue
            x1 = dG.points()[..., 0]
            fields = [mu1f(x1)] if mu1f is not None else []
            if name == "Gd1":      # b(t) on Γd1 (:253) split into its cos(ωt) and sin(ωt) parts
                m1 = fields[0]
                ck, sk = np.cos(k * x1), np.sin(k * x1)
                e0 = c["eta0"]
                fields += [-m1 * k * e0 * ck + ah * m1 * om * e0 * sk,      # w, cos part
                           m1 * om * e0 * sk,                               # u, cos part
                           -m1 * k * e0 * sk - ah * m1 * om * e0 * ck,      # w, sin part
                           -m1 * om * e0 * ck]                              # u, sin part
            d = sys_.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(self.V_Om), dG.surface_slot(self.V_Gk)],
                                measure_slot=1, weights=fields)
            d.add_term(MAT_K, bh * g, L.OP_VAL, KAP, L.OP_VAL, KAP)           # βₕ g κ u
            d.add_term(MAT_K, bh * g * ah, L.OP_VAL, PHI, L.OP_VAL, KAP)      # βₕ αₕ g κ w
            d.add_term(MAT_C, bh, L.OP_VAL, KAP, L.OP_VAL, PHI)               # βₕ ϕₜ u
            d.add_term(MAT_C, bh * ah, L.OP_VAL, PHI, L.OP_VAL, PHI)          # βₕ αₕ ϕₜ w
            d.add_term(MAT_C, -1.0, L.OP_VAL, PHI, L.OP_VAL, KAP)             # -κₜ w
            if mu1f is not None:
                d.add_term(MAT_K, -k, L.OP_VAL, PHI, L.OP_VAL, KAP, weight=0)
                d.add_term(MAT_K, 1.0, L.OP_VAL, KAP, L.OP_DDIR, PHI, weight=0, trial_dir=ey)
                d.add_term(MAT_K, ah, L.OP_VAL, PHI, L.OP_DDIR, PHI, weight=0, trial_dir=ey)
            if name == "Gd1":
                d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=1)
                d.add_rhs(0, 1.0, L.OP_VAL, KAP, weight_re=2)
                d.add_rhs(1, 1.0, L.OP_VAL, PHI, weight_re=3)
                d.add_rhs(1, 1.0, L.OP_VAL, KAP, weight_re=4)
        ETA = 1
        for name, ai in (("Gb1", c["a1"]), ("Gb2", c["a2"])):
            dG = self.dG[name]
            d = sys_.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(self.V_Om),
                                                  dG.surface_slot(self.V_Ge, need_hess=True)], measure_slot=1)
            d.add_term(MAT_K, g, L.OP_VAL, ETA, L.OP_VAL, ETA)
            d.add_term(MAT_K, ai, L.OP_LAPL, ETA, L.OP_LAPL, ETA)
            d.add_term(MAT_C, 1.0, L.OP_VAL, ETA, L.OP_VAL, PHI)              # ϕₜ v
            d.add_term(MAT_C, -1.0, L.OP_VAL, PHI, L.OP_VAL, ETA)             # -ηₜ w
            d.add_term(MAT_M, c["d0"], L.OP_VAL, ETA, L.OP_VAL, ETA)          # d₀ ηₜₜ v
        self._skeleton_domains(sys_, MAT_K)
        dGin = self.dG["Gin"]
        V_aux = TestFESpace(self.Gin, ReferenceFE(lagrangian, float, 1))
        V_aux.offset = 0
        x = dGin.points()
        amp = -(c["eta0"] * om) * (np.cosh(k * x[..., 1]) / np.sinh(k * c["H"]))
        d = sys_.add_domain(2, dGin.nq, dGin.wq, [dGin.trace_slot(self.V_Om), dGin.surface_slot(V_aux)],
                            measure_slot=1, weights=[amp * np.cos(k * x[..., 0]), amp * np.sin(k * x[..., 0])])
        d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=0)
        d.add_rhs(1, 1.0, L.OP_VAL, PHI, weight_re=1)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    def assemble(self):
        self.sys.assemble()
        cu = self.gamma_t / (self.beta_t * self.dt)
        cuu = 1.0 / (self.beta_t * self.dt ** 2)
        self.sys.combine(MAT_A, [MAT_K, MAT_C, MAT_M], [1.0, cu, cuu])

    def solve(self, nsteps=None, solver=None):
        """`solve(ode_solver,op,(x₀,v₀,a₀),t₀,tf)` + the step loop (:266-297): returns
        (ts, η DOF history (nsteps, n_η), final (u,v,a), statistics)."""
        c = self.c
        opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-13)
        opts.update(solver or {})
        self.sys.solver_setup(MAT_A, **opts)
        nst = nsteps if nsteps is not None else nsteps_for(0.0, self.tf, self.dt)
        ts = self.dt * np.arange(1, nst + 1)
        tcoef = np.stack([np.cos(c["om"] * ts), np.sin(c["om"] * ts)], axis=1)
        z = np.zeros(self.X.ndofs)
        o = self.X.offsets
        u, v, a, hist, st = self.sys.newmark_run(MAT_K, MAT_C, MAT_M, nst, self.dt, self.gamma_t, self.beta_t,
                                                 tcoef, z, z, z, out_range=(int(o[2]), int(o[3])), out_stride=1)
        return ts, hist, (u, v, a), st

    def postprocess(self, hist):
        """xs and |η|/η₀ per step in the reference's per-cell sorted layout (:275-290)."""
        xs, cd = self.eta_nodes()
        local = cd - 0          # V_Ge.cell_dofs are field-local ids
        return ((xs - 1.5 * self.c["Lb"]) / self.c["Lb"]).reshape(-1), np.abs(hist[:, local.reshape(-1)]) / self.c["eta0"]


def run_Khabakhpasheva_time_domain(params, device=0, nsteps=None, solver=None):
    prob = KhabakhpashevaTimeProblem(params, device=device)
    prob.assemble()
    ts, hist, _, prob.stats = prob.solve(nsteps=nsteps, solver=solver)
    xs, eta_rel = prob.postprocess(hist)
    return ts, xs, eta_rel
