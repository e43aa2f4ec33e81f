"""B200 mirror of `run_Liu` (src/Liu.jl:18-135): same parameters, same weak form on the
triangular model `models/Liu.json` (variable bathymetry), assembled and solved by libvlfs_b200 on
the GPU.  Configuration 5-3-1 (scripts/5-3-1-Liu.jl:20-32)."""
from dataclasses import dataclass
import numpy as np
from scipy.optimize import brentq
from .. import _lib as L
from ..geometry import DiscreteModelFromFile, Interior, Boundary, Skeleton
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System


@dataclass
class Liu_params:                       # src/Liu.jl:12-15
    name: str = "Liu"
    omega: float = 0.2                  # ω
    mesh_file: str = "models/Liu.json"  # (:65; a parameter here so that tests can point at a copy)


class LiuProblem:
    """Everything up to and including `op = AffineFEOperator(a,l,X,Y)` (:20-116)."""

    def __init__(self, params, device=0, system_cls=B200System, model=None):
        p = self.params = params
        om = p.omega
        m, EI, H0, Lb = 500, 1.0e10, 60, 300.0
        g, rho = 9.81, 1025
        d0, a1 = m / rho, EI / rho
        k = abs(brentq(lambda kk: np.sqrt(g * kk * np.tanh(kk * H0)) - om, 1e-8, 10.0, xtol=1e-15, rtol=8.9e-16))
        order = 4
        h = Lb / 50
        gamma = 1.0 * order * (order - 1) / h
        bh = 0.5
        ah = -1j * om / g * (1 - bh) / bh
        mu0, Ld, xd_out = 6.0, 4 * Lb, 9 * Lb
        self.c = c = dict(Lb=Lb, H0=H0, g=g, d0=d0, a1=a1, k=k, om=om, eta0=0.01, order=order, gamma=gamma,
                          bh=bh, ah=ah)
        mu1_in = lambda x1: mu0 * (1.0 - np.sin(np.pi / 2 * x1 / Ld))
        mu1_out = lambda x1: mu0 * (1.0 - np.cos(np.pi / 2 * (x1 - xd_out) / Ld))
        self.model = model = model if model is not None else DiscreteModelFromFile(p.mesh_file)
        Om = Interior(model)
        Gin = Boundary(model, tags="inlet")
        Gb = Boundary(model, tags="beam")
        Gd1 = Boundary(model, tags="damping_in")
        Gd2 = Boundary(model, tags="damping_out")
        Gf = Boundary(model, tags="free_surface")
        Gk = Boundary(model, tags=["free_surface", "damping_in", "damping_out"])
        Lb_ = Skeleton(Gb)
        self.Om, self.Gin, self.Gb, self.Gd1, self.Gd2, self.Gf, self.Gk, self.Lb = Om, Gin, Gb, Gd1, Gd2, Gf, Gk, Lb_
        degree = 2 * order
        cplx = np.complex128
        reffe = ReferenceFE(lagrangian, float, order)
        V_Om = TestFESpace(Om, reffe, conformity="H1", vector_type=cplx)
        V_Gk = TestFESpace(Gk, reffe, conformity="H1", vector_type=cplx)
        V_Ge = TestFESpace(Gb, reffe, conformity="H1", vector_type=cplx)
        self.X = X = MultiFieldFESpace([TrialFESpace(V_Om), TrialFESpace(V_Gk), TrialFESpace(V_Ge)])
        self.V_Om, self.V_Gk, self.V_Ge = V_Om, V_Gk, V_Ge
        sys_ = self.sys = system_cls(X.ndofs, True, nmat=1, nvec=1, device=device)
        ey = (0.0, 1.0, 0.0)                                               # ∇ₙ(ϕ) = ∇(ϕ)⋅(0,1) (:106)
        dOm = Measure(Om, degree)
        d = sys_.add_domain(2, dOm.nq, dOm.wq, [dOm.own_slot(V_Om)])
        d.add_term(0, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)                     # ∇(w)⋅∇(ϕ) (:108)
        PHI, KAP = 0, 1
        for trian, mu1f in ((Gf, None), (Gd1, mu1_in), (Gd2, mu1_out)):    # (:109-111)
            dG = Measure(trian, degree)
            x1 = dG.points()[..., 0]
            fields = [mu1f(x1)] if mu1f is not None else []
            if trian is Gd1:                                               # ηd, ∇ₙϕd (:61-62)
                m1 = fields[0]
                eta_d = m1 * k * c["eta0"] * np.exp(1j * k * x1)
                dphi_d = m1 * (-1j * om * c["eta0"] * np.exp(1j * k * x1))
                fields += [eta_d.real, eta_d.imag, dphi_d.real, dphi_d.imag]
            d = sys_.add_domain(2, dG.nq, dG.wq, [dG.trace_slot(V_Om), dG.surface_slot(V_Gk)],
                                measure_slot=1, weights=fields)
            d.add_term(0, bh * g, L.OP_VAL, KAP, L.OP_VAL, KAP)            # βₕ g κ u
            d.add_term(0, -1j * om * bh, L.OP_VAL, KAP, L.OP_VAL, PHI)     # -iω βₕ ϕ u
            d.add_term(0, bh * g * ah + 1j * om, L.OP_VAL, PHI, L.OP_VAL, KAP)   # (βₕ g αₕ + iω) κ w
            d.add_term(0, -1j * om * bh * ah, L.OP_VAL, PHI, L.OP_VAL, PHI)      # -iω βₕ αₕ ϕ w
            if mu1f is not None:
                d.add_term(0, -k, L.OP_VAL, PHI, L.OP_VAL, KAP, weight=0)                    # -μ₂ κ w
                d.add_term(0, 1.0, L.OP_VAL, KAP, L.OP_DDIR, PHI, weight=0, trial_dir=ey)   # μ₁ ∇ₙϕ u
                d.add_term(0, ah, L.OP_VAL, PHI, L.OP_DDIR, PHI, weight=0, trial_dir=ey)    # αₕ μ₁ ∇ₙϕ w
            if trian is Gd1:                                               # l (:114)
                d.add_rhs(0, -1.0, L.OP_VAL, PHI, weight_re=1, weight_im=2)   # -ηd w
                d.add_rhs(0, ah, L.OP_VAL, PHI, weight_re=3, weight_im=4)     # αₕ ∇ₙϕd w
                d.add_rhs(0, 1.0, L.OP_VAL, KAP, weight_re=3, weight_im=4)    # ∇ₙϕd u
        ETA = 1
        dGb = Measure(Gb, degree)                                          # (:112)
        d = sys_.add_domain(2, dGb.nq, dGb.wq, [dGb.trace_slot(V_Om), dGb.surface_slot(V_Ge, need_hess=True)],
                            measure_slot=1)
        d.add_term(0, -om ** 2 * d0 + g, L.OP_VAL, ETA, L.OP_VAL, ETA)
        d.add_term(0, a1, L.OP_LAPL, ETA, L.OP_LAPL, ETA)
        d.add_term(0, -1j * om, L.OP_VAL, ETA, L.OP_VAL, PHI)
        d.add_term(0, 1j * om, L.OP_VAL, PHI, L.OP_VAL, ETA)
        dLb = Measure(Lb_, degree)                                         # (:113); slot 0 = η⁺, 1 = η⁻
        d = sys_.add_domain(2, dLb.nq, dLb.wq, [dLb.skeleton_slot(V_Ge, "plus"), dLb.skeleton_slot(V_Ge, "minus")],
                            measure_slot=0, measure_kind=L.MEASURE_FACET)
        jump_n, mean = {0: 1.0, 1: 1.0}, {0: 0.5, 1: 0.5}
        d.add_term(0, -a1, L.OP_GRAD_NOWN, jump_n, L.OP_LAPL, mean)
        d.add_term(0, -a1, L.OP_LAPL, mean, L.OP_GRAD_NOWN, jump_n)
        d.add_term(0, a1 * gamma, L.OP_GRAD_NOWN, jump_n, L.OP_GRAD_NOWN, jump_n)
        dGin = Measure(Gin, degree)                                        # w vᵢₙ (:42,114)
        V_aux = TestFESpace(Gin, ReferenceFE(lagrangian, float, 1))
        V_aux.offset = 0
        x = dGin.points()
        vin = (c["eta0"] * om) * (np.cosh(k * (x[..., 1] + 0.075 * Lb)) / np.sinh(k * H0)) * np.exp(1j * k * x[..., 0])
        d = sys_.add_domain(2, dGin.nq, dGin.wq, [dGin.trace_slot(V_Om), dGin.surface_slot(V_aux)],
                            measure_slot=1, weights=[vin.real, vin.imag])
        d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=0, weight_im=1)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    def assemble(self):
        self.sys.assemble()

    def postprocess(self, x):
        """(xs, |η|/η₀) (:119-126): DOF nodes of every beam cell, sorted with the first cell's
        permutation, abscissa (x - 6 Lb)/Lb."""
        xc = self.V_Ge.cell_dof_coords()[:, :, 0]
        p = np.argsort(xc[0], kind="stable")
        eta = self.X.split(x)[2]
        return (((xc[:, p] - 6 * self.c["Lb"]) / self.c["Lb"]).reshape(-1),
                (np.abs(eta[self.V_Ge.cell_dofs[:, p]]) / self.c["eta0"]).reshape(-1))


def run_Liu(params, device=0, solver=None):
    """`run_Liu(params)` (src/Liu.jl:18-135): returns (xs, |η|/η₀)."""
    prob = LiuProblem(params, device=device)
    prob.assemble()
    opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    opts.update(solver or {})
    prob.sys.solver_setup(0, **opts)
    x, prob.stats = prob.sys.solve(vec=0)
    return prob.postprocess(x)
