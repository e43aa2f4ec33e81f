"""B200 mirror of `run_periodic_beam` (src/Periodic_Beam.jl:23-165): same parameters, same
weak form; K, C, M assembled on one union pattern, A = K + γ/(βΔt) C + 1/(βΔt²) M combined,
factorised and stepped (Newmark) by libvlfs_b200 on the GPU."""
from dataclasses import dataclass
import numpy as np
from .. import _lib as L
from ..geometry import CartesianDiscreteModel, add_tag_from_tags, Interior, Boundary, Skeleton
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System

MAT_K, MAT_C, MAT_M, MAT_A = 0, 1, 2, 3


@dataclass
class Periodic_Beam_params:             # src/Periodic_Beam.jl:12-21
    name: str = "PeriodicBeam"
    n: int = 4
    dt: float = 0.001
    tf: float = 1.0
    orderphi: int = 2                   # orderϕ
    ordereta: int = 2                   # orderη
    k: int = 10
    vtk_output: bool = False


def nsteps_for(t0, tf, dt):
    """States produced by `solve(ode_solver,op,x0,t0,tf)`: it advances while t < tf - ε."""
    n, t = 0, t0
    while t < tf - 100 * np.finfo(float).eps * max(abs(tf), 1.0):
        t += dt
        n += 1
    return n


class PeriodicBeamProblem:
    """Everything up to `op = TransientConstantFEOperator(m,c,a,b,X,Y)` (:30-104)."""

    def __init__(self, params, device=0, system_cls=B200System):
        p = self.params = params
        n, op_, oe, k = p.n, p.orderphi, p.ordereta, p.k
        Lx, H = 2.0 * np.pi, 1.0
        g, rho_w, rho_b, h_b = 9.81, 1.0e3, 1.0e2, 1.0e-2
        om = np.sqrt(g * k * np.tanh(k * H))
        EI_b = rho_b * h_b * om ** 2 / (k ** 4)
        d0 = rho_b * h_b / rho_w
        Drho = EI_b / rho_w
        h = Lx / n
        gamma = 10.0 * oe * (oe + 1)
        self.c = dict(L=Lx, H=H, g=g, om=om, d0=d0, Drho=Drho, eta0=0.01, h=h, gamma=gamma, k=k,
                      gamma_t=0.5, beta_t=0.25)
        model = CartesianDiscreteModel((0.0, Lx, 0.0, H), (2 * n, n), isperiodic=(True, False))
        add_tag_from_tags(model, "bottom", [1, 2, 5])
        add_tag_from_tags(model, "beam", [3, 4, 6])
        self.model = model
        Om = Interior(model)
        G = Boundary(model, tags="beam")
        Lam = Skeleton(G)
        self.Om, self.G, self.Lam = Om, G, Lam
        degree = 2 * oe
        dOm, dG, dLam = Measure(Om, degree), Measure(G, degree), Measure(Lam, degree)
        V_Om = TestFESpace(Om, ReferenceFE(lagrangian, float, op_), conformity="H1")
        V_G = TestFESpace(G, ReferenceFE(lagrangian, float, oe), conformity="H1")
        self.X = X = MultiFieldFESpace([TrialFESpace(V_Om), TrialFESpace(V_G)])
        self.V_Om, self.V_G = V_Om, V_G

        sys_ = self.sys = system_cls(X.ndofs, False, nmat=4, nvec=1, device=device)
        # a: ∫ ∇(ϕ)⋅∇(w) dΩ (:100)
        # (one coefficient field per domain: the analytic solution at the points, for l2_Ω/l2_Γ)
        self._xq_Om, self._xq_G = dOm.points(), dG.points()
        d = self.dom_Om = sys_.add_domain(2, dOm.nq, dOm.wq, [dOm.own_slot(V_Om)],
                                          weights=[np.zeros(self._xq_Om.shape[:2])])
        d.add_term(MAT_K, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)
        # Γ: slot 0 = ϕ trace, slot 1 = η  (m :98, c :99, a :101)
        d = self.dom_G = sys_.add_domain(2, dG.nq, dG.wq,
                                         [dG.trace_slot(V_Om), dG.surface_slot(V_G, need_hess=True)],
                                         measure_slot=1, weights=[np.zeros(self._xq_G.shape[:2])])
        PHI, ETA = 0, 1
        d.add_term(MAT_K, g, L.OP_VAL, ETA, L.OP_VAL, ETA)              # g η v
        d.add_term(MAT_K, Drho, L.OP_LAPL, ETA, L.OP_LAPL, ETA)         # Dᵨ Δ(η) Δ(v)
        d.add_term(MAT_C, 1.0, L.OP_VAL, ETA, L.OP_VAL, PHI)            # ϕₜ v
        d.add_term(MAT_C, -1.0, L.OP_VAL, PHI, L.OP_VAL, ETA)           # -ηₜ w
        d.add_term(MAT_M, d0, L.OP_VAL, ETA, L.OP_VAL, ETA)             # d₀ ηₜₜ v
        # Λ (:102-103): slot 0 = η⁺, slot 1 = η⁻ ; jump(∇(v)⋅nΛ) = ∇v⁺⋅n⁺ + ∇v⁻⋅n⁻
        d = self.dom_L = sys_.add_domain(2, dLam.nq, dLam.wq,
                                         [dLam.skeleton_slot(V_G, "plus"), dLam.skeleton_slot(V_G, "minus")],
                                         measure_slot=0, measure_kind=L.MEASURE_FACET)
        jump_n, mean = {0: 1.0, 1: 1.0}, {0: 0.5, 1: 0.5}
        d.add_term(MAT_K, -Drho, L.OP_GRAD_NOWN, jump_n, L.OP_LAPL, mean)
        d.add_term(MAT_K, -Drho, L.OP_LAPL, mean, L.OP_GRAD_NOWN, jump_n)
        d.add_term(MAT_K, Drho * gamma / h, L.OP_GRAD_NOWN, jump_n, L.OP_GRAD_NOWN, jump_n)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    # analytic fields (:44-45) and time derivatives, for `interpolate_everywhere` (:111-113)
    def eta_ex(self, x, t, d=0):
        c = self.c
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [c["eta0"] * np.cos(ph), c["eta0"] * c["om"] * np.sin(ph),
                -c["eta0"] * c["om"] ** 2 * np.cos(ph)][d]

    def phi_ex(self, x, t, d=0):
        c = self.c
        amp = c["eta0"] * c["om"] / c["k"] * np.cosh(c["k"] * x[..., 1]) / np.sinh(c["k"] * c["H"])
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [amp * np.sin(ph), -amp * c["om"] * np.cos(ph), -amp * c["om"] ** 2 * np.sin(ph)][d]

    def interpolate(self, t, d=0):
        xd = self.X.dof_coords()
        o = self.X.offsets
        x = np.zeros(self.X.ndofs)
        x[o[0]:o[1]] = self.phi_ex(xd[o[0]:o[1]], t, d)
        x[o[1]:o[2]] = self.eta_ex(xd[o[1]:o[2]], t, d)
        return x

    def assemble(self):
        """K, C, M and the Newmark system matrix A (:104-108)."""
        c, p = self.c, self.params
        self.sys.assemble()
        cu = c["gamma_t"] / (c["beta_t"] * p.dt)
        cuu = 1.0 / (c["beta_t"] * p.dt ** 2)
        self.sys.combine(MAT_A, [MAT_K, MAT_C, MAT_M], [1.0, cu, cuu])

    def solve(self, solver=None, keep_states=True):
        """`solve(ode_solver,op,(x₀,v₀,a₀),t₀,tf)` (:116): returns (u,v,a) at tf, the states
        of every step (nsteps, ndofs) and the solver statistics."""
        c, p = self.c, self.params
        opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-13)
        opts.update(solver or {})
        self.sys.solver_setup(MAT_A, **opts)
        nst = nsteps_for(0.0, p.tf, p.dt)
        u, v, a, out, st = self.sys.newmark_run(
            MAT_K, MAT_C, MAT_M, nst, p.dt, c["gamma_t"], c["beta_t"], None,
            self.interpolate(0.0, 0), self.interpolate(0.0, 1), self.interpolate(0.0, 2),
            out_range=(0, self.X.ndofs if keep_states else 0), out_stride=1)
        return u, v, a, out, st


def run_periodic_beam(params, device=0, solver=None):
    """`run_periodic_beam(params)` (src/Periodic_Beam.jl:23-165): returns
    (e_ϕ, e_η, E_kin_f, E_pot_f, E_kin_s, E_ela_s, E_kin_f₀, E_kin_s₀, E_pot_f₀, E_ela_s₀, t_global);
    the per-step integrals (:141-148) are evaluated on the GPU (vlfs_eval_functional)."""
    prob = PeriodicBeamProblem(params, device=device)
    prob.assemble()
    u, v, a, states, prob.stats = prob.solve(solver=solver)
    c, p = prob.c, params
    PHI, ETA = 0, 1
    e_phi, e_eta, Ekf, Epf, Eks, Ees, ts = [], [], [], [], [], [], []
    prev = prob.interpolate(0.0, 0)
    for n in range(states.shape[0]):
        t = (n + 1) * p.dt
        x = states[n]
        prob.dom_Om.set_weights(0, prob.phi_ex(prob._xq_Om, t))
        prob.dom_G.set_weights(0, prob.eta_ex(prob._xq_G, t))
        e_phi.append(np.sqrt(prob.dom_Om.functional(x, L.OP_VAL, 0, weight_f=0)))
        e_eta.append(np.sqrt(prob.dom_G.functional(x, L.OP_VAL, ETA, weight_f=0)))
        Ekf.append(0.5 * prob.dom_Om.functional(x, L.OP_GRAD, 0))
        Epf.append(0.5 * c["g"] * prob.dom_G.functional(x, L.OP_VAL, ETA))
        Eks.append(0.5 * c["d0"] * prob.dom_G.functional((x - prev) / p.dt, L.OP_VAL, ETA))
        Ees.append(0.5 * c["Drho"] * prob.dom_G.functional(x, L.OP_LAPL, ETA))
        ts.append(t)
        prev = x
    E_kin_f0 = 0.25 * c["d0"] * c["om"] ** 2 * c["eta0"] ** 2 * c["L"]
    E_kin_s0 = 0.25 * c["g"] * c["eta0"] ** 2 * c["L"]
    E_pot_f0 = 0.25 * c["Drho"] * c["k"] ** 4 * c["eta0"] ** 2 * c["L"]
    E_ela_s0 = 0.25 * c["g"] * c["eta0"] ** 2 * c["L"]
    return (np.array(e_phi), np.array(e_eta), np.array(Ekf), np.array(Epf), np.array(Eks), np.array(Ees),
            E_kin_f0, E_kin_s0, E_pot_f0, E_ela_s0, np.array(ts))
