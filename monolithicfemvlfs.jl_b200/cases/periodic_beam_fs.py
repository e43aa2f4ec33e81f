"""B200 mirror of `run_periodic_beam_FS` (src/Periodic_Beam_FS.jl:22-213): same parameters, same
weak form (beam on x ∈ [π/2, 3π/2], free surface elsewhere, three fields); K, C, M assembled on one
union pattern, A = K + γ/(βΔt) C + 1/(βΔt²) M combined, factorised and stepped (Newmark) by
libvlfs_b200 on the GPU; the per-step errors and energies are GPU functionals."""
from dataclasses import dataclass
import numpy as np
from .. import _lib as L
from ..geometry import CartesianDiscreteModel, add_tag_from_tags, Interior, Boundary, Triangulation, Skeleton
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System
from .periodic_beam import nsteps_for, MAT_K, MAT_C, MAT_M, MAT_A
from .khabakhpasheva import isapprox


@dataclass
class Periodic_Beam_FS_params:          # src/Periodic_Beam_FS.jl:12-20
    name: str = "PeriodicBeamFS"
    n: int = 4
    dt: float = 0.001
    tf: float = 1.0
    order: int = 2
    k: int = 10
    vtk_output: bool = False


class PeriodicBeamFSProblem:
    """Everything up to `op = TransientConstantFEOperator(m,c,a,b,X,Y)` (:27-147)."""

    def __init__(self, params, device=0, system_cls=B200System):
        p = self.params = params
        n, order, k, dt = p.n, p.order, p.k, p.dt
        Lx, H = 2.0 * np.pi, 1.0
        g, rho_w, rho_b, h_b = 9.81, 1.0e3, 1.0e2, 1.0e-2
        om = np.sqrt(g * k * np.tanh(k * H))
        EI_b = rho_b * h_b * om ** 2 / (k ** 4)
        d0, Drho = rho_b * h_b / rho_w, EI_b / rho_w
        gamma_t, beta_t, bh = 0.5, 0.25, 0.5
        ah = gamma_t / (beta_t * dt) / g * (1 - bh) / bh
        h = Lx / n
        gamma = 10.0 * order * (order - 1) / h
        Lb, xb0 = np.pi, 0.5 * np.pi
        xb1 = xb0 + Lb
        self.c = dict(L=Lx, H=H, g=g, om=om, d0=d0, Drho=Drho, eta0=0.01, k=k, gamma_t=gamma_t, beta_t=beta_t,
                      bh=bh, ah=ah, Lb=Lb)
        model = CartesianDiscreteModel((0.0, Lx, 0.0, H), (2 * n, n), isperiodic=(True, False))
        add_tag_from_tags(model, "surface", [3, 4, 6])
        self.model = model
        Om = Interior(model)
        G = Boundary(model, tags="surface")
        xc = G.cell_coords().mean(axis=1)                                   # is_beam (:84-88)
        beam = (xb0 <= xc[:, 0]) & (xc[:, 0] <= xb1) & isapprox(xc[:, 1], H)
        Gb = Triangulation(G, np.nonzero(beam)[0])
        Gfs = Triangulation(G, np.nonzero(~beam)[0])
        Lam = Skeleton(Gb)
        self.Om, self.G, self.Gb, self.Gfs, self.Lam = Om, G, Gb, Gfs, Lam
        degree = 2 * order
        dOm, dGb, dGfs, dLam = Measure(Om, degree), Measure(Gb, degree), Measure(Gfs, degree), Measure(Lam, degree)
        reffe = ReferenceFE(lagrangian, float, order)
        V_Om = TestFESpace(Om, reffe, conformity="H1")
        V_Gfs = TestFESpace(Gfs, reffe, conformity="H1")
        V_Gb = TestFESpace(Gb, reffe, conformity="H1")
        self.X = X = MultiFieldFESpace([TrialFESpace(V_Om), TrialFESpace(V_Gfs), TrialFESpace(V_Gb)])
        self.V_Om, self.V_Gfs, self.V_Gb = V_Om, V_Gfs, V_Gb

        sys_ = self.sys = system_cls(X.ndofs, False, nmat=4, nvec=1, device=device)
        # (one coefficient field per domain: the analytic solution at the points, for the L2 errors)
        self._xq_Om, self._xq_Gfs, self._xq_Gb = dOm.points(), dGfs.points(), dGb.points()
        d = self.dom_Om = sys_.add_domain(2, dOm.nq, dOm.wq, [dOm.own_slot(V_Om)],
                                          weights=[np.zeros(self._xq_Om.shape[:2])])
        d.add_term(MAT_K, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)                  # ∇(w)⋅∇(ϕ) (:142)
        # Γfs: slot 0 = ϕ trace, slot 1 = κ  (c :139, a :143)
        PHI, KAP = 0, 1
        d = self.dom_Gfs = sys_.add_domain(2, dGfs.nq, dGfs.wq, [dGfs.trace_slot(V_Om), dGfs.surface_slot(V_Gfs)],
                                           measure_slot=1, weights=[np.zeros(self._xq_Gfs.shape[:2])])
        d.add_term(MAT_C, bh, L.OP_VAL, KAP, L.OP_VAL, PHI)                 # βₕ ϕₜ u
        d.add_term(MAT_C, bh * ah, L.OP_VAL, PHI, L.OP_VAL, PHI)            # βₕ αₕ ϕₜ w
        d.add_term(MAT_C, -1.0, L.OP_VAL, PHI, L.OP_VAL, KAP)               # -κₜ w
        d.add_term(MAT_K, bh * g, L.OP_VAL, KAP, L.OP_VAL, KAP)             # βₕ g κ u
        d.add_term(MAT_K, bh * ah * g, L.OP_VAL, PHI, L.OP_VAL, KAP)        # βₕ αₕ g κ w
        # Γb: slot 0 = ϕ trace, slot 1 = η  (m :138, c :140, a :144)
        ETA = 1
        d = self.dom_Gb = sys_.add_domain(2, dGb.nq, dGb.wq,
                                          [dGb.trace_slot(V_Om), dGb.surface_slot(V_Gb, need_hess=True)],
                                          measure_slot=1, weights=[np.zeros(self._xq_Gb.shape[:2])])
        d.add_term(MAT_K, g, L.OP_VAL, ETA, L.OP_VAL, ETA)                  # g η v
        d.add_term(MAT_K, Drho, L.OP_LAPL, ETA, L.OP_LAPL, ETA)             # Dᵨ Δ(v) Δ(η)
        d.add_term(MAT_C, 1.0, L.OP_VAL, ETA, L.OP_VAL, PHI)                # ϕₜ v
        d.add_term(MAT_C, -1.0, L.OP_VAL, PHI, L.OP_VAL, ETA)               # -ηₜ w
        d.add_term(MAT_M, d0, L.OP_VAL, ETA, L.OP_VAL, ETA)                 # d₀ ηₜₜ v
        # Λb (:145): slot 0 = η⁺, slot 1 = η⁻
        d = self.dom_L = sys_.add_domain(2, dLam.nq, dLam.wq,
                                         [dLam.skeleton_slot(V_Gb, "plus"), dLam.skeleton_slot(V_Gb, "minus")],
                                         measure_slot=0, measure_kind=L.MEASURE_FACET)
        jump_n, mean = {0: 1.0, 1: 1.0}, {0: 0.5, 1: 0.5}
        d.add_term(MAT_K, -Drho, L.OP_GRAD_NOWN, jump_n, L.OP_LAPL, mean)
        d.add_term(MAT_K, -Drho, L.OP_LAPL, mean, L.OP_GRAD_NOWN, jump_n)
        d.add_term(MAT_K, Drho * gamma, L.OP_GRAD_NOWN, jump_n, L.OP_GRAD_NOWN, jump_n)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    def eta_ex(self, x, t, d=0):
        c = self.c
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [c["eta0"] * np.cos(ph), c["eta0"] * c["om"] * np.sin(ph), -c["eta0"] * c["om"] ** 2 * np.cos(ph)][d]

    def phi_ex(self, x, t, d=0):
        c = self.c
        amp = c["eta0"] * c["om"] / c["k"] * np.cosh(c["k"] * x[..., 1]) / np.sinh(c["k"] * c["H"])
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [amp * np.sin(ph), -amp * c["om"] * np.cos(ph), -amp * c["om"] ** 2 * np.sin(ph)][d]

    def interpolate(self, t, d=0):
        """`interpolate_everywhere([ϕ(t),η(t),η(t)],X(t))` (:154-156)."""
        xd, o = self.X.dof_coords(), self.X.offsets
        x = np.zeros(self.X.ndofs)
        x[o[0]:o[1]] = self.phi_ex(xd[o[0]:o[1]], t, d)
        x[o[1]:o[3]] = self.eta_ex(xd[o[1]:o[3]], t, d)
        return x

    def assemble(self):
        c, p = self.c, self.params
        self.sys.assemble()
        cu = c["gamma_t"] / (c["beta_t"] * p.dt)
        cuu = 1.0 / (c["beta_t"] * p.dt ** 2)
        self.sys.combine(MAT_A, [MAT_K, MAT_C, MAT_M], [1.0, cu, cuu])

    def solve(self, solver=None, keep_states=True):
        c, p = self.c, self.params
        opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=10, maxit=20, rtol=1e-13)
        opts.update(solver or {})
        self.sys.solver_setup(MAT_A, **opts)
        nst = nsteps_for(0.0, p.tf, p.dt)
        return self.sys.newmark_run(
            MAT_K, MAT_C, MAT_M, nst, p.dt, c["gamma_t"], c["beta_t"], None,
            self.interpolate(0.0, 0), self.interpolate(0.0, 1), self.interpolate(0.0, 2),
            out_range=(0, self.X.ndofs if keep_states else 0), out_stride=1)


def run_periodic_beam_FS(params, device=0, solver=None):
    """`run_periodic_beam_FS(params)` (src/Periodic_Beam_FS.jl:22-213): returns
    (e_ϕ, e_η, E_kin_f, E_pot_f, E_kin_s, E_ela_s, E_kin_f₀, E_kin_s₀, E_pot_f₀, E_ela_s₀, t_global);
    the per-step integrals (:191-198) are evaluated on the GPU (vlfs_eval_functional)."""
    prob = PeriodicBeamFSProblem(params, device=device)
    prob.assemble()
    u, v, a, states, prob.stats = prob.solve(solver=solver)
    c, p = prob.c, params
    KAP, ETA = 1, 1
    e_phi, e_eta, Ekf, Epf, Eks, Ees, ts = [], [], [], [], [], [], []
    prev = prob.interpolate(0.0, 0)
    for n in range(states.shape[0]):
        t = (n + 1) * p.dt
        x = states[n]
        prob.dom_Om.set_weights(0, prob.phi_ex(prob._xq_Om, t))
        prob.dom_Gfs.set_weights(0, prob.eta_ex(prob._xq_Gfs, t))
        prob.dom_Gb.set_weights(0, prob.eta_ex(prob._xq_Gb, t))
        e_phi.append(np.sqrt(prob.dom_Om.functional(x, L.OP_VAL, 0, weight_f=0)))
        e_eta.append(np.sqrt(prob.dom_Gfs.functional(x, L.OP_VAL, KAP, weight_f=0))
                     + np.sqrt(prob.dom_Gb.functional(x, L.OP_VAL, ETA, weight_f=0)))
        Ekf.append(0.5 * prob.dom_Om.functional(x, L.OP_GRAD, 0))
        Epf.append(0.5 * c["g"] * prob.dom_Gfs.functional(x, L.OP_VAL, KAP)
                   + 0.5 * c["g"] * prob.dom_Gb.functional(x, L.OP_VAL, ETA))
        Eks.append(0.5 * c["d0"] * prob.dom_Gb.functional((x - prev) / p.dt, L.OP_VAL, ETA))
        Ees.append(0.5 * c["Drho"] * prob.dom_Gb.functional(x, L.OP_LAPL, ETA))
        ts.append(t)
        prev = x
    E_kin_s0 = 0.25 * c["d0"] * c["om"] ** 2 * c["eta0"] ** 2 * c["Lb"]
    E_kin_f0 = 0.25 * c["g"] * c["eta0"] ** 2 * c["L"]
    E_ela_s0 = 0.25 * c["Drho"] * c["k"] ** 4 * c["eta0"] ** 2 * c["Lb"]
    E_pot_f0 = 0.25 * c["g"] * c["eta0"] ** 2 * c["L"]
    return (np.array(e_phi), np.array(e_eta), np.array(Ekf), np.array(Epf), np.array(Eks), np.array(Ees),
            E_kin_f0, E_kin_s0, E_pot_f0, E_ela_s0, np.array(ts))
