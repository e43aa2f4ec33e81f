"""Oracle restatement of `run_periodic_beam` (`src/Periodic_Beam.jl:23-165`).

2-D periodic fluid domain with an Euler-Bernoulli beam over the whole free surface, time
domain, Newmark(γ=½,β=¼).  The analytic travelling wave (`:44-45`) makes this the
known-answer anchor of the oracle: L2 errors must converge like h^(order+1)
(`scripts/5-1-1-periodic-beam-spatial-convergence.jl:140-141,183-209`).
Test infrastructure only.
"""
import numpy as np
import scipy.sparse as sp
from ..model import cartesian_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO
from ..newmark import newmark, nsteps_for


def constants(n, orderphi, ordereta, k):
    """`src/Periodic_Beam.jl:30-56`."""
    L, H = 2.0 * np.pi, 1.0
    g, rho_w, rho_b, h_b = 9.81, 1.0e3, 1.0e2, 1.0e-2
    om = np.sqrt(g * k * np.tanh(k * H))
    EI_b = rho_b * h_b * om ** 2 / (k ** 4)
    d0 = rho_b * h_b / rho_w
    Drho = EI_b / rho_w
    eta0 = 0.01
    h = L / n
    gamma = 10.0 * ordereta * (ordereta + 1)
    return dict(L=L, H=H, g=g, om=om, d0=d0, Drho=Drho, eta0=eta0, h=h, gamma=gamma, k=k,
                gamma_t=0.5, beta_t=0.25)


class PeriodicBeamCase:
    def __init__(self, n=4, dt=0.001, tf=1.0, orderphi=2, ordereta=2, k=10):
        self.p = dict(n=n, dt=dt, tf=tf, orderphi=orderphi, ordereta=ordereta, k=k)
        c = self.c = constants(n, orderphi, ordereta, k)
        self.model = m = cartesian_model((0.0, c["L"], 0.0, c["H"]), (2 * n, n),
                                         isperiodic=(True, False))
        m.add_tag_from_entities("bottom", [1, 2, 5])      # :67
        m.add_tag_from_entities("beam", [3, 4, 6])        # :68
        self.Om = VolumeTriangulation(m)
        self.G = boundary(m, "beam")
        self.Lam = SkeletonTriangulation(self.G)
        self.degree = 2 * ordereta                         # :79
        self.V_phi = FESpace(self.Om, orderphi)
        self.V_eta = FESpace(self.G, ordereta)
        self.X = MultiFieldFESpace([self.V_phi, self.V_eta])

    # analytic fields (:44-45) and their time derivatives
    def eta_ex(self, x, t, d=0):
        c = self.c
        ph = c["k"] * x[..., 0] - c["om"] * t
        if d == 0:
            return c["eta0"] * np.cos(ph)
        if d == 1:
            return c["eta0"] * c["om"] * np.sin(ph)
        return -c["eta0"] * c["om"] ** 2 * np.cos(ph)

    def phi_ex(self, x, t, d=0):
        c = self.c
        amp = c["eta0"] * c["om"] / c["k"] * np.cosh(c["k"] * x[..., 1]) / np.sinh(c["k"] * c["H"])
        ph = c["k"] * x[..., 0] - c["om"] * t
        if d == 0:
            return amp * np.sin(ph)
        if d == 1:
            return -amp * c["om"] * np.cos(ph)
        return -amp * c["om"] ** 2 * np.sin(ph)

    # ------------------------------------------------------------------
    def assemble(self):
        """K, C, M on the union pattern of the three jacobians (`:96-104`)."""
        c = self.c
        op, oe, deg = self.p["orderphi"], self.p["ordereta"], self.degree
        o_phi, o_eta = self.X.offsets[:2]
        N = self.X.ndofs
        K, C, M = (COO(N, np.float64) for _ in range(3))

        def add(rows, cols, k=None, cc=None, mm=None):
            z = np.zeros((rows.shape[0], rows.shape[1], cols.shape[1]))
            K.add_block(rows, cols, z if k is None else k)
            C.add_block(rows, cols, z if cc is None else cc)
            M.add_block(rows, cols, z if mm is None else mm)

        # Ω: ∫ ∇ϕ·∇w
        dOm = VolumeMeasure(self.Om, deg)
        o = dOm.ops(op)
        dofs = self.V_phi.cell_dofs + o_phi
        add(dofs, dofs, k=np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))
        # Γ: m = d₀ ηₜₜ v ; c = ϕₜ v − ηₜ w ; a = g η v + Dρ Δη Δv
        dG = SurfaceMeasure(self.G, deg)
        S = dG.surface_ops(oe, need_hess=True)
        T = dG.trace_ops(op)
        rows_w = self.V_phi.cell_dofs[self.G.cell] + o_phi
        rows_v = self.V_eta.cell_dofs + o_eta
        mass = np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
        add(rows_v, rows_v,
            k=c["g"] * mass + c["Drho"] * np.einsum("eq,eqi,eqj->eij", S.dV, S.lapl, S.lapl),
            mm=c["d0"] * mass)
        add(rows_v, rows_w, cc=np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
        add(rows_w, rows_v, cc=-np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        # Λ: Dρ( −{Δη}⟦∇v·n⟧ − ⟦∇η·n⟧{Δv} + γ/h ⟦∇v·n⟧⟦∇η·n⟧ );  ⟦a·n⟧ = a⁺·n⁺ + a⁻·n⁻
        dL = SkeletonMeasure(self.Lam, deg)
        P = dL.side_ops(oe, "plus")
        Mi = dL.side_ops(oe, "minus")
        rows_p = self.V_eta.cell_dofs[self.Lam.plus] + o_eta
        rows_m = self.V_eta.cell_dofs[self.Lam.minus] + o_eta
        sides = [(P, rows_p), (Mi, rows_m)]
        pen = c["gamma"] / c["h"]
        for (Sv, rv) in sides:
            Jv = np.einsum("eqib,eqb->eqi", Sv.grad, Sv.normal)
            Mv = 0.5 * Sv.lapl
            for (Se, re_) in sides:
                Je = np.einsum("eqjb,eqb->eqj", Se.grad, Se.normal)
                Me = 0.5 * Se.lapl
                blk = c["Drho"] * (-np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Me)
                                   - np.einsum("eq,eqi,eqj->eij", P.dS, Mv, Je)
                                   + pen * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je))
                add(rv, re_, k=blk)
        return K.tocsr(), C.tocsr(), M.tocsr()

    # ------------------------------------------------------------------
    def interpolate(self, t, d=0):
        """`interpolate_everywhere([ϕ(t),η(t)],X(t))` (`:111-113`): nodal values."""
        x = np.zeros(self.X.ndofs)
        o_phi, o_eta = self.X.offsets[:2]
        x[o_phi:o_phi + self.V_phi.ndofs] = self.phi_ex(self.V_phi.dof_coords(), t, d)
        x[o_eta:o_eta + self.V_eta.ndofs] = self.eta_ex(self.V_eta.dof_coords(), t, d)
        return x

    def functionals(self):
        """Callables for the per-step integrals of `:141-148`."""
        c = self.c
        op, oe, deg = self.p["orderphi"], self.p["ordereta"], self.degree
        o_phi, o_eta = self.X.offsets[:2]
        dOm = VolumeMeasure(self.Om, deg).ops(op)
        dG = SurfaceMeasure(self.G, deg).surface_ops(oe, need_hess=True)
        cd_phi = self.V_phi.cell_dofs + o_phi
        cd_eta = self.V_eta.cell_dofs + o_eta

        def l2_phi(x, t):
            uh = np.einsum("eqi,ei->eq", dOm.val, x[cd_phi])
            return np.sqrt(np.sum(dOm.dV * (self.phi_ex(dOm.x, t) - uh) ** 2))

        def l2_eta(x, t):
            uh = np.einsum("eqi,ei->eq", dG.val, x[cd_eta])
            return np.sqrt(np.sum(dG.dV * (self.eta_ex(dG.x, t) - uh) ** 2))

        def energies(x, xprev, dt):
            gphi = np.einsum("eqib,ei->eqb", dOm.grad, x[cd_phi])
            eta = np.einsum("eqi,ei->eq", dG.val, x[cd_eta])
            eta_t = np.einsum("eqi,ei->eq", dG.val, (x - xprev)[cd_eta]) / dt
            lap = np.einsum("eqi,ei->eq", dG.lapl, x[cd_eta])
            return (0.5 * np.sum(dOm.dV * np.sum(gphi * gphi, axis=-1)),
                    0.5 * c["g"] * np.sum(dG.dV * eta * eta),
                    0.5 * c["d0"] * np.sum(dG.dV * eta_t * eta_t),
                    0.5 * c["Drho"] * np.sum(dG.dV * lap * lap))
        return l2_phi, l2_eta, energies

    def run(self, keep_states=False):
        """Returns dict with e_phi, e_eta, energies per step, t and the final state."""
        c, p = self.c, self.p
        K, C, M = self.assemble()
        nst = nsteps_for(0.0, p["tf"], p["dt"])
        l2_phi, l2_eta, energies = self.functionals()
        out = dict(t=[], e_phi=[], e_eta=[], E=[], states=[])
        prev = [self.interpolate(0.0, 0)]

        def cb(n, t, u, v, a):
            out["t"].append(t)
            out["e_phi"].append(l2_phi(u, t))
            out["e_eta"].append(l2_eta(u, t))
            out["E"].append(energies(u, prev[0], p["dt"]))
            prev[0] = u.copy()
            if keep_states:
                out["states"].append(u.copy())

        u, v, a = newmark(K, C, M, None, self.interpolate(0.0, 0), self.interpolate(0.0, 1),
                          self.interpolate(0.0, 2), p["dt"], c["gamma_t"], c["beta_t"], nst, callback=cb)
        out["u"], out["v"], out["a"] = u, v, a
        # analytic energy constants (`:129-132`, names as in the reference)
        out["E0"] = dict(E_kin_f0=0.25 * c["d0"] * c["om"] ** 2 * c["eta0"] ** 2 * c["L"],
                         E_kin_s0=0.25 * c["g"] * c["eta0"] ** 2 * c["L"],
                         E_pot_f0=0.25 * c["Drho"] * c["k"] ** 4 * c["eta0"] ** 2 * c["L"],
                         E_ela_s0=0.25 * c["g"] * c["eta0"] ** 2 * c["L"])
        return out
