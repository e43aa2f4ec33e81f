"""Oracle restatement of `run_periodic_beam_FS` (`src/Periodic_Beam_FS.jl:22-213`).

2-D periodic fluid domain whose surface is split into an Euler-Bernoulli beam (x ∈ [π/2, 3π/2])
and a free surface; three fields (ϕ, κ on Γfs, η on Γb), time domain, Newmark(γ=½,β=¼), zero
right-hand side, analytic travelling wave as initial condition.  Configuration 5-1-4
(`scripts/5-1-4-periodic-beam-free-surface-energy.jl`).  Test infrastructure only.

One line of the reference is not restated literally: `ηₙ = x₀[2]` (`:189`) takes the free-surface
field κ as the "previous η" of the first step, a function that does not live on Γb; the first
E_kin_s sample here uses the interpolated η(0) on Γb, which is what every later step uses
(`:208`).
"""
import numpy as np
from ..model import cartesian_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO
from ..newmark import newmark, nsteps_for


def isapprox(a, b):
    return np.abs(a - b) <= np.sqrt(np.finfo(float).eps) * np.maximum(np.abs(a), np.abs(b))


def constants(n, dt, order, k):
    """`:27-61`."""
    L, H = 2.0 * np.pi, 1.0
    g, rho_w, rho_b, h_b = 9.81, 1.0e3, 1.0e2, 1.0e-2
    om = np.sqrt(g * k * np.tanh(k * H))
    EI_b = rho_b * h_b * om ** 2 / (k ** 4)
    gamma_t, beta_t = 0.5, 0.25
    bh = 0.5
    ah = gamma_t / (beta_t * dt) / g * (1 - bh) / bh
    h = L / n
    return dict(L=L, H=H, g=g, om=om, d0=rho_b * h_b / rho_w, Drho=EI_b / rho_w, eta0=0.01, h=h,
                gamma=10.0 * order * (order - 1) / h, k=k, gamma_t=gamma_t, beta_t=beta_t, bh=bh, ah=ah,
                Lb=np.pi, xb0=0.5 * np.pi, xb1=0.5 * np.pi + np.pi)


class PeriodicBeamFSCase:
    def __init__(self, n=4, dt=0.001, tf=1.0, order=2, k=10):
        self.p = dict(n=n, dt=dt, tf=tf, order=order, k=k)
        c = self.c = constants(n, dt, order, k)
        self.model = m = cartesian_model((0.0, c["L"], 0.0, c["H"]), (2 * n, n), isperiodic=(True, False))
        m.add_tag_from_entities("surface", [3, 4, 6])          # :76
        self.Om = VolumeTriangulation(m)
        self.G = G = boundary(m, "surface")
        xc = G.cell_coords().mean(axis=1)                       # is_beam (:84-88)
        beam = (c["xb0"] <= xc[:, 0]) & (xc[:, 0] <= c["xb1"]) & isapprox(xc[:, 1], c["H"])
        self.Gb = G.sub(np.nonzero(beam)[0])
        self.Gfs = G.sub(np.nonzero(~beam)[0])
        self.Lb = SkeletonTriangulation(self.Gb)
        self.degree = 2 * order
        self.V_phi = FESpace(self.Om, order)
        self.V_kap = FESpace(self.Gfs, order)
        self.V_eta = FESpace(self.Gb, order)
        self.X = MultiFieldFESpace([self.V_phi, self.V_kap, self.V_eta])

    def eta_ex(self, x, t, d=0):
        c = self.c
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [c["eta0"] * np.cos(ph), c["eta0"] * c["om"] * np.sin(ph), -c["eta0"] * c["om"] ** 2 * np.cos(ph)][d]

    def phi_ex(self, x, t, d=0):
        c = self.c
        amp = c["eta0"] * c["om"] / c["k"] * np.cosh(c["k"] * x[..., 1]) / np.sinh(c["k"] * c["H"])
        ph = c["k"] * x[..., 0] - c["om"] * t
        return [amp * np.sin(ph), -amp * c["om"] * np.cos(ph), -amp * c["om"] ** 2 * np.sin(ph)][d]

    def assemble(self):
        """K, C, M on the union pattern of the three jacobians (`:138-147`)."""
        c, order, deg = self.c, self.p["order"], self.degree
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        K, C, M = (COO(N, np.float64) for _ in range(3))

        def add(rows, cols, k=None, cc=None, mm=None):
            z = np.zeros((rows.shape[0], rows.shape[1], cols.shape[1]))
            K.add_block(rows, cols, z if k is None else k)
            C.add_block(rows, cols, z if cc is None else cc)
            M.add_block(rows, cols, z if mm is None else mm)

        g, bh, ah = c["g"], c["bh"], c["ah"]
        o = VolumeMeasure(self.Om, deg).ops(order)
        dofs = self.V_phi.cell_dofs + o_phi
        add(dofs, dofs, k=np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))            # ∇(w)⋅∇(ϕ)
        # Γfs: c = βₕ ϕₜ (u + αₕ w) − κₜ w ; a = βₕ (u + αₕ w) g κ
        dG = SurfaceMeasure(self.Gfs, deg)
        S, T = dG.surface_ops(order), dG.trace_ops(order)
        rows_w = self.V_phi.cell_dofs[self.Gfs.cell] + o_phi
        rows_u = self.V_kap.cell_dofs + o_kap
        uu = np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
        uw = np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val)
        wu = np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val)
        ww = np.einsum("eq,eqi,eqj->eij", S.dV, T.val, T.val)
        add(rows_u, rows_u, k=bh * g * uu)
        add(rows_u, rows_w, cc=bh * uw)
        add(rows_w, rows_u, k=bh * ah * g * wu, cc=-wu)
        add(rows_w, rows_w, cc=bh * ah * ww)
        # Γb: m = d₀ ηₜₜ v ; c = ϕₜ v − ηₜ w ; a = g η v + Dρ Δv Δη
        dG = SurfaceMeasure(self.Gb, deg)
        S, T = dG.surface_ops(order, need_hess=True), dG.trace_ops(order)
        rows_w = self.V_phi.cell_dofs[self.Gb.cell] + o_phi
        rows_v = self.V_eta.cell_dofs + o_eta
        mass = np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
        add(rows_v, rows_v, k=g * mass + c["Drho"] * np.einsum("eq,eqi,eqj->eij", S.dV, S.lapl, S.lapl),
            mm=c["d0"] * mass)
        add(rows_v, rows_w, cc=np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
        add(rows_w, rows_v, cc=-np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        # Λb
        dL = SkeletonMeasure(self.Lb, deg)
        P, Mi = dL.side_ops(order, "plus"), dL.side_ops(order, "minus")
        rows_p = self.V_eta.cell_dofs[self.Lb.plus] + o_eta
        rows_m = self.V_eta.cell_dofs[self.Lb.minus] + o_eta
        sides = [(P, rows_p), (Mi, rows_m)]
        for (Sv, rv) in sides:
            Jv = np.einsum("eqib,eqb->eqi", Sv.grad, Sv.normal)
            Mv = 0.5 * Sv.lapl
            for (Se, re_) in sides:
                Je = np.einsum("eqjb,eqb->eqj", Se.grad, Se.normal)
                Me = 0.5 * Se.lapl
                add(rv, re_, k=c["Drho"] * (-np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Me)
                                            - np.einsum("eq,eqi,eqj->eij", P.dS, Mv, Je)
                                            + c["gamma"] * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je)))
        return K.tocsr(), C.tocsr(), M.tocsr()

    def interpolate(self, t, d=0):
        """`interpolate_everywhere([ϕ(t),η(t),η(t)],X(t))` (`:154-156`)."""
        x = np.zeros(self.X.ndofs)
        o = self.X.offsets
        x[o[0]:o[1]] = self.phi_ex(self.V_phi.dof_coords(), t, d)
        x[o[1]:o[2]] = self.eta_ex(self.V_kap.dof_coords(), t, d)
        x[o[2]:o[3]] = self.eta_ex(self.V_eta.dof_coords(), t, d)
        return x

    def functionals(self):
        c, order, deg = self.c, self.p["order"], self.degree
        o = self.X.offsets
        dOm = VolumeMeasure(self.Om, deg).ops(order)
        dF = SurfaceMeasure(self.Gfs, deg).surface_ops(order)
        dB = SurfaceMeasure(self.Gb, deg).surface_ops(order, need_hess=True)
        cd_phi = self.V_phi.cell_dofs + o[0]
        cd_kap = self.V_kap.cell_dofs + o[1]
        cd_eta = self.V_eta.cell_dofs + o[2]

        def errors(x, t):
            uh = np.einsum("eqi,ei->eq", dOm.val, x[cd_phi])
            e_phi = np.sqrt(np.sum(dOm.dV * (self.phi_ex(dOm.x, t) - uh) ** 2))
            kh = np.einsum("eqi,ei->eq", dF.val, x[cd_kap])
            eh = np.einsum("eqi,ei->eq", dB.val, x[cd_eta])
            e_eta = (np.sqrt(np.sum(dF.dV * (self.eta_ex(dF.x, t) - kh) ** 2))
                     + np.sqrt(np.sum(dB.dV * (self.eta_ex(dB.x, t) - eh) ** 2)))
            return e_phi, e_eta

        def energies(x, xprev, dt):
            gphi = np.einsum("eqib,ei->eqb", dOm.grad, x[cd_phi])
            kap = np.einsum("eqi,ei->eq", dF.val, x[cd_kap])
            eta = np.einsum("eqi,ei->eq", dB.val, x[cd_eta])
            eta_t = np.einsum("eqi,ei->eq", dB.val, (x - xprev)[cd_eta]) / dt
            lap = np.einsum("eqi,ei->eq", dB.lapl, x[cd_eta])
            return (0.5 * np.sum(dOm.dV * np.sum(gphi * gphi, axis=-1)),
                    0.5 * c["g"] * np.sum(dF.dV * kap * kap) + 0.5 * c["g"] * np.sum(dB.dV * eta * eta),
                    0.5 * c["d0"] * np.sum(dB.dV * eta_t * eta_t),
                    0.5 * c["Drho"] * np.sum(dB.dV * lap * lap))
        return errors, energies

    def run(self, keep_states=False):
        c, p = self.c, self.p
        K, C, M = self.assemble()
        nst = nsteps_for(0.0, p["tf"], p["dt"])
        errors, energies = self.functionals()
        out = dict(t=[], e_phi=[], e_eta=[], E=[], states=[])
        prev = [self.interpolate(0.0, 0)]

        def cb(n, t, u, v, a):
            e1, e2 = errors(u, t)
            out["t"].append(t); out["e_phi"].append(e1); out["e_eta"].append(e2)
            out["E"].append(energies(u, prev[0], p["dt"]))
            prev[0] = u.copy()
            if keep_states:
                out["states"].append(u.copy())

        u, v, a = newmark(K, C, M, None, self.interpolate(0.0, 0), self.interpolate(0.0, 1),
                          self.interpolate(0.0, 2), p["dt"], c["gamma_t"], c["beta_t"], nst, callback=cb)
        out["u"], out["v"], out["a"] = u, v, a
        # analytic energy constants (`:177-180`, names as in the reference)
        out["E0"] = dict(E_kin_s0=0.25 * c["d0"] * c["om"] ** 2 * c["eta0"] ** 2 * c["Lb"],
                         E_kin_f0=0.25 * c["g"] * c["eta0"] ** 2 * c["L"],
                         E_ela_s0=0.25 * c["Drho"] * c["k"] ** 4 * c["eta0"] ** 2 * c["Lb"],
                         E_pot_f0=0.25 * c["g"] * c["eta0"] ** 2 * c["L"])
        return out
