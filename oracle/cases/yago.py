"""Oracle restatement of `run_Yago_freq_domain` (`src/Yago_freq_domain.jl:22-209`).

3-D frequency-domain Kirchhoff plate on the free surface of a finite-depth
fluid: hex mesh, Q2 potential, Q`order` free-surface elevation and plate
deflection, complex arithmetic.  Test infrastructure only.
"""
import numpy as np
from ..model import cartesian_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO, VEC


def constants(nx, ny, nz, order, lfactor, dfactor):
    """`src/Yago_freq_domain.jl:29-94`."""
    c = {}
    L, B, H, hb = 300, 60, 58.5, 2.0
    nLO, nBO = 10, 18
    LO, BO = nLO * L, nBO * B
    xb0 = 4.5 * L
    xb1 = xb0 + L
    yb0 = -B / 2
    yb1 = yb0 + B
    g, rho, E, rhob, nu = 9.81, 1025, 11.9e9, 256.25, 0.13
    I = hb ** 3 / 12
    D = E * I / (1.0 - nu ** 2)
    d0 = rhob * hb / rho
    mu = E / (2 * (1 + nu))
    lam = nu * E / (1 - nu ** 2)
    C = np.zeros((3, 3, 3, 3))
    dl = lambda a, b: 1.0 if a == b else 0.0
    for i in range(2):
        for j in range(2):
            for k in range(2):
                for l in range(2):
                    C[i, j, k, l] = I / rho * (mu * (dl(i, k) * dl(j, l) + dl(i, l) * dl(j, k))
                                               + lam * dl(i, j) * dl(k, l))
    wl = L * lfactor
    k = 2 * np.pi / wl
    om = np.sqrt(g * k * np.tanh(k * H))
    eta0 = 0.01
    h = LO / (nLO * nx)
    bh = 0.5
    ah = -1j * om / g * (1 - bh) / bh
    gamma = 1.0 * order * (order + 1) / h
    mu0 = 2.5
    Ld = dfactor * L
    Ld0 = dfactor * L
    xd = LO - Ld
    xd0 = Ld0
    c.update(L=L, B=B, H=H, LO=LO, BO=BO, xb0=xb0, xb1=xb1, yb0=yb0, yb1=yb1, g=g, rho=rho,
             D=D, d0=d0, C=C, k=k, om=om, eta0=eta0, h=h, bh=bh, ah=ah, gamma=gamma,
             mu0=mu0, Ld=Ld, Ld0=Ld0, xd=xd, xd0=xd0, nLO=nLO, nBO=nBO)
    return c


def mu1(c, x1):
    """`src/Yago_freq_domain.jl:91`."""
    return (c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - c["xd"]) / c["Ld"])) * (x1 > c["xd"])
            + c["mu0"] * (1 - np.cos(np.pi / 2 * (c["Ld0"] - x1) / c["Ld0"])) * (x1 < c["xd0"]))


class YagoCase:
    def __init__(self, nx=32, ny=4, nz=4, order=4, lfactor=0.4, dfactor=3):
        self.p = dict(nx=nx, ny=ny, nz=nz, order=order, lfactor=lfactor, dfactor=dfactor)
        c = self.c = constants(nx, ny, nz, order, lfactor, dfactor)
        H = c["H"]

        def f_z(x):                                   # :100-106
            if x == H:
                return H
            i = x / (H / nz)
            return H - H / (2.5 ** i)

        domain = (0.0, c["LO"], -c["BO"] / 2, c["BO"] / 2, 0.0, H)
        partition = (c["nLO"] * nx, c["nBO"] * ny, nz)
        self.model = m = cartesian_model(domain, partition,
                                         map_fn=lambda x: (x[0], x[1], f_z(x[2])))
        m.add_tag_from_entities("surface", [22])      # :112
        m.add_tag_from_entities("inlet", [25])        # :113
        self.Om = VolumeTriangulation(m)
        self.G = boundary(m, "surface")
        self.Gin = boundary(m, "inlet")
        xG = self.G.cell_coords()
        is_plate = np.all((c["xb0"] <= xG[:, :, 0]) & (xG[:, :, 0] <= c["xb1"])
                          & (c["yb0"] <= xG[:, :, 1]) & (xG[:, :, 1] <= c["yb1"]), axis=1)
        self.Gb = self.G.sub(np.nonzero(is_plate)[0])
        self.Gf = self.G.sub(np.nonzero(~is_plate)[0])
        self.Lb = SkeletonTriangulation(self.Gb)
        self.degree = 2 * order
        self.V_phi = FESpace(self.Om, 2)
        self.V_kap = FESpace(self.Gf, order)
        self.V_eta = FESpace(self.Gb, order)
        self.X = MultiFieldFESpace([self.V_phi, self.V_kap, self.V_eta])

    # ------------------------------------------------------------------
    def assemble(self, chunk=2048):
        c, order, deg = self.c, self.p["order"], self.degree
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        A = COO(N, np.complex128)
        b = VEC(N, np.complex128)
        g, om, bh, ah, k = c["g"], c["om"], c["bh"], c["ah"], c["k"]

        # Ω: ∫ ∇ϕ·∇w   (:162)
        dOm = VolumeMeasure(self.Om, deg)
        dofs = self.V_phi.cell_dofs + o_phi
        for s in range(0, self.Om.ncells, chunk):
            sl = slice(s, min(s + chunk, self.Om.ncells))
            op = dOm.ops(2, cells=sl)
            K = np.einsum("eq,eqib,eqjb->eij", op.dV, op.grad, op.grad)
            A.add_block(dofs[sl], dofs[sl], K)

        # Γf (:163)
        dGf = SurfaceMeasure(self.Gf, deg)
        S = dGf.surface_ops(order)
        T = dGf.trace_ops(2)
        Dz = T.grad[..., 2]
        m1 = mu1(c, S.x[..., 0])
        m2 = m1 * k
        rows_w = self.V_phi.cell_dofs[self.Gf.cell] + o_phi
        rows_u = self.V_kap.cell_dofs + o_kap
        dV = S.dV
        A.add_block(rows_u, rows_u, bh * g * np.einsum("eq,eqi,eqj->eij", dV, S.val, S.val))
        A.add_block(rows_u, rows_w,
                    np.einsum("eq,eqi,eqj->eij", dV, S.val, -1j * om * bh * T.val)
                    + np.einsum("eq,eqi,eqj->eij", dV * m1, S.val, Dz))
        A.add_block(rows_w, rows_u,
                    np.einsum("eq,eqi,eqj->eij", dV * (bh * g * ah + 1j * om - m2), T.val, S.val))
        A.add_block(rows_w, rows_w,
                    np.einsum("eq,eqi,eqj->eij", dV, T.val, -1j * om * bh * ah * T.val)
                    + np.einsum("eq,eqi,eqj->eij", dV * m1 * ah, T.val, Dz))
        # RHS on Γf (:166): -∫(ηd w - ∇ₙϕd (u + αₕ w))
        x1 = S.x[..., 0]
        eta_in = c["eta0"] * np.exp(1j * k * x1)
        vz_in = -1j * om * c["eta0"] * np.exp(1j * k * x1)
        eta_d = m2 * eta_in * (x1 < c["xd0"])
        dphi_d = m1 * vz_in * (x1 < c["xd0"])
        b.add(rows_w, np.einsum("eq,eqi->ei", dV * (-eta_d + ah * dphi_d), T.val))
        b.add(rows_u, np.einsum("eq,eqi->ei", dV * dphi_d, S.val))

        # Γb (:164)
        dGb = SurfaceMeasure(self.Gb, deg)
        S = dGb.surface_ops(order, need_hess=True)
        T = dGb.trace_ops(2)
        rows_w = self.V_phi.cell_dofs[self.Gb.cell] + o_phi
        rows_v = self.V_eta.cell_dofs + o_eta
        CH = np.einsum("abcd,eqicd->eqiab", c["C"], S.hess)
        A.add_block(rows_v, rows_v,
                    (g - om ** 2 * c["d0"]) * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
                    + np.einsum("eq,eqiab,eqjab->eij", S.dV, S.hess, CH))
        A.add_block(rows_v, rows_w, -1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
        A.add_block(rows_w, rows_v, 1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))

        # Λb (:165)
        dLb = SkeletonMeasure(self.Lb, deg)
        P = dLb.side_ops(order, "plus")
        M = dLb.side_ops(order, "minus")
        npl = P.normal
        rows_p = self.V_eta.cell_dofs[self.Lb.plus] + o_eta
        rows_m = self.V_eta.cell_dofs[self.Lb.minus] + o_eta
        pen = c["D"] / c["rho"] * c["gamma"]
        sides = [(P, rows_p, +1.0), (M, rows_m, -1.0)]
        for (Sv, rv, sv) in sides:            # test side
            Jv = sv * Sv.grad                                            # jump(∇v)
            Mv = 0.5 * np.einsum("abcd,eqicd,eqb->eqia", c["C"], Sv.hess, npl)   # mean(C⊙∇∇v)·n⁺
            for (Se, re_, se) in sides:       # trial side
                Je = se * Se.grad
                Me = 0.5 * np.einsum("abcd,eqjcd,eqb->eqja", c["C"], Se.hess, npl)
                blk = (-np.einsum("eq,eqia,eqja->eij", P.dS, Jv, Me)
                       - np.einsum("eq,eqia,eqja->eij", P.dS, Mv, Je)
                       + pen * np.einsum("eq,eqia,eqja->eij", P.dS, Jv, Je))
                A.add_block(rv, re_, blk)

        # Γin RHS (:166): ∫ vin w
        dGin = SurfaceMeasure(self.Gin, deg)
        S = dGin.surface_ops(1)
        T = dGin.trace_ops(2)
        x = S.x
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 2]) / np.sinh(k * c["H"])) * np.exp(1j * k * x[..., 0])
        rows_w = self.V_phi.cell_dofs[self.Gin.cell] + o_phi
        b.add(rows_w, np.einsum("eq,eqi->ei", S.dV * vin, T.val))

        self.ncoo = A.ncoo
        return A.tocsc(), b.v

    # ------------------------------------------------------------------
    def rao(self, x):
        """`src/Yago_freq_domain.jl:178-199`: |η|/η₀ at plate-cell vertices on y=0."""
        c = self.c
        eta = x[self.X.offsets[2]:self.X.offsets[3]]
        xe = self.Gb.cell_coords()
        xps, vals = [], []
        for ie in range(self.Gb.ncells):
            for iv in range(xe.shape[1]):
                xi = xe[ie, iv]
                s = -2 * (xi[0] - (c["xb0"] + c["L"] / 2)) / c["L"]
                if -1 <= s <= 1 and xi[1] == 0.0:
                    xps.append(s)
                    vals.append(abs(eta[self.V_eta.cell_dofs[ie, iv]]) / c["eta0"])
        return np.array(xps), np.array(vals)
