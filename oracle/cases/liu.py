"""Oracle restatement of `run_Liu` (`src/Liu.jl:18-135`): 2-D wave tank with variable
bathymetry on the triangular model `models/Liu.json`, one Euler-Bernoulli beam, damping zones at
both ends; P4 Lagrangian elements everywhere; complex (frequency domain).  Configuration 5-3-1
(`scripts/5-3-1-Liu.jl:20-32`: ω = 0.4 and 0.8).  Test infrastructure only.
"""
import numpy as np
from scipy.optimize import brentq
from ..model import json_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO, VEC


def constants(omega):
    """`src/Liu.jl:21-62`."""
    m, EI, H0, Lb = 500, 1.0e10, 60, 300.0
    g, rho = 9.81, 1025
    d0, a1 = m / rho, EI / rho
    f = lambda k: np.sqrt(g * k * np.tanh(k * H0)) - omega          # find_zero(f, 0.2) (:36-37)
    k = abs(brentq(f, 1e-8, 10.0, xtol=1e-15, rtol=8.9e-16))
    order = 4
    h = Lb / 50
    return dict(Lb=Lb, H0=H0, g=g, rho=rho, d0=d0, a1=a1, k=k, om=omega, eta0=0.01, order=order, h=h,
                gamma=1.0 * order * (order - 1) / h, bh=0.5, mu0=6.0, Ld=4 * Lb, xd_out=9 * Lb)


class LiuCase:
    def __init__(self, mesh_file, omega=0.2):
        self.p = dict(mesh_file=mesh_file, omega=omega)
        c = self.c = constants(omega)
        order = c["order"]
        self.model = m = json_model(mesh_file)
        self.Om = VolumeTriangulation(m)
        self.Gin = boundary(m, ["inlet"])
        self.Gb = boundary(m, ["beam"])
        self.Gd1 = boundary(m, ["damping_in"])
        self.Gd2 = boundary(m, ["damping_out"])
        self.Gf = boundary(m, ["free_surface"])
        self.Gk = boundary(m, ["free_surface", "damping_in", "damping_out"])
        self.Lb = SkeletonTriangulation(self.Gb)
        self.degree = 2 * order
        self.V_phi = FESpace(self.Om, order)
        self.V_kap = FESpace(self.Gk, order)
        self.V_eta = FESpace(self.Gb, order)
        self.X = MultiFieldFESpace([self.V_phi, self.V_kap, self.V_eta])

    # damping profiles (:55-60)
    def mu1_in(self, x1):
        return self.c["mu0"] * (1.0 - np.sin(np.pi / 2 * x1 / self.c["Ld"]))

    def mu1_out(self, x1):
        return self.c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - self.c["xd_out"]) / self.c["Ld"]))

    def assemble(self):
        c, order, deg = self.c, self.c["order"], self.degree
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        A = COO(N, np.complex128)
        b = VEC(N, np.complex128)
        g, om, bh, k = c["g"], c["om"], c["bh"], c["k"]
        ah = -1j * om / g * (1 - bh) / bh
        # Ω (:108)
        dOm = VolumeMeasure(self.Om, deg)
        o = dOm.ops(order)
        dofs = self.V_phi.cell_dofs + o_phi
        A.add_block(dofs, dofs, np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))
        # Γf, Γd1, Γd2 (:109-111)
        for trian, mu1f in [(self.Gf, None), (self.Gd1, self.mu1_in), (self.Gd2, self.mu1_out)]:
            dG = SurfaceMeasure(trian, deg)
            S = dG.surface_ops(order)
            T = dG.trace_ops(order)
            Dy = T.grad[..., 1]                       # ∇ₙ(ϕ) = ∇(ϕ)⋅(0,1) (:106)
            rows_w = self.V_phi.cell_dofs[trian.cell] + o_phi
            rows_u = self.V_kap.cell_dofs[self.Gk.position_of(trian)] + o_kap
            dV = S.dV
            m1 = mu1f(S.x[..., 0]) if mu1f is not None else np.zeros_like(dV)
            m2 = m1 * k
            A.add_block(rows_u, rows_u, bh * g * np.einsum("eq,eqi,eqj->eij", dV, S.val, S.val))
            A.add_block(rows_u, rows_w, -1j * om * bh * np.einsum("eq,eqi,eqj->eij", dV, S.val, T.val)
                        + np.einsum("eq,eqi,eqj->eij", dV * m1, S.val, Dy))
            A.add_block(rows_w, rows_u, np.einsum("eq,eqi,eqj->eij", dV * (bh * g * ah + 1j * om - m2), T.val, S.val))
            A.add_block(rows_w, rows_w, -1j * om * bh * ah * np.einsum("eq,eqi,eqj->eij", dV, T.val, T.val)
                        + ah * np.einsum("eq,eqi,eqj->eij", dV * m1, T.val, Dy))
            if trian is self.Gd1:                      # l (:114)
                x1 = S.x[..., 0]
                eta_d = m2 * c["eta0"] * np.exp(1j * k * x1)
                dphi_d = m1 * (-1j * om * c["eta0"] * np.exp(1j * k * x1))
                b.add(rows_w, np.einsum("eq,eqi->ei", dV * (-eta_d + ah * dphi_d), T.val))
                b.add(rows_u, np.einsum("eq,eqi->ei", dV * dphi_d, S.val))
        # Γb (:112)
        dG = SurfaceMeasure(self.Gb, deg)
        S = dG.surface_ops(order, need_hess=True)
        T = dG.trace_ops(order)
        rows_w = self.V_phi.cell_dofs[self.Gb.cell] + o_phi
        rows_v = self.V_eta.cell_dofs + o_eta
        A.add_block(rows_v, rows_v, (-om ** 2 * c["d0"] + g) * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
                    + c["a1"] * np.einsum("eq,eqi,eqj->eij", S.dV, S.lapl, S.lapl))
        A.add_block(rows_v, rows_w, -1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
        A.add_block(rows_w, rows_v, 1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        # Λb (:113)
        dL = SkeletonMeasure(self.Lb, deg)
        P = dL.side_ops(order, "plus")
        M = dL.side_ops(order, "minus")
        rows_p = self.V_eta.cell_dofs[self.Lb.plus] + o_eta
        rows_m = self.V_eta.cell_dofs[self.Lb.minus] + o_eta
        sides = [(P, rows_p), (M, rows_m)]
        for (Sv, rv) in sides:
            Jv = np.einsum("eqib,eqb->eqi", Sv.grad, Sv.normal)
            Mv = 0.5 * Sv.lapl
            for (Se, re_) in sides:
                Je = np.einsum("eqjb,eqb->eqj", Se.grad, Se.normal)
                Me = 0.5 * Se.lapl
                A.add_block(rv, re_, c["a1"] * (-np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Me)
                                                - np.einsum("eq,eqi,eqj->eij", P.dS, Mv, Je)
                                                + c["gamma"] * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je)))
        # Γin (:114): vᵢₙ (:42)
        dGin = SurfaceMeasure(self.Gin, deg)
        S = dGin.surface_ops(1)
        T = dGin.trace_ops(order)
        x = S.x
        vin = (c["eta0"] * om) * (np.cosh(k * (x[..., 1] + 0.075 * c["Lb"])) / np.sinh(k * c["H0"])) * np.exp(1j * k * x[..., 0])
        b.add(self.V_phi.cell_dofs[self.Gin.cell] + o_phi, np.einsum("eq,eqi->ei", S.dV * vin, T.val))
        self.ncoo = A.ncoo
        return A.tocsc(), b.v

    def postprocess(self, x):
        """(xs, |η|/η₀) per DOF node of every beam cell, each cell sorted with the permutation of
        the first cell (`src/Liu.jl:119-126`)."""
        c = self.c
        xc = self.V_eta.cell_dof_coords()[:, :, 0]
        p = np.argsort(xc[0], kind="stable")
        eta = x[self.X.offsets[2]:self.X.offsets[3]]
        vals = eta[self.V_eta.cell_dofs]
        return ((xc[:, p] - 6 * c["Lb"]) / c["Lb"]).reshape(-1), (np.abs(vals[:, p]) / c["eta0"]).reshape(-1)

    def solve(self):
        import scipy.sparse.linalg as spla
        A, b = self.assemble()
        x = spla.splu(A.tocsc()).solve(b)
        return x, self.postprocess(x)
