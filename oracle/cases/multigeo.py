"""Oracle restatement of `run_MultiGeo_freq_domain` (`src/Multi_geo_freq_domain.jl:19-151`).

3-D frequency-domain Kirchhoff plates of arbitrary shape on the free surface of a tank
meshed with tetrahedra (gmsh): P2 potential, P`order` free-surface elevation and plate
deflection, cell-wise penalty length h = skeleton edge length.  Simplex quadrature: the
oracle and the product both use the Duffy-collapsed Gauss-Jacobi rule of SURVEY.md A13 (the
non-polynomial damping coefficients make the entries rule-dependent; parity is defined
against this stated rule).  Test infrastructure only.
"""
import numpy as np
from ..model import json_model
from ..gmsh import gmsh_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO, VEC


def constants(order):
    """`src/Multi_geo_freq_domain.jl:26-96`."""
    g, rho, E, rhob, nu, hb = 9.81, 1025, 11.9e9, 256.25, 0.13, 0.1
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
    H = 3
    k = 2 * np.pi / 1.0
    om = np.sqrt(g * k * np.tanh(k * H))
    bh = 0.5
    ah = -1j * om / g * (1 - bh) / bh
    return dict(g=g, rho=rho, D=D, d0=d0, C=C, H=H, k=k, om=om, eta0=0.01, bh=bh, ah=ah,
                mu0=2.5, Ld=3, Ld0=3, xd=10, xd0=-1)


def mu1(c, x1):
    """`src/Multi_geo_freq_domain.jl:93`."""
    return (c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - c["xd"]) / c["Ld"])) * (x1 > c["xd"])
            + c["mu0"] * (1 - np.cos(np.pi / 2 * (c["Ld0"] - x1) / c["Ld0"])) * (x1 < c["xd0"]))


class MultiGeoCase:
    def __init__(self, mesh_file, order=4):
        self.p = dict(mesh_file=mesh_file, order=order)
        c = self.c = constants(order)
        self.model = m = gmsh_model(mesh_file) if mesh_file.endswith(".msh") else json_model(mesh_file)
        self.Om = VolumeTriangulation(m)
        self.Gin = boundary(m, ["inlet"])
        self.Gb = boundary(m, ["plate"])
        self.Gf = boundary(m, ["freeSurface"])
        self.Lb = SkeletonTriangulation(self.Gb)
        self.degree = 2 * order
        self.V_phi = FESpace(self.Om, 2)
        self.V_kap = FESpace(self.Gf, order)
        self.V_eta = FESpace(self.Gb, order)
        self.X = MultiFieldFESpace([self.V_phi, self.V_kap, self.V_eta])

    def edge_lengths(self):
        """`get_cell_measure(Λb)` (`:82`)."""
        sk = self.Lb
        X = sk.surf.cell_coords()[sk.plus]
        lv = np.array(sk.surf.poly.faces[1], dtype=np.int64)[sk.plus_lf]
        a = X[np.arange(len(X)), lv[:, 0]]
        b = X[np.arange(len(X)), lv[:, 1]]
        return np.linalg.norm(b - a, axis=1)

    def assemble(self, chunk=4096):
        c, order, deg = self.c, self.p["order"], self.degree
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        A = COO(N, np.complex128)
        b = VEC(N, np.complex128)
        g, om, bh, ah, k = c["g"], c["om"], c["bh"], c["ah"], c["k"]
        dOm = VolumeMeasure(self.Om, deg)
        dofs = self.V_phi.cell_dofs + o_phi
        for s in range(0, self.Om.ncells, chunk):
            sl = slice(s, min(s + chunk, self.Om.ncells))
            o = dOm.ops(2, cells=sl)
            A.add_block(dofs[sl], dofs[sl], np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))
        # Γf (:127)
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
        A.add_block(rows_u, rows_w, np.einsum("eq,eqi,eqj->eij", dV, S.val, -1j * om * bh * T.val)
                    + np.einsum("eq,eqi,eqj->eij", dV * m1, S.val, Dz))
        A.add_block(rows_w, rows_u, np.einsum("eq,eqi,eqj->eij", dV * (bh * g * ah + 1j * om - m2), T.val, S.val))
        A.add_block(rows_w, rows_w, np.einsum("eq,eqi,eqj->eij", dV, T.val, -1j * om * bh * ah * T.val)
                    + np.einsum("eq,eqi,eqj->eij", dV * m1 * ah, T.val, Dz))
        x1 = S.x[..., 0]
        eta_in = c["eta0"] * np.exp(1j * k * x1)
        vz_in = -1j * om * c["eta0"] * np.exp(1j * k * x1)
        eta_d = m2 * eta_in * (x1 < c["xd0"])
        dphi_d = m1 * vz_in * (x1 < c["xd0"])
        b.add(rows_w, np.einsum("eq,eqi->ei", dV * (-eta_d + ah * dphi_d), T.val))
        b.add(rows_u, np.einsum("eq,eqi->ei", dV * dphi_d, S.val))
        # Γb (:128)
        dGb = SurfaceMeasure(self.Gb, deg)
        S = dGb.surface_ops(order, need_hess=True)
        T = dGb.trace_ops(2)
        rows_w = self.V_phi.cell_dofs[self.Gb.cell] + o_phi
        rows_v = self.V_eta.cell_dofs + o_eta
        CH = np.einsum("abcd,eqicd->eqiab", c["C"], S.hess)
        A.add_block(rows_v, rows_v, (g - om ** 2 * c["d0"]) * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
                    + np.einsum("eq,eqiab,eqjab->eij", S.dV, S.hess, CH))
        A.add_block(rows_v, rows_w, -1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
        A.add_block(rows_w, rows_v, 1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        # Λb (:129): γ = order(order+1)/h with h the length of each edge
        dLb = SkeletonMeasure(self.Lb, deg)
        P = dLb.side_ops(order, "plus")
        M = dLb.side_ops(order, "minus")
        npl = P.normal
        rows_p = self.V_eta.cell_dofs[self.Lb.plus] + o_eta
        rows_m = self.V_eta.cell_dofs[self.Lb.minus] + o_eta
        pen = (c["D"] / c["rho"] * order * (order + 1) / self.edge_lengths())[:, None]
        sides = [(P, rows_p, +1.0), (M, rows_m, -1.0)]
        for (Sv, rv, sv) in sides:
            Jv = sv * Sv.grad
            Mv = 0.5 * np.einsum("abcd,eqicd,eqb->eqia", c["C"], Sv.hess, npl)
            for (Se, re_, se) in sides:
                Je = se * Se.grad
                Me = 0.5 * np.einsum("abcd,eqjcd,eqb->eqja", c["C"], Se.hess, npl)
                blk = (-np.einsum("eq,eqia,eqja->eij", P.dS, Jv, Me)
                       - np.einsum("eq,eqia,eqja->eij", P.dS, Mv, Je)
                       + np.einsum("eq,eqia,eqja->eij", P.dS * pen, Jv, Je))
                A.add_block(rv, re_, blk)
        # Γin rhs (:130)
        dGin = SurfaceMeasure(self.Gin, deg)
        S = dGin.surface_ops(1)
        T = dGin.trace_ops(2)
        x = S.x
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 2]) / np.sinh(k * c["H"])) * np.exp(1j * k * x[..., 0])
        b.add(self.V_phi.cell_dofs[self.Gin.cell] + o_phi, np.einsum("eq,eqi->ei", S.dV * vin, T.val))
        self.ncoo = A.ncoo
        return A.tocsc(), b.v
