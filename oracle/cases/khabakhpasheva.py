"""Oracle restatement of `run_Khabakhpasheva_freq_domain`
(`src/Khabakhpasheva_freq_domain.jl:22-261`) and `run_Khabakhpasheva_time_domain`
(`src/Khabakhpasheva_time_domain.jl:23-306`).

2-D numerical wave tank, two Euler-Bernoulli beams joined by a rotational spring, damping
zones at both ends; Q`order` everywhere; complex (frequency domain) or real + Newmark (time
domain).  Test infrastructure only.
"""
import numpy as np
from ..model import cartesian_model
from ..fem import (VolumeTriangulation, boundary, SkeletonTriangulation, FESpace,
                   MultiFieldFESpace, VolumeMeasure, SurfaceMeasure, SkeletonMeasure)
from ..assembly import COO, VEC
from ..newmark import newmark, nsteps_for


def isapprox(a, b):
    """Julia `a ≈ b` for Float64: |a-b| <= sqrt(eps)*max(|a|,|b|)."""
    return np.abs(a - b) <= np.sqrt(np.finfo(float).eps) * np.maximum(np.abs(a), np.abs(b))


def constants(nx, ny, order, xi):
    """`src/Khabakhpasheva_freq_domain.jl:28-91` (identical in the time-domain file :29-101)."""
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
    d0 = m / rho
    a1, a2 = EI1 / rho, EI2 / rho
    kr = xi * a1 / Lb
    lam = alpha * Lb
    k = 2 * np.pi / lam
    om = np.sqrt(g * k * np.tanh(k * H))
    T = 2 * np.pi / om
    eta0 = 0.01
    nx_total = int(np.ceil(nx / beta) * np.ceil(LO / Lb))
    h = LO / nx_total
    gamma = 1.0 * order * (order - 1) / h
    bh = 0.5
    mu0 = 2.5
    return dict(Lb=Lb, H=H, Ld=Ld, LO=LO, x0=x0, xd_in=xd_in, xb0=xb0, xbj=xbj, xb1=xb1,
                xd_out=xd_out, g=g, rho=rho, d0=d0, a1=a1, a2=a2, kr=kr, k=k, om=om, T=T,
                eta0=eta0, nx_total=nx_total, h=h, gamma=gamma, bh=bh, mu0=mu0)


class KhabakhpashevaGeometry:
    """Mesh, triangulations and FE spaces shared by the two drivers
    (`src/Khabakhpasheva_freq_domain.jl:93-224`)."""

    def __init__(self, nx=20, ny=5, order=4, xi=0.0):
        self.p = dict(nx=nx, ny=ny, order=order, xi=xi)
        c = self.c = constants(nx, ny, order, xi)
        H = c["H"]

        def f_y(y):                                    # :96-102
            if y == H:
                return H
            i = y / (H / ny)
            return H - H / (2.5 ** i)

        self.model = m = cartesian_model((c["x0"], c["LO"], 0.0, H), (c["nx_total"], ny),
                                         map_fn=lambda x: (x[0], f_y(x[1])))
        m.add_tag_from_entities("surface", [3, 4, 6])
        m.add_tag_from_entities("inlet", [7])
        self.Om = VolumeTriangulation(m)
        self.G = G = boundary(m, "surface")
        self.Gin = boundary(m, "inlet")
        xG = G.cell_coords()
        xc = xG.mean(axis=1)                           # centroid (:121-141)
        top = isapprox(xc[:, 1], H)
        b1 = (c["xb0"] <= xc[:, 0]) & (xc[:, 0] <= c["xbj"]) & top
        b2 = (c["xbj"] <= xc[:, 0]) & (xc[:, 0] <= c["xb1"]) & top
        d1 = (c["x0"] <= xc[:, 0]) & (xc[:, 0] <= c["xd_in"]) & top
        d2 = (c["xd_out"] <= xc[:, 0]) & top
        ids = lambda mask: np.nonzero(mask)[0]
        self.Gb1, self.Gb2 = G.sub(ids(b1)), G.sub(ids(b2))
        self.Gd1, self.Gd2 = G.sub(ids(d1)), G.sub(ids(d2))
        self.Gfs = G.sub(ids(~(b1 | b2 | d1 | d2)))
        self.Geta = G.sub(ids(b1 | b2))
        self.Gkap = G.sub(ids(~(b1 | b2)))
        self.Lb1 = SkeletonTriangulation(self.Gb1)
        self.Lb2 = SkeletonTriangulation(self.Gb2)
        # joint: surface vertices (ascending parent id) at (xbj, H)  (:173-181)
        act = G.active_model()
        xv = m.node_coords[act["parent"][0]]   # vertex == node on non-periodic Cartesian models
        jmask = isapprox(xv[:, 0], c["xbj"]) & isapprox(xv[:, 1], H)
        self.Lj = SkeletonTriangulation(G, face_mask=jmask)
        self.degree = 2 * order
        self.V_phi = FESpace(self.Om, order)
        self.V_kap = FESpace(self.Gkap, order)
        self.V_eta = FESpace(self.Geta, order)
        self.X = MultiFieldFESpace([self.V_phi, self.V_kap, self.V_eta])

    # damping profiles (:85-91)
    def mu1_in(self, x1):
        return self.c["mu0"] * (1.0 - np.sin(np.pi / 2 * x1 / self.c["Ld"]))

    def mu1_out(self, x1):
        return self.c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - self.c["xd_out"]) / self.c["Ld"]))

    def _rows(self, trian):
        """Global DOF rows (w, u, v) of the surface cells of `trian`."""
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        w = self.V_phi.cell_dofs[trian.cell] + o_phi
        return w, o_kap, o_eta

    def eta_postprocess(self, x):
        """(xs, |η|/η₀) per DOF node of every beam cell, each cell sorted with the permutation
        of the first cell (`src/Khabakhpasheva_freq_domain.jl:250-257`)."""
        c = self.c
        xc = self.V_eta.cell_dof_coords()[:, :, 0]
        p = np.argsort(xc[0], kind="stable")
        eta = x[self.X.offsets[2]:self.X.offsets[3]]
        vals = eta[self.V_eta.cell_dofs]
        return ((xc[:, p] - c["xb0"]) / c["Lb"]).reshape(-1), (np.abs(vals[:, p]) / c["eta0"]).reshape(-1)


def _skeleton_blocks(add, case, skel, space_off, coef, gamma, order, deg, mode):
    """Λ terms: `mode='beam'`  a(−⟦∇v·n⟧{Δη} − {Δv}⟦∇η·n⟧ + γ⟦∇v·n⟧⟦∇η·n⟧)  (:234-235),
    `mode='joint'`  ⟦∇v·n⟧ kᵣ ⟦∇η·n⟧  (:236)."""
    if skel.nfaces == 0:
        return
    dL = SkeletonMeasure(skel, deg)
    P = dL.side_ops(order, "plus")
    M = dL.side_ops(order, "minus")
    pos = case.Geta.position_of(skel.surf) if skel.surf is not case.Geta else np.arange(skel.surf.ncells)
    rows_p = case.V_eta.cell_dofs[pos[skel.plus]] + space_off
    rows_m = case.V_eta.cell_dofs[pos[skel.minus]] + space_off
    sides = [(P, rows_p), (M, rows_m)]
    for (Sv, rv) in sides:
        Jv = np.einsum("eqib,eqb->eqi", Sv.grad, Sv.normal)
        Mv = 0.5 * Sv.lapl
        for (Se, re_) in sides:
            Je = np.einsum("eqjb,eqb->eqj", Se.grad, Se.normal)
            Me = 0.5 * Se.lapl
            if mode == "beam":
                blk = coef * (-np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Me)
                              - np.einsum("eq,eqi,eqj->eij", P.dS, Mv, Je)
                              + gamma * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je))
            else:
                blk = coef * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je)
            add(rv, re_, blk)


class KhabakhpashevaFreqCase(KhabakhpashevaGeometry):
    def assemble(self):
        c, order, deg = self.c, self.p["order"], self.degree
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        A = COO(N, np.complex128)
        b = VEC(N, np.complex128)
        g, om, bh, k = c["g"], c["om"], c["bh"], c["k"]
        ah = -1j * om / g * (1 - bh) / bh
        # Ω
        dOm = VolumeMeasure(self.Om, deg)
        o = dOm.ops(order)
        dofs = self.V_phi.cell_dofs + o_phi
        A.add_block(dofs, dofs, np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))
        # free-surface type domains (:229-231)
        for trian, mu1f in [(self.Gfs, None), (self.Gd1, self.mu1_in), (self.Gd2, self.mu1_out)]:
            if trian.ncells == 0:
                continue
            dG = SurfaceMeasure(trian, deg)
            S = dG.surface_ops(order)
            T = dG.trace_ops(order)
            Dy = T.grad[..., 1]
            rows_w = self.V_phi.cell_dofs[trian.cell] + o_phi
            rows_u = self.V_kap.cell_dofs[self.Gkap.position_of(trian)] + o_kap
            dV = S.dV
            m1 = mu1f(S.x[..., 0]) if mu1f is not None else np.zeros_like(dV)
            m2 = m1 * k
            A.add_block(rows_u, rows_u, bh * g * np.einsum("eq,eqi,eqj->eij", dV, S.val, S.val))
            A.add_block(rows_u, rows_w, -1j * om * bh * np.einsum("eq,eqi,eqj->eij", dV, S.val, T.val)
                        + np.einsum("eq,eqi,eqj->eij", dV * m1, S.val, Dy))
            A.add_block(rows_w, rows_u, np.einsum("eq,eqi,eqj->eij", dV * (bh * g * ah + 1j * om - m2), T.val, S.val))
            A.add_block(rows_w, rows_w, -1j * om * bh * ah * np.einsum("eq,eqi,eqj->eij", dV, T.val, T.val)
                        + ah * np.einsum("eq,eqi,eqj->eij", dV * m1, T.val, Dy))
            if trian is self.Gd1:                      # rhs (:237)
                x1 = S.x[..., 0]
                eta_in = c["eta0"] * np.exp(1j * k * x1)
                vz_in = -1j * om * c["eta0"] * np.exp(1j * k * x1)
                eta_d = m2 * eta_in
                dphi_d = m1 * vz_in
                b.add(rows_w, np.einsum("eq,eqi->ei", dV * (-eta_d + ah * dphi_d), T.val))
                b.add(rows_u, np.einsum("eq,eqi->ei", dV * dphi_d, S.val))
        # beams (:232-233)
        for trian, ai in [(self.Gb1, c["a1"]), (self.Gb2, c["a2"])]:
            dG = SurfaceMeasure(trian, deg)
            S = dG.surface_ops(order, need_hess=True)
            T = dG.trace_ops(order)
            rows_w = self.V_phi.cell_dofs[trian.cell] + o_phi
            rows_v = self.V_eta.cell_dofs[self.Geta.position_of(trian)] + o_eta
            A.add_block(rows_v, rows_v,
                        (-om ** 2 * c["d0"] + g) * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
                        + ai * np.einsum("eq,eqi,eqj->eij", S.dV, S.lapl, S.lapl))
            A.add_block(rows_v, rows_w, -1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
            A.add_block(rows_w, rows_v, 1j * om * np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        # skeletons (:234-236)
        _skeleton_blocks(A.add_block, self, self.Lb1, o_eta, c["a1"], c["gamma"], order, deg, "beam")
        _skeleton_blocks(A.add_block, self, self.Lb2, o_eta, c["a2"], c["gamma"], order, deg, "beam")
        self._joint(A.add_block, o_eta, c["kr"], order, deg)
        # Γin rhs
        dGin = SurfaceMeasure(self.Gin, deg)
        S = dGin.surface_ops(1)
        T = dGin.trace_ops(order)
        x = S.x
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 1]) / np.sinh(k * c["H"])) * np.exp(1j * k * x[..., 0])
        b.add(self.V_phi.cell_dofs[self.Gin.cell] + o_phi, np.einsum("eq,eqi->ei", S.dV * vin, T.val))
        self.ncoo = A.ncoo
        return A.tocsc(), b.v

    def _joint(self, add, o_eta, kr, order, deg):
        """Λj lives on Skeleton(Γ,mask): its plus/minus cells are cells of Γ; η is looked up
        through the common background facet."""
        skel = self.Lj
        if skel.nfaces == 0:
            return
        dL = SkeletonMeasure(skel, deg)
        P = dL.side_ops(order, "plus")
        M = dL.side_ops(order, "minus")
        lut = {int(f): i for i, f in enumerate(self.Geta.facets)}
        pos = np.array([lut[int(f)] for f in self.G.facets[np.concatenate([skel.plus, skel.minus])]])
        rows_p = self.V_eta.cell_dofs[pos[:skel.nfaces]] + o_eta
        rows_m = self.V_eta.cell_dofs[pos[skel.nfaces:]] + o_eta
        sides = [(P, rows_p), (M, rows_m)]
        for (Sv, rv) in sides:
            Jv = np.einsum("eqib,eqb->eqi", Sv.grad, Sv.normal)
            for (Se, re_) in sides:
                Je = np.einsum("eqjb,eqb->eqj", Se.grad, Se.normal)
                add(rv, re_, kr * np.einsum("eq,eqi,eqj->eij", P.dS, Jv, Je))


class KhabakhpashevaTimeCase(KhabakhpashevaGeometry):
    """`src/Khabakhpasheva_time_domain.jl:237-266`: three real matrices + time-dependent rhs."""

    def time_constants(self):
        c = self.c
        gt, bt = 0.5, 0.25
        dt = c["T"] / 40
        tf = 50 * c["T"]
        cu = gt / (bt * dt)
        ah = cu / c["g"] * (1 - c["bh"]) / c["bh"]
        return dict(gamma_t=gt, beta_t=bt, dt=dt, tf=tf, ah=ah)

    def assemble(self):
        c, order, deg = self.c, self.p["order"], self.degree
        tc = self.time_constants()
        o_phi, o_kap, o_eta = self.X.offsets[:3]
        N = self.X.ndofs
        K, C, M = (COO(N, np.float64) for _ in range(3))
        g, bh, k, ah = c["g"], c["bh"], c["k"], tc["ah"]

        def add(rows, cols, kk=None, cc=None, mm=None):
            z = np.zeros((rows.shape[0], rows.shape[1], cols.shape[1]))
            K.add_block(rows, cols, z if kk is None else kk)
            C.add_block(rows, cols, z if cc is None else cc)
            M.add_block(rows, cols, z if mm is None else mm)

        dOm = VolumeMeasure(self.Om, deg)
        o = dOm.ops(order)
        dofs = self.V_phi.cell_dofs + o_phi
        add(dofs, dofs, kk=np.einsum("eq,eqib,eqjb->eij", o.dV, o.grad, o.grad))
        for trian, mu1f in [(self.Gfs, None), (self.Gd1, self.mu1_in), (self.Gd2, self.mu1_out)]:
            if trian.ncells == 0:
                continue
            dG = SurfaceMeasure(trian, deg)
            S = dG.surface_ops(order)
            T = dG.trace_ops(order)
            Dy = T.grad[..., 1]
            rows_w = self.V_phi.cell_dofs[trian.cell] + o_phi
            rows_u = self.V_kap.cell_dofs[self.Gkap.position_of(trian)] + o_kap
            dV = S.dV
            m1 = mu1f(S.x[..., 0]) if mu1f is not None else np.zeros_like(dV)
            m2 = m1 * k
            uk = np.einsum("eq,eqi,eqj->eij", dV, S.val, S.val)
            uw = np.einsum("eq,eqi,eqj->eij", dV, S.val, T.val)
            wu = np.einsum("eq,eqi,eqj->eij", dV, T.val, S.val)
            ww = np.einsum("eq,eqi,eqj->eij", dV, T.val, T.val)
            add(rows_u, rows_u, kk=bh * g * uk)
            add(rows_u, rows_w, kk=np.einsum("eq,eqi,eqj->eij", dV * m1, S.val, Dy), cc=bh * uw)
            add(rows_w, rows_u, kk=bh * g * ah * wu - np.einsum("eq,eqi,eqj->eij", dV * m2, T.val, S.val), cc=-wu)
            add(rows_w, rows_w, kk=ah * np.einsum("eq,eqi,eqj->eij", dV * m1, T.val, Dy), cc=bh * ah * ww)
        for trian, ai in [(self.Gb1, c["a1"]), (self.Gb2, c["a2"])]:
            dG = SurfaceMeasure(trian, deg)
            S = dG.surface_ops(order, need_hess=True)
            T = dG.trace_ops(order)
            rows_w = self.V_phi.cell_dofs[trian.cell] + o_phi
            rows_v = self.V_eta.cell_dofs[self.Geta.position_of(trian)] + o_eta
            mass = np.einsum("eq,eqi,eqj->eij", S.dV, S.val, S.val)
            add(rows_v, rows_v, kk=g * mass + ai * np.einsum("eq,eqi,eqj->eij", S.dV, S.lapl, S.lapl),
                mm=c["d0"] * mass)
            add(rows_v, rows_w, cc=np.einsum("eq,eqi,eqj->eij", S.dV, S.val, T.val))
            add(rows_w, rows_v, cc=-np.einsum("eq,eqi,eqj->eij", S.dV, T.val, S.val))
        addK = lambda r, cc_, blk: add(r, cc_, kk=blk)
        _skeleton_blocks(addK, self, self.Lb1, o_eta, c["a1"], c["gamma"], order, deg, "beam")
        _skeleton_blocks(addK, self, self.Lb2, o_eta, c["a2"], c["gamma"], order, deg, "beam")
        KhabakhpashevaFreqCase._joint(self, addK, o_eta, c["kr"], order, deg)
        return K.tocsr(), C.tocsr(), M.tocsr()

    def rhs(self, t):
        """b(t) (`:253`), assembled literally at time t."""
        c, order, deg = self.c, self.p["order"], self.degree
        tc = self.time_constants()
        o_phi, o_kap = self.X.offsets[:2]
        k, om, ah = c["k"], c["om"], tc["ah"]
        b = VEC(self.X.ndofs, np.float64)
        dGin = SurfaceMeasure(self.Gin, deg)
        S = dGin.surface_ops(1)
        T = dGin.trace_ops(order)
        x = S.x
        vin = -(c["eta0"] * om) * (np.cosh(k * x[..., 1]) / np.sinh(k * c["H"])) * np.cos(k * x[..., 0] - om * t)
        b.add(self.V_phi.cell_dofs[self.Gin.cell] + o_phi, np.einsum("eq,eqi->ei", S.dV * vin, T.val))
        dG = SurfaceMeasure(self.Gd1, deg)
        S = dG.surface_ops(order)
        T = dG.trace_ops(order)
        x1 = S.x[..., 0]
        m1 = self.mu1_in(x1)
        eta_d = m1 * k * c["eta0"] * np.cos(k * x1 - om * t)
        dphi_d = m1 * om * c["eta0"] * np.sin(k * x1 - om * t)
        rows_w = self.V_phi.cell_dofs[self.Gd1.cell] + o_phi
        rows_u = self.V_kap.cell_dofs[self.Gkap.position_of(self.Gd1)] + o_kap
        b.add(rows_w, np.einsum("eq,eqi->ei", S.dV * (-eta_d + ah * dphi_d), T.val))
        b.add(rows_u, np.einsum("eq,eqi->ei", S.dV * dphi_d, S.val))
        return b.v

    def run(self, nsteps=None):
        """Returns (ts, xs, eta_rel per step) as the reference (`:275-305`)."""
        tc = self.time_constants()
        K, C, M = self.assemble()
        nst = nsteps if nsteps is not None else nsteps_for(0.0, tc["tf"], tc["dt"])
        z = np.zeros(self.X.ndofs)
        ts, hist = [], []

        def cb(n, t, u, v, a):
            ts.append(t)
            hist.append(u[self.X.offsets[2]:self.X.offsets[3]].copy())

        u, v, a = newmark(K, C, M, self.rhs, z, z, z, tc["dt"], tc["gamma_t"], tc["beta_t"], nst, callback=cb)
        return np.array(ts), np.array(hist), (u, v, a)
