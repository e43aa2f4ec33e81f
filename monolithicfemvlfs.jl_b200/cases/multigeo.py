"""B200 mirror of `run_MultiGeo_freq_domain` (src/Multi_geo_freq_domain.jl:19-151): same
parameters, same weak form on the tetrahedral gmsh model, assembled and solved by
libvlfs_b200 on the GPU."""
from dataclasses import dataclass
import numpy as np
from .. import _lib as L
from ..geometry import DiscreteModelFromFile, Interior, Boundary, Skeleton, refine_uniform
from ..fe import ReferenceFE, TestFESpace, TrialFESpace, MultiFieldFESpace, Measure, lagrangian
from ..assembler import B200System


@dataclass
class MultiGeo_freq_domain_params:      # src/Multi_geo_freq_domain.jl:11-17
    name: str = "MultigeoFreq"
    mesh_file: str = "models/multi_geo_coarse.json"
    order: int = 4
    dfactor: float = 3
    vtk_output: bool = False
    refine: int = 0                     # (not in the reference) levels of uniform refinement: synthetic M2 meshes


class MultiGeoProblem:
    """Everything up to and including `op = AffineFEOperator(a,b,X,Y)` (:26-131)."""

    def __init__(self, params, device=0, system_cls=B200System, model=None):
        p = self.params = params
        order = p.order
        g, rho, E, rhob, nu, hb = 9.81, 1025, 11.9e9, 256.25, 0.13, 0.1
        I = hb ** 3 / 12
        D_ = E * I / (1.0 - nu ** 2)
        d0 = rhob * hb / rho
        mu_ = E / (2 * (1 + nu)); lam = nu * E / (1 - nu ** 2)
        Ct = np.zeros((3, 3, 3, 3))
        dl = np.eye(2)
        for i in range(2):
            for j in range(2):
                for k in range(2):
                    for l in range(2):
                        Ct[i, j, k, l] = I / rho * (mu_ * (dl[i, k] * dl[j, l] + dl[i, l] * dl[j, k])
                                                    + lam * dl[i, j] * dl[k, l])
        H = 3
        k = 2 * np.pi / 1.0
        om = np.sqrt(g * k * np.tanh(k * H))
        bh = 0.5
        ah = -1j * om / g * (1 - bh) / bh
        self.c = c = dict(g=g, rho=rho, D=D_, d0=d0, H=H, k=k, om=om, eta0=0.01, bh=bh, ah=ah,
                          mu0=2.5, Ld=3, Ld0=3, xd=10, xd0=-1)
        model = model if model is not None else DiscreteModelFromFile(p.mesh_file)
        for _ in range(int(p.refine)):
            model = refine_uniform(model)
        self.model = model
        Om = Interior(model)
        Gin = Boundary(model, tags="inlet")
        Gb = Boundary(model, tags="plate")
        Gf = Boundary(model, tags="freeSurface")
        Lb_ = Skeleton(Gb)
        self.Om, self.Gin, self.Gb, self.Gf, self.Lb = Om, Gin, Gb, Gf, Lb_
        degree = 2 * order
        dOm, dGb, dGf, dGin, dLb = (Measure(t, degree) for t in (Om, Gb, Gf, Gin, Lb_))
        cplx = np.complex128
        V_Om = TestFESpace(Om, ReferenceFE(lagrangian, float, 2), vector_type=cplx)
        V_Gk = TestFESpace(Gf, ReferenceFE(lagrangian, float, order), vector_type=cplx)
        V_Ge = TestFESpace(Gb, ReferenceFE(lagrangian, float, order), vector_type=cplx)
        self.X = X = MultiFieldFESpace([TrialFESpace(V_Om), TrialFESpace(V_Gk), TrialFESpace(V_Ge)])
        self.V_Om, self.V_Gk, self.V_Ge = V_Om, V_Gk, V_Ge
        sys_ = self.sys = system_cls(X.ndofs, True, nmat=1, nvec=1, device=device)
        ez = (0.0, 0.0, 1.0)
        d = sys_.add_domain(3, dOm.nq, dOm.wq, [dOm.own_slot(V_Om)])
        d.add_term(0, 1.0, L.OP_GRAD, 0, L.OP_GRAD, 0)                        # ∇(ϕ)⋅∇(w)  (:126)
        # Γf (:127)
        x1 = dGf.points()[..., 0]
        m1 = (c["mu0"] * (1.0 - np.cos(np.pi / 2 * (x1 - c["xd"]) / c["Ld"])) * (x1 > c["xd"])
              + c["mu0"] * (1 - np.cos(np.pi / 2 * (c["Ld0"] - x1) / c["Ld0"])) * (x1 < c["xd0"]))
        eta_d = m1 * k * c["eta0"] * np.exp(1j * k * x1) * (x1 < c["xd0"])
        dphi_d = m1 * (-1j * om * c["eta0"] * np.exp(1j * k * x1)) * (x1 < c["xd0"])
        d = sys_.add_domain(3, dGf.nq, dGf.wq, [dGf.trace_slot(V_Om), dGf.surface_slot(V_Gk)], measure_slot=1,
                            weights=[m1, eta_d.real, eta_d.imag, dphi_d.real, dphi_d.imag])
        PHI, KAP = 0, 1
        d.add_term(0, bh * g, L.OP_VAL, KAP, L.OP_VAL, KAP)
        d.add_term(0, -1j * om * bh, L.OP_VAL, KAP, L.OP_VAL, PHI)
        d.add_term(0, 1.0, L.OP_VAL, KAP, L.OP_DDIR, PHI, weight=0, trial_dir=ez)
        d.add_term(0, bh * g * ah + 1j * om, L.OP_VAL, PHI, L.OP_VAL, KAP)
        d.add_term(0, -k, L.OP_VAL, PHI, L.OP_VAL, KAP, weight=0)
        d.add_term(0, -1j * om * bh * ah, L.OP_VAL, PHI, L.OP_VAL, PHI)
        d.add_term(0, ah, L.OP_VAL, PHI, L.OP_DDIR, PHI, weight=0, trial_dir=ez)
        d.add_rhs(0, -1.0, L.OP_VAL, PHI, weight_re=1, weight_im=2)
        d.add_rhs(0, ah, L.OP_VAL, PHI, weight_re=3, weight_im=4)
        d.add_rhs(0, 1.0, L.OP_VAL, KAP, weight_re=3, weight_im=4)
        # Γb (:128)
        d = sys_.add_domain(3, dGb.nq, dGb.wq, [dGb.trace_slot(V_Om), dGb.surface_slot(V_Ge, need_hess=True)],
                            measure_slot=1, Ctensor=Ct)
        ETA = 1
        d.add_term(0, g - om ** 2 * d0, L.OP_VAL, ETA, L.OP_VAL, ETA)
        d.add_term(0, 1.0, L.OP_HESS, ETA, L.OP_CHESS, ETA)
        d.add_term(0, -1j * om, L.OP_VAL, ETA, L.OP_VAL, PHI)
        d.add_term(0, 1j * om, L.OP_VAL, PHI, L.OP_VAL, ETA)
        # Λb (:129): γ = order(order+1)/h, h = get_cell_measure(Λb) -> coefficient field 1/h
        hinv = np.repeat((1.0 / Lb_.cell_measure())[:, None], dLb.nq, axis=1)
        d = sys_.add_domain(3, dLb.nq, dLb.wq, [dLb.skeleton_slot(V_Ge, "plus"), dLb.skeleton_slot(V_Ge, "minus")],
                            measure_slot=0, measure_kind=L.MEASURE_FACET, Ctensor=Ct, weights=[hinv])
        jump, mean = {0: 1.0, 1: -1.0}, {0: 0.5, 1: 0.5}
        d.add_term(0, -1.0, L.OP_GRAD, jump, L.OP_CHESS_NPLUS, mean)
        d.add_term(0, -1.0, L.OP_CHESS_NPLUS, mean, L.OP_GRAD, jump)
        d.add_term(0, D_ / rho * order * (order + 1), L.OP_GRAD, jump, L.OP_GRAD, jump, weight=0)
        # Γin rhs (:130)
        V_aux = TestFESpace(Gin, ReferenceFE(lagrangian, float, 1))
        V_aux.offset = 0
        x = dGin.points()
        vin = (c["eta0"] * om) * (np.cosh(k * x[..., 2]) / np.sinh(k * H)) * np.exp(1j * k * x[..., 0])
        d = sys_.add_domain(3, dGin.nq, dGin.wq, [dGin.trace_slot(V_Om), dGin.surface_slot(V_aux)],
                            measure_slot=1, weights=[vin.real, vin.imag])
        d.add_rhs(0, 1.0, L.OP_VAL, PHI, weight_re=0, weight_im=1)
        if hasattr(sys_, "set_dof_coords"):
            sys_.set_dof_coords(X.dof_coords())
        self.nnz = sys_.build_pattern()

    def assemble(self):
        self.sys.assemble()


def run_MultiGeo_freq_domain(params, device=0, solver=None):
    prob = MultiGeoProblem(params, device=device)
    prob.assemble()
    opts = dict(method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    opts.update(solver or {})
    prob.sys.solver_setup(0, **opts)
    x, prob.stats = prob.sys.solve(vec=0)
    if params.vtk_output:                                   # src/Multi_geo_freq_domain.jl:145-147
        from ..vtk import writevtk
        phi, kap, et = prob.X.split(x)
        name = getattr(params, "name", "MultiGeo")
        writevtk(prob.Om, name + "_O", [("phi_re", prob.V_Om, phi.real), ("phi_im", prob.V_Om, phi.imag)], nsubcells=10)
        writevtk(prob.Gf, name + "_Gk", [("eta_re", prob.V_Gk, kap.real), ("eta_im", prob.V_Gk, kap.imag)], nsubcells=10)
        writevtk(prob.Gb, name + "_Ge", [("eta_re", prob.V_Ge, et.real), ("eta_im", prob.V_Ge, et.imag)], nsubcells=10)
    return prob.X.split(x)
