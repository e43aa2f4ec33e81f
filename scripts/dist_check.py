"""Real multi-GPU correctness check (run under torchrun, one rank per GPU):
distributed assembly + NCCL-halo SpMV + distributed GMRES against the single-GPU path of rank 0.
  gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import vlfs_b200
from vlfs_b200 import _lib as L, dist as D
from vlfs_b200.cases.yago import YagoProblem, Yago_freq_domain_params

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
kw = dict(nx=8, ny=2, nz=2, order=4, lfactor=0.4, dfactor=4)
mk = lambda *a, **k: D.DistributedSystem(*a, rank=rank, nranks=world, **{**k, "device": local})
prob = YagoProblem(Yago_freq_domain_params(**kw), device=local, system_cls=mk)
ops = D.GpuOps(prob.sys); ops.bind_stream()
prob.assemble()
member = D.Member(prob.sys, ops)
fleet = D.Fleet([member], world, dist=dist)
schur = D.SchurSolver(fleet)
schur.setup()
N = prob.sys.ndofs_global
# SpMV
x = np.sin(0.1 * np.arange(N)) + 1j * np.cos(0.1 * np.arange(N))
xl = prob.sys.to_local(x); xl[prob.sys.n_owned:] = 0
xs, ys = [ops.from_host(xl)], [ops.zeros()]
fleet.matvec(xs, ys)
y_owned = ops.to_host(ys[0])[:prob.sys.n_owned]
# solve
bs, x0 = [ops.rhs_vector()], [ops.zeros()]
st = fleet.gmres(bs, x0, restart=8, maxit=8, rtol=1e-11, apply=schur.apply)
x_owned = ops.to_host(x0[0])[:prob.sys.n_owned]
gathered = [None] * world
dist.all_gather_object(gathered, (prob.sys.owned, y_owned, x_owned))
if rank == 0:
    ref = YagoProblem(Yago_freq_domain_params(**kw), device=local)
    ref.assemble()
    yr = ref.sys.spmv(x)
    ref.sys.solver_setup(0, method=L.SOLVER_GMRES, precond=L.PRECOND_SPARSE_LU, restart=30, maxit=60, rtol=1e-12)
    xr, _ = ref.sys.solve(vec=0)
    y = np.zeros(N, dtype=complex); xd = np.zeros(N, dtype=complex)
    for owned, yo, xo in gathered:
        y[owned] = yo; xd[owned] = xo
    out = dict(world=world, N=N, spmv_err=float(np.abs(y - yr).max() / np.abs(yr).max()),
               solve_err=float(np.linalg.norm(xd - xr) / np.linalg.norm(xr)), gmres=st,
               halo_bytes=fleet.halo_bytes, interface_unknowns=int(schur.ng))
    print(json.dumps(out))
    assert out["spmv_err"] <= 1e-12 and out["solve_err"] <= 1e-8, out
dist.destroy_process_group()
