"""Newmark-beta stepping on a constant operator (oracle; test infrastructure only).

Restates what Gridap 0.17 does for
`solve(Newmark(LUSolver(),dt,γ,β), TransientConstant[Matrix]FEOperator(m,c,a,b,X,Y), (x₀,v₀,a₀), t₀, tf)`
(reference call sites `src/Periodic_Beam.jl:104-116`, `src/Khabakhpasheva_time_domain.jl:257-266`;
SURVEY.md §8 row a7): the matrix  A = K + γ/(βΔt) C + 1/(βΔt²) M  is assembled and factorised
once (sparse LU), every step solves

    A u₁ = b(t₁) - C v* - M a*,
    v* = -γ/(βΔt) u₀ + (1-γ/β) v₀ + Δt (1-γ/(2β)) a₀,
    a* = -1/(βΔt²) u₀ - 1/(βΔt) v₀ - (1-2β)/(2β) a₀,
    v₁ = v* + γ/(βΔt) u₁,   a₁ = a* + 1/(βΔt²) u₁

and the iterator yields (u₁, t₁) for n = 1 … ⌈(tf-t₀)/Δt⌉.
"""
import numpy as np
import scipy.sparse.linalg as spla


def nsteps_for(t0, tf, dt):
    """Number of states Gridap's iterator produces: it advances while t < tf - ε."""
    n = 0
    t = t0
    while t < tf - 100 * np.finfo(float).eps * max(abs(tf), 1.0):
        t += dt
        n += 1
    return n


def newmark(K, C, M, rhs, u0, v0, a0, dt, gamma, beta, nsteps, t0=0.0, callback=None):
    """rhs(t) -> vector (or None for zero).  Returns the final (u,v,a); `callback(n,t,u,v,a)`
    is called after every step."""
    cu = gamma / (beta * dt)
    cuu = 1.0 / (beta * dt * dt)
    A = (K + cu * C + cuu * M).tocsc()
    lu = spla.splu(A)
    u, v, a = (np.array(z, dtype=float) for z in (u0, v0, a0))
    t = t0
    for n in range(1, nsteps + 1):
        t = t0 + n * dt
        vs = -cu * u + (1.0 - gamma / beta) * v + dt * (1.0 - gamma / (2 * beta)) * a
        as_ = -cuu * u - v / (beta * dt) - (1.0 - 2 * beta) / (2 * beta) * a
        b = rhs(t) if rhs is not None else None
        r = -(C @ vs) - (M @ as_)
        if b is not None:
            r = r + b
        u1 = lu.solve(r)
        v = vs + cu * u1
        a = as_ + cuu * u1
        u = u1
        if callback is not None:
            callback(n, t, u, v, a)
    return u, v, a
