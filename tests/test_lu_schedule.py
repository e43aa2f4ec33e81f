"""The blocked right-looking schedule of the multifrontal LU (csrc/direct.cu::factor_levels and the
three regions of `schur_update`), restated in NumPy and checked against a plain partial LU without
pivoting: panels of NB pivots inside outer blocks of `outer` panels; after every panel only the
block's column strip (mode 1) and row strip (mode 2) are updated, the trailing matrix (mode 0)
once per outer block with the whole block as depth - CPU only, guards the algorithm not the CUDA."""
import numpy as np
import pytest


def plain_partial_lu(F, npiv):
    F = F.copy()
    for j in range(npiv):
        F[j + 1:, j] /= F[j, j]
        F[j + 1:, j + 1:] -= np.outer(F[j + 1:, j], F[j, j + 1:])
    return F


def blocked_partial_lu(F, npiv, NB, outer):
    F = F.copy()
    nf = F.shape[0]
    npan = (npiv + NB - 1) // NB
    for K0 in range(0, npan, outer):
        K1 = min(K0 + outer, npan)
        Kend = min(K1 * NB, npiv)
        for k in range(K0, K1):
            k0 = k * NB
            if k0 >= npiv:
                continue
            k1 = min(k0 + NB, npiv)
            D = F[k0:k1, k0:k1]                                   # diag_lu
            for j in range(k1 - k0):
                D[j + 1:, j] /= D[j, j]
                D[j + 1:, j + 1:] -= np.outer(D[j + 1:, j], D[j, j + 1:])
            L11 = np.tril(D, -1) + np.eye(k1 - k0)
            U11 = np.triu(D)
            F[k1:, k0:k1] = F[k1:, k0:k1] @ np.linalg.inv(U11)    # panel_trsm, L-panel (all rows below)
            F[k0:k1, k1:] = np.linalg.inv(L11) @ F[k0:k1, k1:]    # panel_trsm, U-panel (all columns right)
            d1 = k1
            if k < K1 - 1:
                # mode 1: rows >= d1, columns [d1, Kend)
                F[d1:, d1:Kend] -= F[d1:, k0:k1] @ F[k0:k1, d1:Kend]
                # mode 2: rows [d1, Kend), columns >= Kend
                F[d1:Kend, Kend:] -= F[d1:Kend, k0:k1] @ F[k0:k1, Kend:]
        d0 = K0 * NB
        if d0 < npiv:                                             # mode 0: trailing matrix, depth = the whole block
            F[Kend:, Kend:] -= F[Kend:, d0:Kend] @ F[d0:Kend, Kend:]
    return F


@pytest.mark.parametrize("nf,npiv,NB,outer", [(60, 60, 8, 4), (75, 50, 8, 3), (41, 17, 4, 4), (33, 33, 32, 4),
                                             (50, 21, 8, 1), (64, 40, 8, 8), (30, 5, 8, 4)])
def test_blocked_schedule_equals_plain_partial_lu(nf, npiv, NB, outer):
    rng = np.random.default_rng(nf * 1000 + npiv)
    A = rng.standard_normal((nf, nf)) + 1j * rng.standard_normal((nf, nf)) + nf * np.eye(nf)
    ref = plain_partial_lu(A, npiv)
    out = blocked_partial_lu(A, npiv, NB, outer)
    assert np.abs(out - ref).max() <= 1e-11 * np.abs(ref).max()
    # the trailing block is the Schur complement of the eliminated pivots
    if npiv == nf:
        return
    S = A[npiv:, npiv:] - A[npiv:, :npiv] @ np.linalg.solve(A[:npiv, :npiv], A[:npiv, npiv:])
    assert np.abs(out[npiv:, npiv:] - S).max() <= 1e-10 * max(np.abs(S).max(), 1.0)
