"""Lanczos driver over the GPU matrix-free H*v: stands in for KrylovKit's `eigsolve(ham, ::PDVec, ...)`
(ext/KrylovKitExt.jl:23-46).  The recurrence needs only the vector ops the C ABI provides
(mul!, add!/axpby!, dot, norm, scale!); the small tridiagonal eigenproblem is solved on the host."""
from __future__ import annotations

import numpy as np

from .dictvectors import GPUDVec, WorkingMemory, mul
from .stochasticstyles import IsDeterministic


def eigsolve_lanczos(ham, start: GPUDVec, howmany=1, krylovdim=60, tol=1e-10, maxiter=20, full_reorth=False):
    """Lowest `howmany` eigenvalues of a Hermitian device Hamiltonian from `start`.
    Thick restarts are replaced by plain restarts from the current Ritz vector.
    Returns (values, vectors, info)."""
    det = IsDeterministic()
    v = GPUDVec(style=det, address_type=start.address_type, ctx=start.ctx).copy_from(start)
    wm = WorkingMemory(v)
    info = {"matvecs": 0, "converged": False, "residual": np.inf}
    theta = None
    for restart in range(maxiter):
        nrm = v.norm(2)
        v.scale_(1.0 / nrm)
        basis = [v]
        alphas, betas = [], []
        w = v.similar()
        for j in range(krylovdim):
            mul(w, ham, basis[j], wm)
            info["matvecs"] += 1
            a = basis[j].dot(w)
            alphas.append(a)
            w.add_(basis[j], -a)
            if j > 0:
                w.add_(basis[j - 1], -betas[j - 1])
            if full_reorth:
                for q in basis:
                    w.add_(q, -q.dot(w))
            b = w.norm(2)
            T = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1)
            evals, evecs = np.linalg.eigh(T)
            resid = abs(b * evecs[-1, 0])
            theta = evals
            if resid < tol or b < 1e-14 or j == krylovdim - 1:
                info["residual"] = resid
                break
            betas.append(b)
            nxt = w.copy().scale_(1.0 / b)
            basis.append(nxt)
            w = v.similar()
        # Ritz vector of the lowest state
        y = evecs[:, 0]
        ritz = basis[0].copy().scale_(y[0])
        for q, c in zip(basis[1:], y[1:]):
            ritz.add_(q, c)
        v = ritz
        if info["residual"] < tol:
            info["converged"] = True
            break
    vals = theta[:howmany]
    return vals, [v], info
