"""Lanczos driver over the GPU matrix-free H*v: stands in for KrylovKit's `eigsolve(ham, ::PDVec, ...)`
(ext/KrylovKitExt.jl:23-46).  The recurrence needs only the vector ops the C ABI provides
(mul!, add!/axpby!, dot, norm, scale!); the small tridiagonal eigenproblem is solved on the host."""
from __future__ import annotations

import numpy as np

from .dictvectors import GPUDVec, WorkingMemory, mul
from .stochasticstyles import IsDeterministic


def eigsolve_lanczos(ham, start: GPUDVec, howmany=1, krylovdim=60, tol=1e-10, maxiter=20, full_reorth=False):
    """Lowest `howmany` eigenvalues of a Hermitian device Hamiltonian from `start`.

    howmany == 1: thick restarts are replaced by plain restarts from the current Ritz vector (up to `maxiter` cycles).
    howmany > 1: a restart from one vector would collapse onto the lowest state, so a single Krylov space of `krylovdim`
    vectors is built (full reorthogonalisation is switched on) and `info["converged"]` says whether ALL requested Ritz pairs
    reached `tol`; raise `krylovdim` otherwise.  Returns (values, vectors, info): `howmany` values and their Ritz vectors."""
    if howmany < 1:
        raise ValueError("howmany must be >= 1")
    if howmany > krylovdim:
        raise ValueError("howmany cannot exceed krylovdim")
    det = IsDeterministic()
    if hasattr(start, "mul_from"):  # dense sector vector (sectors.py): no working memory, same vector operations
        v, wm = start.copy(), None
    else:
        v = GPUDVec(style=det, address_type=start.address_type, ctx=start.ctx).copy_from(start)
        wm = WorkingMemory(v)
    info = {"matvecs": 0, "converged": False, "residual": np.inf}
    if howmany > 1:
        full_reorth, maxiter = True, 1
    theta, ritz_vectors = None, [v]
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
            # residual of Ritz pair i is |beta_j * (last component of its eigenvector)|; all requested pairs must converge
            nreq = min(howmany, len(evals))
            resid = float(np.max(np.abs(b * evecs[-1, :nreq]))) if len(evals) >= howmany else np.inf
            theta = evals
            if resid < tol or b < 1e-14 or j == krylovdim - 1:
                info["residual"] = resid if len(evals) >= howmany else np.inf
                if b < 1e-14 and len(evals) >= howmany:  # invariant subspace: the Ritz pairs are exact
                    info["residual"] = 0.0
                break
            betas.append(b)
            nxt = w.copy().scale_(1.0 / b)
            basis.append(nxt)
            w = v.similar()
        ritz_vectors = []
        for i in range(min(howmany, evecs.shape[1])):
            y = evecs[:, i]
            ritz = basis[0].copy().scale_(y[0]) if y[0] != 0.0 else basis[0].similar()
            for q, c in zip(basis[1:], y[1:]):
                if c != 0.0:
                    ritz.add_(q, c)
            ritz_vectors.append(ritz)
        v = ritz_vectors[0]
        if info["residual"] < tol:
            info["converged"] = True
            break
    if len(theta) < howmany:
        raise ValueError(f"the Krylov space closed after {len(theta)} vectors: fewer than howmany={howmany} states overlap the start vector")
    return theta[:howmany], ritz_vectors, info
