"""Conservation / energy / entropy residual functionals (host NumPy version).

Restates Analysis/conservation.jl:145-189 for post-run checks; the device version is
`sse_functionals`.  State arrays are (N_e, N_c, N_p) C-ordered."""
from __future__ import annotations

import numpy as np

from . import _abi


def _mass_matrix(ra, gf, mass_solver):
    """mass_matrix(mass_solver, k) (mass_matrix.jl:140-153) for all k: (N_e, N_p, N_p)."""
    if mass_solver == _abi.SSE_MASS_DIAGONAL:
        return np.einsum("ki,ij->kij", ra.W[None, :] * gf.J_q, np.eye(ra.N_p))
    V = ra.V
    if mass_solver == _abi.SSE_MASS_CHOLESKY:                  # V' WJ V   mass_matrix.jl:140-143
        return np.einsum("qa,kq,qb->kab", V, ra.W[None, :] * gf.J_q, V)
    Minv = np.einsum("qa,kq,qb->kab", V, ra.W[None, :] / gf.J_q, V)
    return np.linalg.inv(Minv)


def conservation_residual(image, dudt):
    """sum_k 1' WJ_k V dudt[:, e, k]  (conservation.jl:145-152)."""
    ra, gf = image.sd.reference_approximation, image.sd.geometric_factors
    WJ = ra.W[None, :] * gf.J_q
    return np.einsum("kq,qa,kea->e", WJ, ra.V, dudt)


def energy_residual(image, u, dudt):
    """sum_k u_k' M_k dudt_k per variable (conservation.jl:154-167)."""
    ra, gf = image.sd.reference_approximation, image.sd.geometric_factors
    M = _mass_matrix(ra, gf, int(image.cfg.mass_solver))
    return np.einsum("kea,kab,keb->e", u, M, dudt)


def conservative_to_entropy(gamma, u_q):
    """Euler entropy variables (euler_navierstokes.jl:100-113); u_q: (..., N_c)."""
    d = u_q.shape[-1] - 2
    rho, E = u_q[..., 0], u_q[..., d + 1]
    k = (0.5 / rho) * np.sum(u_q[..., 1:d + 1] ** 2, axis=-1)
    p = (gamma - 1) * (E - k)
    w = np.empty_like(u_q)
    w[..., 0] = (gamma - np.log(p / rho ** gamma)) / (gamma - 1) - k / p
    w[..., 1:d + 1] = u_q[..., 1:d + 1] / p[..., None]
    w[..., d + 1] = -rho / p
    return w


def entropy_residual(image, u, dudt):
    """sum_k (P_k w(V u_k))' M_k dudt_k (conservation.jl:169-189), P_k = M_k^-1 V' WJ_k."""
    ra, gf = image.sd.reference_approximation, image.sd.geometric_factors
    gamma = float(image.cfg.gamma)
    WJ = ra.W[None, :] * gf.J_q
    u_q = np.einsum("qa,kea->kqe", ra.V, u)
    w_q = conservative_to_entropy(gamma, u_q)
    # (P w)' M dudt = (M^-1 V' WJ w)' M dudt = (V' WJ w)' dudt  for symmetric M
    return float(np.einsum("qa,kq,kqe,kea->", ra.V, WJ, w_q, dudt))
