"""Conservation laws, numerical-flux tags and initial data (host descriptors).

Mirrors the type surface of src/ConservationLaws (ConservationLaws.jl:40-72,
linear_advection_diffusion.jl:1-47, euler_navierstokes.jl:23-38) and the grid
functions used by the BASELINE configs (GridFunctions.jl:118-134,
euler_navierstokes.jl:234-320).  The pointwise physics itself is evaluated on the
device (csrc/physics.cuh); nothing here is on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

import numpy as np

# ids shared with include/sse_b200.h
PDE_ADVECTION, PDE_ADVECTION_DIFFUSION, PDE_EULER, PDE_BURGERS, PDE_VISCOUS_BURGERS = 0, 1, 2, 3, 4
FLUX_LAX_FRIEDRICHS, FLUX_CENTRAL, FLUX_ENTROPY_CONSERVATIVE = 0, 1, 2
TWO_POINT_CONSERVATIVE, TWO_POINT_ENTROPY_CONSERVATIVE = 0, 1


@dataclass(frozen=True)
class LinearAdvectionEquation:
    a: Tuple[float, ...]
    pde_id: int = PDE_ADVECTION

    @property
    def d(self):
        return len(self.a)

    N_c = 1
    second_order = False


@dataclass(frozen=True)
class InviscidBurgersEquation:
    """burgers.jl:1-21, 46: flux a u^2 / 2; `InviscidBurgersEquation()` is the 1-D law with a = (1,)."""
    a: Tuple[float, ...] = (1.0,)
    pde_id: int = PDE_BURGERS

    @property
    def d(self):
        return len(self.a)

    N_c = 1
    second_order = False


@dataclass(frozen=True)
class ViscousBurgersEquation:
    """burgers.jl:23-49: flux a u^2 / 2 - b q (BR1); `ViscousBurgersEquation(b)` is the 1-D law with a = (1,)."""
    a: Tuple[float, ...] = (1.0,)
    b: float = 0.0
    pde_id: int = PDE_VISCOUS_BURGERS

    @property
    def d(self):
        return len(self.a)

    N_c = 1
    second_order = True


@dataclass(frozen=True)
class LinearAdvectionDiffusionEquation:
    a: Tuple[float, ...]
    b: float
    pde_id: int = PDE_ADVECTION_DIFFUSION

    @property
    def d(self):
        return len(self.a)

    N_c = 1
    second_order = True


@dataclass(frozen=True)
class EulerEquations:
    d: int
    gamma: float = 1.4
    pde_id: int = PDE_EULER
    second_order = False

    @property
    def N_c(self):
        return self.d + 2


# numerical fluxes (ConservationLaws.jl:52-72)
@dataclass(frozen=True)
class LaxFriedrichsNumericalFlux:
    lam: float = 1.0
    flux_id: int = FLUX_LAX_FRIEDRICHS

    @property
    def half_lambda(self):
        return 0.5 * self.lam


@dataclass(frozen=True)
class CentralNumericalFlux:
    flux_id: int = FLUX_CENTRAL
    half_lambda: float = 0.0


@dataclass(frozen=True)
class EntropyConservativeNumericalFlux:
    flux_id: int = FLUX_ENTROPY_CONSERVATIVE
    half_lambda: float = 0.0


@dataclass(frozen=True)
class BR1:
    pass


@dataclass(frozen=True)
class ConservativeFlux:
    two_point_id: int = TWO_POINT_CONSERVATIVE


@dataclass(frozen=True)
class EntropyConservativeFlux:
    two_point_id: int = TWO_POINT_ENTROPY_CONSERVATIVE


# --------------------------------------------------------------------------
# initial data: callables  f(x: list of arrays) -> array (..., N_c)
# --------------------------------------------------------------------------
def initial_data_sine(A, k):
    k = np.atleast_1d(k)
    return lambda x: (A * np.prod([np.sin(k[m] * x[m]) for m in range(len(x))], axis=0))[..., None]


def initial_data_gassner(k, eps):
    """InitialDataGassner (GridFunctions.jl:60-68, 144-146)."""
    return lambda x: (np.sin(k * x[0]) + eps)[..., None]


def initial_data_cosine(A, k):
    k = np.atleast_1d(k)
    return lambda x: (A * np.prod([np.cos(k[m] * x[m]) for m in range(len(x))], axis=0))[..., None]


def isentropic_vortex(gamma=1.4, Ma=0.4, theta=np.pi / 4, R=1.0, beta=1.0, sigma=1.0,
                      x_0=(0.0, 0.0)):
    """euler_navierstokes.jl:255-266"""
    def f(x):
        xr = ((x[0] - x_0[0]) / R, (x[1] - x_0[1]) / R)
        Om = beta * np.exp(-0.5 / sigma ** 2 * (xr[0] ** 2 + xr[1] ** 2))
        dv = (-xr[1] * Om, xr[0] * Om)
        dT = -0.5 * (gamma - 1) * Om ** 2
        rho = (1 + dT) ** (1 / (gamma - 1))
        v = (Ma * np.cos(theta) + dv[0], Ma * np.sin(theta) + dv[1])
        p = rho ** gamma / gamma
        E = p / (gamma - 1) + 0.5 * rho * (v[0] ** 2 + v[1] ** 2)
        return np.stack([rho, rho * v[0], rho * v[1], E], axis=-1)
    return f


def taylor_green_vortex(gamma=1.4, Ma=0.1):
    """euler_navierstokes.jl:311-320"""
    def f(x):
        p = (1 / (Ma ** 2 * gamma)) + 0.0625 * (2 * np.cos(2 * x[0]) + 2 * np.cos(2 * x[1])
                                              + np.cos(2 * x[0]) * np.cos(2 * x[2])
                                              + np.cos(2 * x[1]) * np.cos(2 * x[2]))
        u = np.sin(x[0]) * np.cos(x[1]) * np.cos(x[2])
        v = -np.cos(x[0]) * np.sin(x[1]) * np.cos(x[2])
        one = np.ones_like(u)
        return np.stack([one, u, v, 0 * one, p / (gamma - 1) + 0.5 * (u ** 2 + v ** 2)], axis=-1)
    return f


def euler_periodic_test(d, gamma=1.4, strength=0.2, L=2.0):
    """euler_navierstokes.jl:289-294"""
    def f(x):
        rho = 1.0 + strength * np.sin(2 * np.pi * sum(x[m] for m in range(d)) / L)
        return np.stack([rho] + [rho] * d + [1.0 / (gamma - 1.0) + 0.5 * rho * d], axis=-1)
    return f


def project_function(f, ra, J_q, xyzq):
    """initialize / project_function (Solvers.jl:388-427): nodal schemes evaluate the
    data, modal schemes take the per-element L2 projection with the true mass matrix.
    Returns u0 with shape (N_e, N_c, N_p)."""
    u_q = f(xyzq)                                          # (N_e, N_q, N_c)
    if ra.V_is_identity:
        return np.ascontiguousarray(np.transpose(u_q, (0, 2, 1)))
    V, W = ra.V, ra.W
    out = np.empty((u_q.shape[0], u_q.shape[2], V.shape[1]))
    for s in range(0, u_q.shape[0], 16384):
        WJ = W[None, :] * J_q[s:s + 16384]                 # (n, N_q)
        VtWJ = V.T[None, :, :] * WJ[:, None, :]            # (n, N_p, N_q)
        u0 = np.linalg.solve(VtWJ @ V, VtWJ @ u_q[s:s + 16384])   # (n, N_p, N_c)
        out[s:s + 16384] = np.transpose(u0, (0, 2, 1))
    return out


def project_function_reference(f, ra, xyzq):
    """Cheap synthetic state for large benchmarks: u = V' W f(x_q) (the L2 projection on the
    reference element, exact for affine elements since V' W V = I)."""
    u_q = f(xyzq)
    if ra.V_is_identity:
        return np.ascontiguousarray(np.transpose(u_q, (0, 2, 1)))
    return np.ascontiguousarray(np.einsum("qa,kqc->kca", ra.V * ra.W[:, None], u_q, optimize=True))
