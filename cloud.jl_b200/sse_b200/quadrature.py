"""1-D Gauss-type quadrature rules and orthonormal Jacobi polynomials.

Host-side setup only (runs once, never on the hot path).  Restates what the
reference obtains from Jacobi.jl / StartUpDG:

* ``quadrature(Line(), GaussQuadrature(q, a, b))``  -> ``zgj/wgj``
  (reference: src/SpatialDiscretizations/quadrature_rules.jl:82-86)
* ``quadrature(Line(), GaussLobattoQuadrature(q, a, b))`` -> ``zglj/wglj``
  (quadrature_rules.jl:76-80)
* ``jacobiP(x, alpha, beta, n)``: the *orthonormal* Jacobi polynomial of the
  Hesthaven--Warburton convention used by StartUpDG/NodesAndModes
  (call sites: src/SpatialDiscretizations/tensor_simplex.jl:96-99,123-130).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import gamma as _gamma

import numpy as np
from scipy.special import roots_jacobi


@dataclass(frozen=True)
class GaussQuadrature:
    """Gauss--Jacobi rule with q+1 nodes for weight (1-x)^a (1+x)^b."""
    q: int
    a: int = 0
    b: int = 0


@dataclass(frozen=True)
class GaussLobattoQuadrature:
    q: int
    a: int = 0
    b: int = 0


def LGQuadrature(q: int) -> GaussQuadrature:
    return GaussQuadrature(q, 0, 0)


def LGLQuadrature(q: int) -> GaussLobattoQuadrature:
    return GaussLobattoQuadrature(q, 0, 0)


def quadrature_line(rule):
    """Nodes and weights on [-1, 1] (ascending)."""
    if isinstance(rule, GaussQuadrature):
        x, w = roots_jacobi(rule.q + 1, float(rule.a), float(rule.b))
        return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)
    if isinstance(rule, GaussLobattoQuadrature):
        n = rule.q + 1
        a, b = float(rule.a), float(rule.b)
        if n == 2:
            xi = np.zeros(0)
        else:
            # interior Gauss--Lobatto--Jacobi nodes are the Gauss--Jacobi nodes
            # of the (a+1, b+1) weight
            xi, _ = roots_jacobi(n - 2, a + 1.0, b + 1.0)
        x = np.concatenate([[-1.0], xi, [1.0]])
        # weights by exactness on the degree <= n-1 orthonormal Jacobi basis
        V = np.stack([jacobiP(x, a, b, k) for k in range(n)], axis=0)
        rhs = np.zeros(n)
        rhs[0] = np.sqrt(2.0 ** (a + b + 1) * _gamma(a + 1) * _gamma(b + 1) / _gamma(a + b + 2))
        w = np.linalg.solve(V, rhs)
        return x, w
    raise TypeError(f"unsupported quadrature rule {rule!r}")


def jacobiP(x, alpha: float, beta: float, N: int) -> np.ndarray:
    """Orthonormal Jacobi polynomial P_N^{(alpha,beta)}(x) (HW08 convention)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    PL = np.zeros((N + 1, x.size))
    gamma0 = (2.0 ** (alpha + beta + 1) / (alpha + beta + 1) * _gamma(alpha + 1)
              * _gamma(beta + 1) / _gamma(alpha + beta + 1))
    PL[0] = 1.0 / np.sqrt(gamma0)
    if N == 0:
        return PL[0]
    gamma1 = (alpha + 1) * (beta + 1) / (alpha + beta + 3) * gamma0
    PL[1] = ((alpha + beta + 2) * x / 2 + (alpha - beta) / 2) / np.sqrt(gamma1)
    aold = 2.0 / (2 + alpha + beta) * np.sqrt((alpha + 1) * (beta + 1) / (alpha + beta + 3))
    for i in range(1, N):
        h1 = 2 * i + alpha + beta
        anew = 2.0 / (h1 + 2) * np.sqrt((i + 1) * (i + 1 + alpha + beta) * (i + 1 + alpha)
                                         * (i + 1 + beta) / (h1 + 1) / (h1 + 3))
        bnew = -(alpha ** 2 - beta ** 2) / h1 / (h1 + 2)
        PL[i + 1] = 1.0 / anew * (-aold * PL[i - 1] + (x - bnew) * PL[i])
        aold = anew
    return PL[N]


def grad_jacobiP(x, alpha: float, beta: float, N: int) -> np.ndarray:
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if N == 0:
        return np.zeros_like(x)
    return np.sqrt(N * (N + alpha + beta + 1.0)) * jacobiP(x, alpha + 1, beta + 1, N - 1)


def vandermonde_1d(q: int, x) -> np.ndarray:
    """Orthonormal Legendre Vandermonde, V[i, j] = P_j(x_i)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.stack([jacobiP(x, 0.0, 0.0, j) for j in range(q + 1)], axis=1)


def grad_vandermonde_1d(q: int, x) -> np.ndarray:
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    return np.stack([grad_jacobiP(x, 0.0, 0.0, j) for j in range(q + 1)], axis=1)
