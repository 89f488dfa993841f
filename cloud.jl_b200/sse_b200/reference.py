"""Reference-element operators (host-side setup, NumPy).

Restates, for the element/approximation types on the hot path, what
``ReferenceApproximation(approx_type, element; ...)`` builds in the reference:

* collapsed-coordinate tensor-product operators on Tri/Tet
  (src/SpatialDiscretizations/tensor_simplex.jl:1-306),
* 1-D nodal (Line) and modal operators used by the 1-D golden tests
  (src/SpatialDiscretizations/tensor_cartesian.jl:1-32,
   src/SpatialDiscretizations/multidimensional.jl:1-40),
* facet nodes / reference normals (src/SpatialDiscretizations/ref_elem_data.jl),
* flux-differencing operators S, C (src/Solvers/operators.jl:163-221),
* ``reference_derivative_operators`` (SpatialDiscretizations.jl:414-418).

Everything here is element-independent and tiny; it stays on the host and is
handed to the CUDA library once through ``sse_create``.
Node ordering follows the reference: tensor nodes are lexicographic with the
*last* collapsed coordinate fastest (quadrature_rules.jl:37-46).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .quadrature import (GaussQuadrature, GaussLobattoQuadrature, LGQuadrature,
                         LGLQuadrature, quadrature_line, jacobiP, vandermonde_1d,
                         grad_vandermonde_1d)


@dataclass(frozen=True)
class NodalTensor:
    p: int


@dataclass(frozen=True)
class ModalTensor:
    p: int


@dataclass(frozen=True)
class ModalMulti:
    p: int


@dataclass(frozen=True)
class NodalMulti:
    p: int


NUM_FACES = {"Line": 2, "Tri": 3, "Tet": 4}
DIM = {"Line": 1, "Tri": 2, "Tet": 3}


# --------------------------------------------------------------------------
# Duffy maps (tensor_simplex.jl:2-11)
# --------------------------------------------------------------------------
def chi_tri(eta1, eta2):
    return 0.5 * (1.0 + eta1) * (1.0 - eta2) - 1.0, eta2


def chi_tet(eta1, eta2, eta3):
    xi_pri1 = 0.5 * (1.0 + eta1) * (1.0 - eta3) - 1.0
    xi_pyr2 = 0.5 * (1.0 + eta2) * (1.0 - eta3) - 1.0
    return 0.5 * (1.0 + xi_pri1) * (1.0 - eta2) - 1.0, xi_pyr2, eta3


def _grid(*x1d):
    """Tensor grid flattened with the first coordinate slowest (reference order)."""
    g = np.meshgrid(*x1d, indexing="ij")
    return [a.reshape(-1) for a in g]


# --------------------------------------------------------------------------
# Warped (Dubiner/Proriol) tensor-product Vandermonde (tensor_simplex.jl:84-140)
# --------------------------------------------------------------------------
@dataclass
class WarpedProduct:
    d: int
    p: int
    A: np.ndarray            # (M1, p+1)
    B: np.ndarray            # (M2, p+1, p+1)
    C: Optional[np.ndarray]  # (M3, p+1, p+1, p+1) for d = 3
    sigma_i: np.ndarray      # modal index (0-based, -1 where unused), shape (p+1,)*d
    sigma_o: np.ndarray      # nodal index (0-based), shape (M1, M2[, M3])

    @property
    def shape(self):
        return (int(self.sigma_o.size), int((self.sigma_i >= 0).sum()))

    def dense(self) -> np.ndarray:
        Nq, Np = self.shape
        V = np.zeros((Nq, Np))
        P1 = self.p + 1
        if self.d == 2:
            for b1 in range(P1):
                for b2 in range(P1):
                    j = self.sigma_i[b1, b2]
                    if j < 0:
                        continue
                    col = np.einsum("a,b->ab", self.A[:, b1], self.B[:, b1, b2])
                    V[self.sigma_o.reshape(-1), j] = col.reshape(-1)
        else:
            for b1 in range(P1):
                for b2 in range(P1):
                    for b3 in range(P1):
                        j = self.sigma_i[b1, b2, b3]
                        if j < 0:
                            continue
                        col = np.einsum("a,b,c->abc", self.A[:, b1], self.B[:, b1, b2],
                                        self.C[:, b1, b2, b3])
                        V[self.sigma_o.reshape(-1), j] = col.reshape(-1)
        return V


def warped_product(elem: str, p: int, eta1d) -> WarpedProduct:
    P1 = p + 1
    if elem == "Tri":
        M1, M2 = len(eta1d[0]), len(eta1d[1])
        sigma_o = np.arange(M1 * M2).reshape(M1, M2)
        sigma_i = -np.ones((P1, P1), dtype=np.int64)
        A = np.zeros((M1, P1))
        B = np.zeros((M2, P1, P1))
        k = 0
        for i in range(P1):
            for j in range(P1 - i):
                sigma_i[i, j] = k
                k += 1
                A[:, i] = np.sqrt(2.0) * jacobiP(eta1d[0], 0, 0, i)
                B[:, i, j] = (1 - eta1d[1]) ** i * jacobiP(eta1d[1], 2 * i + 1, 0, j)
        return WarpedProduct(2, p, A, B, None, sigma_i, sigma_o)
    if elem == "Tet":
        M1, M2, M3 = (len(e) for e in eta1d)
        sigma_o = np.arange(M1 * M2 * M3).reshape(M1, M2, M3)
        sigma_i = -np.ones((P1, P1, P1), dtype=np.int64)
        A = np.zeros((M1, P1))
        B = np.zeros((M2, P1, P1))
        C = np.zeros((M3, P1, P1, P1))
        l = 0
        for i in range(P1):
            for j in range(P1 - i):
                for k in range(P1 - i - j):
                    sigma_i[i, j, k] = l
                    l += 1
                    A[:, i] = np.sqrt(2.0) * jacobiP(eta1d[0], 0, 0, i)
                    B[:, i, j] = (1 - eta1d[1]) ** i * jacobiP(eta1d[1], 2 * i + 1, 0, j)
                    C[:, i, j, k] = (2 * (1 - eta1d[2]) ** (i + j)
                                     * jacobiP(eta1d[2], 2 * i + 2 * j + 2, 0, k))
        return WarpedProduct(3, p, A, B, C, sigma_i, sigma_o)
    raise ValueError(elem)


# --------------------------------------------------------------------------
# Collapsed-coordinate reference metrics (tensor_simplex.jl:13-82)
# --------------------------------------------------------------------------
def reference_geometric_factors(elem: str, rules):
    ab = [(r.a, r.b) for r in rules]
    x1d = [quadrature_line(r)[0] for r in rules]
    eta = _grid(*x1d)
    N = eta[0].size
    d = len(rules)
    Lref = np.zeros((N, d, d))
    if elem == "Tri":
        if ab == [(0, 0), (0, 0)]:
            Jref = 0.5 * (1.0 - eta[1])
            Lref[:, 0, 0] = 1.0
            Lref[:, 0, 1] = 0.5 * (1.0 + eta[0])
            Lref[:, 1, 1] = 0.5 * (1.0 - eta[1])
        elif ab == [(0, 0), (1, 0)]:
            Jref = 0.5 * np.ones(N)
            Lref[:, 0, 0] = 1.0 / (1.0 - eta[1])
            Lref[:, 0, 1] = 0.5 * (1.0 + eta[0]) / (1.0 - eta[1])
            Lref[:, 1, 1] = 0.5
        else:
            raise ValueError("Chosen Jacobi weight not supported")
    elif elem == "Tet":
        if ab == [(0, 0), (0, 0), (0, 0)]:
            h2, h3 = 0.5 * (1.0 - eta[1]), 0.5 * (1.0 - eta[2])
            Jref = h2 * h3 ** 2
            Lref[:, 0, 0] = h3
            Lref[:, 0, 1] = 0.5 * (1.0 + eta[0]) * h3
            Lref[:, 0, 2] = 0.5 * (1.0 + eta[0]) * h3
            Lref[:, 1, 1] = h2 * h3
            Lref[:, 1, 2] = 0.5 * (1.0 + eta[1]) * h2 * h3
            Lref[:, 2, 2] = h2 * h3 ** 2
        elif ab == [(0, 0), (0, 0), (1, 0)]:
            Jref = 0.125 * (1.0 - eta[1]) * (1.0 - eta[2])
            Lref[:, 0, 0] = 0.5
            Lref[:, 0, 1] = 0.25 * (1.0 + eta[0])
            Lref[:, 0, 2] = 0.25 * (1.0 + eta[0])
            Lref[:, 1, 1] = 0.25 * (1.0 - eta[1])
            Lref[:, 1, 2] = 0.125 * (1.0 + eta[1]) * (1.0 - eta[1])
            Lref[:, 2, 2] = 0.125 * (1.0 - eta[1]) * (1.0 - eta[2])
        else:
            raise ValueError("Chosen Jacobi weight not supported")
    else:
        raise ValueError(elem)
    return Jref, Lref


def simplex_quadrature(elem: str, rules):
    """quadrature(::Tri / ::Tet, tensor rules) (quadrature_rules.jl:125-166)."""
    xw = [quadrature_line(r) for r in rules]
    eta = _grid(*[x for x, _ in xw])
    wg = _grid(*[w for _, w in xw])
    ab = [(r.a, r.b) for r in rules]
    if elem == "Tri":
        w2 = wg[0] * wg[1]
        r, s = chi_tri(eta[0], eta[1])
        if ab == [(0, 0), (0, 0)]:
            return (r, s), 0.5 * (1 - eta[1]) * w2
        if ab == [(0, 0), (1, 0)]:
            return (r, s), 0.5 * w2
    if elem == "Tet":
        w3 = wg[0] * wg[1] * wg[2]
        r, s, t = chi_tet(*eta)
        if ab == [(0, 0), (0, 0), (0, 0)]:
            return (r, s, t), 0.125 * (1 - eta[1]) * (1 - eta[2]) ** 2 * w3
        if ab == [(0, 0), (0, 0), (1, 0)]:
            return (r, s, t), 0.125 * (1 - eta[1]) * (1 - eta[2]) * w3
    raise ValueError("Chosen Jacobi weight not supported")


# --------------------------------------------------------------------------
# Degree-N geometry element (what the reference takes from StartUpDG's
# RefElemData: interpolation nodes, Drst, Vq, Vf).  Any unisolvent node set and
# any basis of P_N give the same polynomial mapping space; we use the
# equispaced lattice and a Legendre-product basis (well conditioned for the
# degrees used for mappings, N <= 6).
# --------------------------------------------------------------------------
def _lattice_nodes(d: int, N: int):
    if d == 1:
        return [np.linspace(-1.0, 1.0, N + 1)]
    pts = []
    if d == 2:
        for j in range(N + 1):
            for i in range(N + 1 - j):
                pts.append((-1 + 2.0 * i / N, -1 + 2.0 * j / N))
    else:
        for k in range(N + 1):
            for j in range(N + 1 - k):
                for i in range(N + 1 - j - k):
                    pts.append((-1 + 2.0 * i / N, -1 + 2.0 * j / N, -1 + 2.0 * k / N))
    pts = np.array(pts)
    return [pts[:, m].copy() for m in range(d)]


def _warp_blend_nodes_tri(N: int):
    """Interpolation nodes of the triangle mapping element: equispaced lattice warped so that every edge carries the
    Gauss-Lobatto nodes, blended into the interior without the alpha-optimisation of Hesthaven & Warburton (alpha = 0).
    This is the node set that reproduces the reference's triangle goldens (runtests.jl:38-60, 111-121) to round-off,
    i.e. what the un-vendored NodesAndModes `nodes(Tri(), N)` returns; the curved geometry is the interpolant of the
    warping function at these nodes, so the set matters whenever the mapping is not polynomial."""
    if N == 1:
        return [np.array([-1.0, 1.0, -1.0]), np.array([-1.0, -1.0, 1.0])]
    gll, _ = quadrature_line(GaussLobattoQuadrature(N))
    req = np.linspace(-1.0, 1.0, N + 1)
    Veq = vandermonde_1d(N, req)

    def warpfactor(rout):
        warp = np.linalg.solve(Veq.T, vandermonde_1d(N, rout).T).T @ (gll - req)
        inside = (np.abs(rout) < 1.0 - 1e-10).astype(float)
        return warp / (1.0 - (inside * rout) ** 2) + warp * (inside - 1.0)

    L1, L3 = [], []
    for n in range(N + 1):
        for m in range(N + 1 - n):
            L1.append(n / N)
            L3.append(m / N)
    L1, L3 = np.array(L1), np.array(L3)
    L2 = 1.0 - L1 - L3
    x, y = -L2 + L3, (-L2 - L3 + 2.0 * L1) / np.sqrt(3.0)
    w1 = 4.0 * L2 * L3 * warpfactor(L3 - L2)
    w2 = 4.0 * L1 * L3 * warpfactor(L1 - L3)
    w3 = 4.0 * L1 * L2 * warpfactor(L2 - L1)
    x = x + w1 + np.cos(2 * np.pi / 3) * w2 + np.cos(4 * np.pi / 3) * w3
    y = y + np.sin(2 * np.pi / 3) * w2 + np.sin(4 * np.pi / 3) * w3
    L1 = (np.sqrt(3.0) * y + 1.0) / 3.0
    L2 = (-3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    L3 = (3.0 * x - np.sqrt(3.0) * y + 2.0) / 6.0
    return [-L2 + L3 - L1, -L2 - L3 + L1]


def _warp_blend_nodes_tet(N: int):
    """Interpolation nodes of the tetrahedron mapping element: Hesthaven & Warburton's warp-and-blend construction (Nodes3D)
    with alpha = 0, the 3-D counterpart of _warp_blend_nodes_tri — Gauss-Lobatto nodes on the edges, the triangle set on
    every face, symmetric under all vertex permutations (tests/test_reference_operators.py)."""
    if N == 1:
        return [np.array([-1.0, 1.0, -1.0, -1.0]), np.array([-1.0, -1.0, 1.0, -1.0]), np.array([-1.0, -1.0, -1.0, 1.0])]
    gll, _ = quadrature_line(GaussLobattoQuadrature(N))
    gx = np.sort(gll)[::-1]
    xeq = np.array([-1.0 + 2.0 * (N - i) / N for i in range(N + 1)])       # descending, like gx

    def evalwarp(xout):                         # interpolant of (gx - xeq) divided by (1 - x^2)
        warp = np.zeros_like(xout)
        for i in range(N + 1):
            dd = (gx[i] - xeq[i]) * np.ones_like(xout)
            for j in range(1, N):
                if i != j:
                    dd = dd * (xout - xeq[j]) / (xeq[i] - xeq[j])
            if i != 0:
                dd = -dd / (xeq[i] - xeq[0])
            if i != N:
                dd = dd / (xeq[i] - xeq[N])
            warp = warp + dd
        return warp

    def evalshift(L1, L2, L3):
        w1, w2, w3 = (L2 * L3 * 4.0 * evalwarp(L3 - L2), L1 * L3 * 4.0 * evalwarp(L1 - L3), L1 * L2 * 4.0 * evalwarp(L2 - L1))
        return (w1 + np.cos(2 * np.pi / 3) * w2 + np.cos(4 * np.pi / 3) * w3,
                np.sin(2 * np.pi / 3) * w2 + np.sin(4 * np.pi / 3) * w3)

    r, s, t = _lattice_nodes(3, N)
    L1, L2, L3, L4 = (1 + t) / 2, (1 + s) / 2, -(1 + r + s + t) / 2, (1 + r) / 2
    v1 = np.array([-1.0, -1 / np.sqrt(3), -1 / np.sqrt(6)])
    v2 = np.array([1.0, -1 / np.sqrt(3), -1 / np.sqrt(6)])
    v3 = np.array([0.0, 2 / np.sqrt(3), -1 / np.sqrt(6)])
    v4 = np.array([0.0, 0.0, 3 / np.sqrt(6)])
    t1 = [v2 - v1, v2 - v1, v3 - v2, v3 - v1]
    t2 = [v3 - 0.5 * (v1 + v2), v4 - 0.5 * (v1 + v2), v4 - 0.5 * (v2 + v3), v4 - 0.5 * (v1 + v3)]
    t1 = [a / np.linalg.norm(a) for a in t1]
    t2 = [a / np.linalg.norm(a) for a in t2]
    XYZ = np.outer(L3, v1) + np.outer(L4, v2) + np.outer(L2, v3) + np.outer(L1, v4)
    shift = np.zeros_like(XYZ)
    tol = 1e-10
    for face, (La, Lb, Lc, Ld) in enumerate([(L1, L2, L3, L4), (L2, L1, L3, L4), (L3, L1, L4, L2), (L4, L1, L3, L2)]):
        w1, w2 = evalshift(Lb, Lc, Ld)
        blend = Lb * Lc * Ld
        denom = (Lb + 0.5 * La) * (Lc + 0.5 * La) * (Ld + 0.5 * La)
        ok = denom > tol
        blend[ok] = blend[ok] / denom[ok]
        shift = shift + np.outer(blend * w1, t1[face]) + np.outer(blend * w2, t2[face])
        onface = (La < tol) & (((Lb > tol).astype(int) + (Lc > tol).astype(int) + (Ld > tol).astype(int)) < 3)
        shift[onface] = np.outer(w1[onface], t1[face]) + np.outer(w2[onface], t2[face])
    XYZ = XYZ + shift
    A = np.stack([0.5 * (v2 - v1), 0.5 * (v3 - v1), 0.5 * (v4 - v1)], axis=1)
    RST = np.linalg.solve(A, XYZ.T - (0.5 * (v2 + v3 + v4 - v1))[:, None])
    return [RST[0].copy(), RST[1].copy(), RST[2].copy()]


def _tensor_nodes(d: int, N: int):
    """Tensor-product Gauss-Lobatto interpolation nodes of the Quad / Hex mapping element (NodesAndModes' nodes(Quad/Hex, N)
    are the tensor product of the 1-D Lobatto nodes), first coordinate slowest."""
    x, _ = quadrature_line(GaussLobattoQuadrature(N))
    g = np.meshgrid(*([x] * d), indexing="ij")
    return [a.reshape(-1).copy() for a in g]


def _poly_basis(d: int, N: int, rst, grad: bool = False, tensor: bool = False):
    """Total-degree-N basis prod_m P_{a_m}(x_m), sum a_m <= N (simplices), or the full tensor basis a_m <= N
    (tensor=True: Quad / Hex), and its gradients."""
    from numpy.polynomial import legendre as L
    n = rst[0].size
    idx = []
    if tensor:
        from itertools import product
        idx = list(product(range(N + 1), repeat=d))
    elif d == 1:
        idx = [(a,) for a in range(N + 1)]
    elif d == 2:
        idx = [(a, b) for a in range(N + 1) for b in range(N + 1 - a)]
    else:
        idx = [(a, b, c) for a in range(N + 1) for b in range(N + 1 - a)
               for c in range(N + 1 - a - b)]
    P = [[None] * (N + 1) for _ in range(d)]
    dP = [[None] * (N + 1) for _ in range(d)]
    for m in range(d):
        for a in range(N + 1):
            c = np.zeros(a + 1)
            c[a] = 1.0
            P[m][a] = L.legval(rst[m], c)
            dP[m][a] = L.legval(rst[m], L.legder(c)) if a > 0 else np.zeros(n)
    V = np.empty((n, len(idx)))
    G = [np.empty((n, len(idx))) for _ in range(d)]
    for j, a in enumerate(idx):
        vals = [P[m][a[m]] for m in range(d)]
        V[:, j] = np.prod(vals, axis=0)
        if grad:
            for m in range(d):
                g = dP[m][a[m]].copy()
                for mm in range(d):
                    if mm != m:
                        g = g * vals[mm]
                G[m][:, j] = g
    return (V, G) if grad else V


@dataclass
class GeometryElement:
    """Degree-N nodal element used only for the mapping (mesh.xyz, Drst, Vq, Vf)."""
    d: int
    N: int
    rst: List[np.ndarray]
    VDM: np.ndarray
    Drst: List[np.ndarray]
    Vq: np.ndarray
    Vf: np.ndarray
    tensor: bool = False

    def interp(self, pts) -> np.ndarray:
        return np.linalg.solve(self.VDM.T, _poly_basis(self.d, self.N, pts, tensor=self.tensor).T).T


def geometry_element(d: int, N: int, rstq, rstf, tensor: bool = False) -> GeometryElement:
    rst = _tensor_nodes(d, N) if tensor else (_warp_blend_nodes_tri(N) if d == 2 else
                                              (_warp_blend_nodes_tet(N) if d == 3 else _lattice_nodes(d, N)))
    VDM, G = _poly_basis(d, N, rst, grad=True, tensor=tensor)
    Drst = [np.linalg.solve(VDM.T, g.T).T for g in G]
    ge = GeometryElement(d, N, rst, VDM, Drst, None, None, tensor)
    ge.Vq = ge.interp(rstq)
    ge.Vf = ge.interp(rstf)
    return ge


# --------------------------------------------------------------------------
# ReferenceApproximation
# --------------------------------------------------------------------------
@dataclass
class ReferenceApproximation:
    approx_type: object
    element: str
    d: int
    p: int
    N_p: int
    N_q: int
    N_f: int
    N_fac: int
    D: List[np.ndarray]                 # collapsed-coordinate derivative operators (N_q x N_q)
    V: np.ndarray                       # dense generalized Vandermonde (N_q x N_p)
    V_warped: Optional[WarpedProduct]   # sum-factorised form of V (ModalTensor) or None
    V_is_identity: bool
    R: np.ndarray                       # dense facet interpolation (N_f x N_q)
    W: np.ndarray                       # (N_q,)
    B: np.ndarray                       # (N_f,)
    J_ref: Optional[np.ndarray]         # reference (collapse) mapping, None = NoMapping
    L_ref: Optional[np.ndarray]         # (N_q, d, d)
    rstq: List[np.ndarray]
    rstf: List[np.ndarray]
    nrstJ: List[np.ndarray]             # reference normals at facet nodes
    geom: GeometryElement
    eta1d: List[np.ndarray] = field(default_factory=list)
    D1d: List[np.ndarray] = field(default_factory=list)
    is_tensor: bool = True
    R_is_selection: bool = False

    @property
    def nodes_per_face(self) -> int:
        return self.N_f // self.N_fac

    # reference_derivative_operators (SpatialDiscretizations.jl:414-422)
    def D_xi(self) -> List[np.ndarray]:
        if self.J_ref is None:
            return self.D
        d = self.d
        return [sum((self.L_ref[:, l, m] / self.J_ref)[:, None] * self.D[l] for l in range(d))
                for m in range(d)]

    # flux_differencing_operators (operators.jl:163-221)
    def flux_differencing_operators(self):
        Dxi = self.D_xi()
        W = self.W
        S = [0.5 * (W[:, None] * Dm - Dm.T * W[None, :]) for Dm in Dxi]
        C = None if self.R_is_selection else self.R.T * self.B[None, :]
        return S, C

    # check_sbp_property (SpatialDiscretizations.jl:441-455)
    def check_sbp_property(self):
        Dxi = self.D_xi()
        out = []
        for m in range(self.d):
            Q = self.W[:, None] * Dxi[m]
            E = self.R.T @ ((self.B * self.nrstJ[m])[:, None] * self.R)
            out.append(np.max(np.abs(Q + Q.T - E)))
        return out


def _ops_1d(rules):
    eta, q, V1, D1, RL, RR = [], [], [], [], [], []
    for r in rules:
        x, _ = quadrature_line(r)
        eta.append(x)
        q.append(len(x) - 1)
        V = vandermonde_1d(q[-1], x)
        V1.append(V)
        D1.append(np.linalg.solve(V.T, grad_vandermonde_1d(q[-1], x).T).T)
        RL.append(np.linalg.solve(V.T, vandermonde_1d(q[-1], [-1.0]).T).T)
        RR.append(np.linalg.solve(V.T, vandermonde_1d(q[-1], [1.0]).T).T)
    return eta, q, V1, D1, RL, RR


def _interp_1d(rule_from, rule_to, q, V1):
    if rule_from == rule_to:
        return np.eye(q + 1)
    xf, _ = quadrature_line(rule_to)
    return np.linalg.solve(V1.T, vandermonde_1d(q, xf).T).T


def _kron(*mats):
    out = mats[0]
    for m in mats[1:]:
        out = np.kron(out, m)
    return out


def reference_approximation(approx_type, element: str, mapping_degree: int = 1,
                            volume_quadrature_rule=None, facet_quadrature_rule=None
                            ) -> ReferenceApproximation:
    p = approx_type.p
    if element in ("Tri", "Tet") and isinstance(approx_type, (ModalMulti, NodalMulti)):
        return _ref_multi(approx_type, element, p, mapping_degree, volume_quadrature_rule, facet_quadrature_rule)
    if element == "Tri":
        return _ref_tri(approx_type, p, mapping_degree, volume_quadrature_rule,
                        facet_quadrature_rule)
    if element == "Tet":
        return _ref_tet(approx_type, p, mapping_degree, volume_quadrature_rule,
                        facet_quadrature_rule)
    if element == "Line":
        return _ref_line(approx_type, p, mapping_degree, volume_quadrature_rule)
    if element in ("Quad", "Hex"):
        return _ref_box(approx_type, element, p, mapping_degree, volume_quadrature_rule, facet_quadrature_rule)
    raise ValueError(f"unsupported element {element}")


# multidimensional.jl:1-75 -- ModalMulti / NodalMulti: dense operators D_m = grad(VDM)_m P, R = V_f P (and V = VDM or I) with
# P = (VDM' W VDM)^-1 VDM' W, on a simplex without a collapsed reference mapping (NoMapping).  The reference takes its volume /
# facet nodes from StartUpDG's tabulated symmetric rules of degree 2p (not vendored); here the collapsed Gauss rules of the
# tensor schemes stand in for them -- they are exact to degree 2p as well, so every operator identity the constructors rely
# on holds (SBP property, exact differentiation and extrapolation of P_p), while the node SET differs from the reference's.
def _ref_multi(approx_type, element, p, mapping_degree, vq, fq):
    t = (_ref_tri if element == "Tri" else _ref_tet)(ModalTensor(p), p, mapping_degree, vq, fq)
    VDM, W = t.V, t.W
    grad = [Dm @ VDM for Dm in t.D_xi()]                      # basis gradients at the volume nodes (D_xi is exact on P_p)
    Vf = t.R @ VDM                                            # basis at the facet nodes (R is exact on P_p)
    P = np.linalg.solve(VDM.T @ (W[:, None] * VDM), VDM.T * W[None, :])
    D = [g @ P for g in grad]
    nodal = isinstance(approx_type, NodalMulti)
    V = np.eye(W.size) if nodal else VDM
    return ReferenceApproximation(approx_type, element, t.d, p, V.shape[1], V.shape[0], t.N_f, t.N_fac,
                                  D, V, None, nodal, Vf @ P, W, t.B, None, None, t.rstq, t.rstf, t.nrstJ, t.geom,
                                  is_tensor=False)


# tensor_simplex.jl:158-219
def _ref_tri(approx_type, p, mapping_degree, vq, fq):
    vq = vq or (LGQuadrature(p), LGQuadrature(p))
    fq = fq or LGQuadrature(p)
    eta, q, V1, D1, RL, RR = _ops_1d(vq)
    J_ref, L_ref = reference_geometric_factors("Tri", vq)
    e1f = _interp_1d(vq[0], fq, q[0], V1[0])
    e2f = _interp_1d(vq[1], fq, q[1], V1[1])
    R = np.vstack([_kron(e1f, RL[1]), _kron(RR[0], e2f), _kron(RL[0], e2f)])
    I1, I2 = np.eye(q[0] + 1), np.eye(q[1] + 1)
    D = [_kron(D1[0], I2), _kron(I1, D1[1])]
    # ref_elem_data.jl:1-64
    r1, w1 = quadrature_line(fq)
    one = np.ones_like(r1)
    rf = np.concatenate([r1, -r1, -one])
    sf = np.concatenate([-one, r1, r1])
    wf = np.concatenate([w1, w1, w1])
    nrJ = np.concatenate([0 * one, one, -one])
    nsJ = np.concatenate([-one, one, 0 * one])
    (rq, sq), wq = simplex_quadrature("Tri", vq)
    geom = geometry_element(2, mapping_degree, [rq, sq], [rf, sf])
    if isinstance(approx_type, ModalTensor):
        Vw = warped_product("Tri", p, eta)
        V, ident = Vw.dense(), False
    else:
        Vw, V, ident = None, np.eye(rq.size), True
    return ReferenceApproximation(approx_type, "Tri", 2, p, V.shape[1], V.shape[0], R.shape[0], 3,
                                  D, V, Vw, ident, R, wq, wf, J_ref, L_ref, [rq, sq], [rf, sf],
                                  [nrJ, nsJ], geom, eta, D1)


# tensor_simplex.jl:221-306
def _ref_tet(approx_type, p, mapping_degree, vq, fq):
    vq = vq or (LGQuadrature(p), LGQuadrature(p), GaussQuadrature(p, 1, 0))
    fq = fq or (LGQuadrature(p), GaussQuadrature(p, 1, 0))
    eta, q, V1, D1, RL, RR = _ops_1d(vq)
    J_ref, L_ref = reference_geometric_factors("Tet", vq)
    e1f1 = _interp_1d(vq[0], fq[0], q[0], V1[0])
    e2f1 = _interp_1d(vq[1], fq[0], q[1], V1[1])
    e2f2 = _interp_1d(vq[1], fq[1], q[1], V1[1])
    e3f2 = _interp_1d(vq[2], fq[1], q[2], V1[2])
    R = np.vstack([_kron(e1f1, RL[1], e3f2), _kron(RR[0], e2f1, e3f2),
                   _kron(RL[0], e2f1, e3f2), _kron(e1f1, e2f2, RL[2])])
    I = [np.eye(n + 1) for n in q]
    D = [_kron(D1[0], I[1], I[2]), _kron(I[0], D1[1], I[2]), _kron(I[0], I[1], D1[2])]
    # ref_elem_data.jl:66-134
    (r2, s2), w2 = simplex_quadrature("Tri", fq)
    ee, zz = np.ones_like(r2), np.zeros_like(r2)
    rf = np.concatenate([r2, -(ee + r2 + s2), -ee, r2])
    sf = np.concatenate([-ee, r2, r2, s2])
    tf = np.concatenate([s2, s2, s2, -ee])
    wf = np.concatenate([w2, w2, w2, w2])
    nrJ = np.concatenate([zz, ee, -ee, zz])
    nsJ = np.concatenate([-ee, ee, zz, zz])
    ntJ = np.concatenate([zz, ee, zz, -ee])
    (rq, sq, tq), wq = simplex_quadrature("Tet", vq)
    geom = geometry_element(3, mapping_degree, [rq, sq, tq], [rf, sf, tf])
    if isinstance(approx_type, ModalTensor):
        Vw = warped_product("Tet", p, eta)
        V, ident = Vw.dense(), False
    else:
        Vw, V, ident = None, np.eye(rq.size), True
    return ReferenceApproximation(approx_type, "Tet", 3, p, V.shape[1], V.shape[0], R.shape[0], 4,
                                  D, V, Vw, ident, R, wq, wf, J_ref, L_ref, [rq, sq, tq],
                                  [rf, sf, tf], [nrJ, nsJ, ntJ], geom, eta, D1)


def _ref_line(approx_type, p, mapping_degree, vq):
    rf = np.array([-1.0, 1.0])
    nrJ = np.array([-1.0, 1.0])
    wf = np.array([1.0, 1.0])
    if isinstance(approx_type, NodalTensor):
        # tensor_cartesian.jl:1-32
        vq = vq or LGLQuadrature(p)
        rq, wq = quadrature_line(vq)
        q = len(rq) - 1
        VDM = vandermonde_1d(q, rq)
        D = np.linalg.solve(VDM.T, grad_vandermonde_1d(q, rq).T).T
        R = np.linalg.solve(VDM.T, vandermonde_1d(q, rf).T).T
        sel = isinstance(vq, GaussLobattoQuadrature)
        if sel:
            R = np.round(R)  # exact selection of the end points
        V = np.eye(q + 1)
        geom = geometry_element(1, mapping_degree, [rq], [rf])
        return ReferenceApproximation(NodalTensor(q), "Line", 1, q, q + 1, q + 1, 2, 2, [D], V, None,
                                      True, R, wq, wf, None, None, [rq], [rf], [nrJ], geom, [rq],
                                      [D], True, sel)
    if isinstance(approx_type, ModalMulti):
        # multidimensional.jl:1-40 with DefaultQuadrature(2p) on a Line =
        # LG(ceil((2p-1)/2)) (quadrature_rules.jl:49-51)
        vq = vq or LGQuadrature(int(np.ceil((2 * p - 1) / 2)))
        rq, wq = quadrature_line(vq)
        VDM = vandermonde_1d(p, rq)
        dVDM = grad_vandermonde_1d(p, rq)
        Vf = vandermonde_1d(p, rf)
        P = np.linalg.solve(VDM.T @ (wq[:, None] * VDM), VDM.T * wq[None, :])
        geom = geometry_element(1, mapping_degree, [rq], [rf])
        return ReferenceApproximation(approx_type, "Line", 1, p, p + 1, rq.size, 2, 2, [dVDM @ P],
                                      VDM, None, False, Vf @ P, wq, wf, None, None, [rq], [rf],
                                      [nrJ], geom, [rq], [], False, False)
    raise ValueError(approx_type)


# tensor_cartesian.jl:34-131 — NodalTensor on Quad / Hex
def _ref_box(approx_type, element, p, mapping_degree, vq, fq):
    """Collocated tensor-product element: V = I, D_m = I (x) D_1D (x) I, W = w (x) w [(x) w]; with Lobatto rules R is a
    selection of the boundary nodes (diagonal-E), otherwise the Kronecker product of 1-D extrapolations.  Node index =
    (i_1 * n + i_2) [* n + i_3]; faces are ordered (xi_1 = -1, xi_1 = +1, xi_2 = -1, ...) and their nodes run over the
    remaining coordinates in the same order (any consistent convention gives the same scheme)."""
    if not isinstance(approx_type, NodalTensor):
        raise ValueError("Quad / Hex elements carry NodalTensor approximations (tensor_cartesian.jl)")
    d = 2 if element == "Quad" else 3
    vq = vq or LGLQuadrature(p)
    fq = fq or vq
    if fq != vq:
        raise NotImplementedError("distinct facet quadrature on Quad / Hex")
    x1, w1 = quadrature_line(vq)
    q = len(x1) - 1
    n = q + 1
    V1 = vandermonde_1d(q, x1)
    D1 = np.linalg.solve(V1.T, grad_vandermonde_1d(q, x1).T).T
    sel = isinstance(vq, GaussLobattoQuadrature)
    RL = np.linalg.solve(V1.T, vandermonde_1d(q, [-1.0]).T).T
    RR = np.linalg.solve(V1.T, vandermonde_1d(q, [1.0]).T).T
    if sel:
        RL, RR = np.round(RL), np.round(RR)
    I1 = np.eye(n)
    D = [_kron(*[D1 if mm == m else I1 for mm in range(d)]) for m in range(d)]
    R = np.vstack([_kron(*[(RL if side == 0 else RR) if mm == m else I1 for mm in range(d)])
                   for m in range(d) for side in (0, 1)])
    grid = np.meshgrid(*([x1] * d), indexing="ij")
    rstq = [a.reshape(-1).copy() for a in grid]
    wq = _kron(*([w1[:, None]] * d)).reshape(-1)
    fgrid = np.meshgrid(*([x1] * (d - 1)), indexing="ij")
    fpts = [a.reshape(-1) for a in fgrid]
    wface = _kron(*([w1[:, None]] * (d - 1))).reshape(-1)
    npf = wface.size
    rstf = [[] for _ in range(d)]
    nrstJ = [[] for _ in range(d)]
    for m in range(d):
        for side in (-1.0, 1.0):
            others = iter(fpts)
            for mm in range(d):
                rstf[mm].append(np.full(npf, side) if mm == m else next(others))
                nrstJ[mm].append(np.full(npf, side) if mm == m else np.zeros(npf))
    rstf = [np.concatenate(a) for a in rstf]
    nrstJ = [np.concatenate(a) for a in nrstJ]
    wf = np.tile(wface, 2 * d)
    geom = geometry_element(d, mapping_degree, rstq, rstf, tensor=True)
    return ReferenceApproximation(NodalTensor(q), element, d, q, n ** d, n ** d, R.shape[0], 2 * d, D, np.eye(n ** d), None,
                                  True, R, wq, wf, None, None, rstq, rstf, nrstJ, geom, [x1] * d, [D1] * d, True, sel)
