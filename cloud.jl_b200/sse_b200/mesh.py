"""Periodic simplex meshes, curvilinear warpings and geometric factors (host setup).

Restates src/SpatialDiscretizations/mesh.jl for the element types on the hot
path.  Python-facing arrays are C-ordered with the *element index first*; the
memory image is therefore exactly the reference's column-major array with the
element index last, e.g. ``Lambda_q[k, n, m, i]`` here is ``Λ_q[i, m, n, k]``
(SpatialDiscretizations.jl:283-289) in the reference.

* ``uniform_periodic_mesh``  mesh.jl:133-179 (incl. the descending-vertex-id tet
  orientation of Warburton's thesis, mesh.jl:150-169)
* ``warp_mesh``              mesh.jl:23-119
* ``geometric_factors``      mesh.jl:229-282 (exact), 284-339 / 410-506 (conservative curl)
* ``mapP``                   StartUpDG ``make_periodic`` semantics: 0-based linear
  index ``k * N_f + i`` of the coincident facet node of the neighbour.
"""
from __future__ import annotations

from dataclasses import dataclass
from itertools import permutations
from typing import List, Optional, Tuple

import numpy as np

from .reference import ReferenceApproximation, geometry_element, GeometryElement


# --------------------------------------------------------------------------
# warpings (mesh.jl:7-119)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class DelReyWarping:
    factor: float
    L: Tuple[float, ...]


@dataclass(frozen=True)
class ChanWarping:
    factor: float
    L: Tuple[float, ...]


@dataclass(frozen=True)
class UniformWarping:
    factor: float
    L: Tuple[float, ...]


def apply_warp(xyz: List[np.ndarray], w) -> List[np.ndarray]:
    d = len(xyz)
    pi = np.pi
    f, L = (w.factor, w.L) if w is not None else (0.0, None)
    if w is None:
        return [a.copy() for a in xyz]
    if isinstance(w, DelReyWarping) and d == 2:
        x, y = xyz
        xn = x + L[0] * f * np.sin(pi * x / L[0]) * np.sin(pi * y / L[1])
        yn = y + L[1] * f * np.exp(1.0 - y / L[1]) * np.sin(pi * x / L[0]) * np.sin(pi * y / L[1])
        return [xn, yn]
    if isinstance(w, DelReyWarping) and d == 3:
        x, y, z = xyz
        xn = x + L[0] * f * np.sin(pi * x / L[0]) * np.sin(pi * y / L[1])
        yn = y + L[1] * f * np.exp((1.0 - y) / L[1]) * np.sin(pi * x / L[0]) * np.sin(pi * y / L[1])
        zn = z + 0.25 * L[2] * f * (np.sin(2 * pi * x / L[0]) * np.sin(2 * pi * y / L[1])) \
            * np.sin(2 * pi * z / L[2])
        return [xn, yn, zn]
    if isinstance(w, ChanWarping) and d == 2:
        x, y = xyz
        xn = x + L[0] * f * np.cos(pi / L[0] * (x - 0.5 * L[0])) * np.cos(3 * pi / L[1] * (y - 0.5 * L[1]))
        yn = y + L[1] * f * np.sin(4 * pi / L[0] * (xn - 0.5 * L[0])) * np.cos(pi / L[1] * (y - 0.5 * L[1]))
        return [xn, yn]
    if isinstance(w, ChanWarping) and d == 3:
        x, y, z = xyz
        yn = y + L[1] * f * np.cos(3 * pi / L[0] * (x - 0.5 * L[0])) \
            * np.cos(pi / L[1] * (y - 0.5 * L[1])) * np.cos(pi / L[2] * (z - 0.5 * L[2]))
        xn = x + L[0] * f * np.cos(pi / L[0] * (x - 0.5 * L[0])) \
            * np.sin(4 * pi / L[1] * (yn - 0.5 * L[1])) * np.cos(pi / L[2] * (z - 0.5 * L[2]))
        zn = z + L[2] * f * np.cos(pi / L[0] * (xn - 0.5 * L[0])) \
            * np.cos(2 * pi / L[1] * (yn - 0.5 * L[1])) * np.cos(pi / L[2] * (z - 0.5 * L[2]))
        return [xn, yn, zn]
    if isinstance(w, UniformWarping):
        eps = f * np.prod([np.sin(2 * pi * (xyz[m] - L[m] / 2) / L[m]) for m in range(d)], axis=0)
        return [xyz[m] + L[m] * eps for m in range(d)]
    raise ValueError(f"unsupported warping {w!r} for d={d}")


# --------------------------------------------------------------------------
# mesh container
# --------------------------------------------------------------------------
@dataclass
class Mesh:
    d: int
    N_e: int                    # locally owned elements
    xyz: List[np.ndarray]       # (N_e, N_nodes) mapping-node coordinates (curved)
    xyzq: List[np.ndarray]      # (N_e, N_q)
    xyzf: List[np.ndarray]      # (N_e, N_f)
    mapP: np.ndarray            # (N_e, N_f) int64, 0-based linear index k*N_f + i (local + ghost)
    limits: Tuple[Tuple[float, float], ...]
    # distributed meshes: ghost facet slots appended after the N_e local elements
    n_ghost: int = 0
    send_idx: Optional[List[np.ndarray]] = None   # per neighbour rank: local facet-node linear ids to send
    recv_off: Optional[List[int]] = None          # per neighbour rank: ghost slot offset (in facet nodes)
    nbr_ranks: Optional[List[int]] = None
    n_boundary: int = 0                           # elements [N_e - n_boundary, N_e) touch ghosts
    elem_gid: Optional[np.ndarray] = None         # global element ids of the local elements


def _affine_nodes(VX, EtoV, geom: GeometryElement):
    """Map the reference interpolation nodes affinely into each straight element."""
    d = geom.d
    if d == 1:
        lam = [0.5 * (1 - geom.rst[0]), 0.5 * (1 + geom.rst[0])]
    elif d == 2:
        r, s = geom.rst
        lam = [-0.5 * (r + s), 0.5 * (1 + r), 0.5 * (1 + s)]
    else:
        r, s, t = geom.rst
        lam = [-0.5 * (1 + r + s + t), 0.5 * (1 + r), 0.5 * (1 + s), 0.5 * (1 + t)]
    return [sum(VX[m][EtoV[:, v]][:, None] * lam[v][None, :] for v in range(d + 1))
            for m in range(d)]


def _cartesian_simplex_cells(d: int, M: Tuple[int, ...], cells: np.ndarray):
    """Vertices (non-periodic ids) and EtoV for the Kuhn split of the given cells.

    Vertex id = ix + (Mx+1)*(iy + (My+1)*iz); translation preserves the order
    of ids, so periodic images of a face see the same relative vertex order.
    Local vertex order per simplex: global ids descending, first two swapped if
    the orientation is negative (mesh.jl:150-169; used for every simplex so that
    collapsed facet nodes conform).
    """
    n1 = [m + 1 for m in M]
    if d == 2:
        ix, iy = cells % M[0], cells // M[0]
        base = np.stack([ix, iy], axis=1)
    else:
        ix, iy, iz = cells % M[0], (cells // M[0]) % M[1], cells // (M[0] * M[1])
        base = np.stack([ix, iy, iz], axis=1)

    def vid(c):
        if d == 2:
            return c[:, 0] + n1[0] * c[:, 1]
        return c[:, 0] + n1[0] * (c[:, 1] + n1[1] * c[:, 2])

    if d == 2:
        # StartUpDG's uniform_mesh(Tri(), Kx, Ky), identified through the reference's triangle goldens (runtests.jl:38-60,
        # 111-121 are reproduced to round-off with exactly this split): every square is cut along the lower-left /
        # upper-right diagonal into (ll, lr, ur) and (ur, ul, ll); the reference keeps this order for triangles.
        ll, lr = vid(base), vid(base + [1, 0])
        ur, ul = vid(base + [1, 1]), vid(base + [0, 1])
        EtoV = np.stack([np.stack([ll, lr, ur], axis=1), np.stack([ur, ul, ll], axis=1)], axis=1).reshape(-1, 3)
        return EtoV, n1
    # StartUpDG's uniform_mesh(Tet(), Kx, Ky, Kz), identified through the reference's tetrahedral golden (runtests.jl:123-129):
    # six tets per cube around the body diagonal from (0, 0, 1) to (1, 1, 0).  With this diagonal the golden L2 error
    # 0.1876141674772107 is reproduced to 4e-5 (the rest is the error quadrature: the un-vendored Jaskowiec-Sukumar rule
    # against a converged collapsed Gauss rule, tests/test_oracle_goldens.py); the other three diagonals are off by 1-3e-2.
    start = np.array([0, 0, 1])
    tets = []
    for perm in permutations(range(d)):
        c = base + start[None, :]
        verts = [vid(c)]
        for ax in perm:
            c = c.copy()
            c[:, ax] += 1 - 2 * start[ax]
            verts.append(vid(c))
        tets.append(np.stack(verts, axis=1))
    EtoV = np.stack(tets, axis=1).reshape(-1, d + 1)       # element = cell * d! + t
    return EtoV, n1


def _vertex_coords(d, n1, limits, ids):
    out = []
    if d == 2:
        comps = [ids % n1[0], ids // n1[0]]
    else:
        comps = [ids % n1[0], (ids // n1[0]) % n1[1], ids // (n1[0] * n1[1])]
    for m in range(d):
        lo, hi = limits[m]
        out.append(lo + (hi - lo) * comps[m] / (n1[m] - 1))
    return out


def _orient(d, EtoV, n1, limits):
    if d == 2:
        return EtoV                                        # triangles keep the generator's vertex order
    E = -np.sort(-EtoV, axis=1)                            # descending global ids
    X = [_vertex_coords(d, n1, limits, E[:, v]) for v in range(d + 1)]
    A = np.stack([np.stack([X[v][m] - X[0][m] for m in range(d)], axis=1)
                  for v in range(1, d + 1)], axis=2)       # (N_e, d, d) columns = edges
    neg = np.linalg.det(A) < 0
    E[neg, 0], E[neg, 1] = E[neg, 1].copy(), E[neg, 0].copy()
    return E


def _match_faces(xf: List[np.ndarray], N_fac: int, limits, tol=1e-8):
    """mapP by coordinate matching of facet nodes of the *straight* periodic mesh.

    xf[m]: (N_e, N_f).  Returns (N_e, N_f) linear indices k*N_f + i.
    """
    d = len(xf)
    N_e, N_f = xf[0].shape
    npf = N_f // N_fac
    L = np.array([hi - lo for lo, hi in limits])
    lo = np.array([l for l, _ in limits])
    X = np.stack(xf, axis=2).reshape(N_e * N_fac, npf, d)          # faces
    cen = np.mod(X.mean(axis=1) - lo, L)
    binw = L * 1e-6
    for shift in (0.0, 0.5, 0.25, 0.75):
        key = np.floor(cen / binw + shift).astype(np.int64)
        key = np.mod(key, np.round(L / binw).astype(np.int64))     # x = L  ==  x = 0
        _, inv, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)
        if np.all(counts == 2):
            break
    else:
        raise RuntimeError("face matching failed: mesh is not a watertight periodic mesh")
    order = np.argsort(inv.reshape(-1), kind="stable")
    a, b = order[0::2], order[1::2]
    mapP = np.empty(N_e * N_f, dtype=np.int64)
    CH = 200000
    for s in range(0, a.size, CH):
        fa, fb = a[s:s + CH], b[s:s + CH]
        Xa, Xb = X[fa], X[fb]
        diff = Xa[:, :, None, :] - Xb[:, None, :, :]
        diff = diff - L * np.round(diff / L)
        dist = np.abs(diff).sum(axis=3)
        ja = np.argmin(dist, axis=2)                                 # for each node of a, node of b
        if np.max(np.take_along_axis(dist, ja[:, :, None], 2)) > 1e-8 * L.max():
            raise RuntimeError("facet nodes do not conform")
        ia = np.arange(npf)[None, :]
        mapP[(fa[:, None] * npf + ia).reshape(-1)] = (fb[:, None] * npf + ja).reshape(-1)
        mapP[(fb[:, None] * npf + ja).reshape(-1)] = np.broadcast_to(fa[:, None] * npf + ia, ja.shape).reshape(-1)
    return mapP.reshape(N_e, N_f)


def uniform_periodic_mesh(ra: ReferenceApproximation, limits, M, warp=None,
                          part: Optional[Tuple[int, int]] = None, structured: Optional[bool] = None) -> Mesh:
    """Periodic Cartesian-derived simplex mesh (1-D: intervals), optionally warped.

    ``part=(rank, world)`` builds only this rank's slab (cells split along the
    last axis), with ghost facet slots and send/recv lists for the halo.
    """
    d = ra.d
    geom = ra.geom
    if d == 1:
        M = (int(M),) if np.isscalar(M) else tuple(M)
        limits = (tuple(limits),) if np.isscalar(limits[0]) else tuple(limits)
        N_e = M[0]
        VX = [np.linspace(limits[0][0], limits[0][1], N_e + 1)]
        EtoV = np.stack([np.arange(N_e), np.arange(N_e) + 1], axis=1)
        xyz = _affine_nodes(VX, EtoV, geom)
        xyzf0 = [xyz[0] @ geom.Vf.T]
        mapP = _match_faces(xyzf0, 2, limits)
        xyz = apply_warp(xyz, warp)
        return Mesh(1, N_e, xyz, [xyz[0] @ geom.Vq.T], [xyz[0] @ geom.Vf.T], mapP, limits)

    M = tuple(int(m) for m in M)
    limits = tuple(tuple(l) for l in limits)
    ncell = int(np.prod(M))
    if ra.element in ("Quad", "Hex"):
        if part is not None and part != (0, 1):
            raise NotImplementedError("partitioned Quad / Hex meshes")
        return _box_mesh(ra, limits, M, warp)
    nsimp = 2 if d == 2 else 6
    if structured is None:
        structured = min(M) >= 4
    if part is None and structured:
        part = (0, 1)
    if part is None:
        cells = np.arange(ncell)
        EtoV, n1 = _cartesian_simplex_cells(d, M, cells)
        EtoV = _orient(d, EtoV, n1, limits)
        nv = int(np.prod(n1))
        VX = _vertex_coords(d, n1, limits, np.arange(nv))
        xyz = _affine_nodes(VX, EtoV, geom)
        xf0 = [x @ geom.Vf.T for x in xyz]
        mapP = _match_faces(xf0, ra.N_fac, limits)
        xyz = apply_warp(xyz, warp)
        return Mesh(d, EtoV.shape[0], xyz, [x @ geom.Vq.T for x in xyz],
                    [x @ geom.Vf.T for x in xyz], mapP, limits,
                    elem_gid=np.arange(EtoV.shape[0]))
    return _partitioned_mesh(ra, limits, M, warp, part)


def _box_mesh(ra: ReferenceApproximation, limits, M, warp) -> Mesh:
    """Periodic Cartesian mesh of Quad / Hex elements (one element per cell, first axis fastest), optionally warped."""
    d, geom = ra.d, ra.geom
    ncell = int(np.prod(M))
    c = np.arange(ncell)
    comps = [c % M[0], (c // M[0]) % M[1]] + ([c // (M[0] * M[1])] if d == 3 else [])
    xyz = []
    for m in range(d):
        lo, hi = limits[m]
        h = (hi - lo) / M[m]
        xyz.append(lo + h * (comps[m][:, None] + 0.5 * (1.0 + geom.rst[m][None, :])))
    mapP = _match_faces([x @ geom.Vf.T for x in xyz], ra.N_fac, limits)
    xyz = apply_warp(xyz, warp)
    return Mesh(d, ncell, xyz, [x @ geom.Vq.T for x in xyz], [x @ geom.Vf.T for x in xyz], mapP, limits,
                elem_gid=np.arange(ncell))


def _template_connectivity(ra: ReferenceApproximation):
    """Connectivity pattern of one cell of the Kuhn mesh, from a 3^d periodic mesh.

    Returns (off, t_nb, node_nb): for simplex t of a cell and facet node i, the
    neighbour is simplex t_nb[t, i] of cell (c + off[t, i]) at facet node node_nb[t, i].
    """
    d = ra.d
    lim = tuple((0.0, 3.0) for _ in range(d))
    m = uniform_periodic_mesh(ra, lim, (3,) * d, structured=False)
    nsimp = 2 if d == 2 else 6
    N_f = ra.N_f
    center = 1 + 3 * 1 + (9 if d == 3 else 0)
    e0 = center * nsimp
    mp = m.mapP[e0:e0 + nsimp]                                # (nsimp, N_f)
    ke, node = mp // N_f, mp % N_f
    cell, t_nb = ke // nsimp, ke % nsimp
    if d == 2:
        cc = np.stack([cell % 3, cell // 3], axis=2)
    else:
        cc = np.stack([cell % 3, (cell // 3) % 3, cell // 9], axis=2)
    off = cc - 1
    return off, t_nb, node


def structured_connectivity(ra: ReferenceApproximation, M, cells=None):
    """mapP rows (global linear ids) for the given cells (default all) of the periodic Kuhn
    mesh with M >= 3 cells per direction, O(N) time."""
    d = ra.d
    off, t_nb, node = _template_connectivity(ra)
    nsimp = 2 if d == 2 else 6
    c = np.arange(int(np.prod(M))) if cells is None else np.asarray(cells)
    comps = [c % M[0], (c // M[0]) % M[1]] + ([c // (M[0] * M[1])] if d == 3 else [])
    nb = 0
    stride = 1
    for m in range(d):
        nb = nb + ((comps[m][:, None, None] + off[None, :, :, m]) % M[m]) * stride
        stride *= M[m]
    mapP = (nb * nsimp + t_nb[None]) * ra.N_f + node[None]
    return mapP.reshape(c.size * nsimp, ra.N_f)


def _partitioned_mesh(ra, limits, M, warp, part) -> Mesh:
    """Slab partition along the last axis; local elements ordered interior first,
    halo-adjacent (boundary) last; ghost facet slots follow the local ones."""
    rank, world = part
    d = ra.d
    nsimp = 2 if d == 2 else 6
    N_f = ra.N_f
    Ml = M[-1]
    if Ml % world != 0 or min(M) < 3:
        raise ValueError("partitioned meshes need M[-1] divisible by world and M >= 3")
    per = Ml // world
    cells_per_layer = int(np.prod(M[:-1]))
    owner_of_cell = lambda c: (c // cells_per_layer) // per
    c0, c1 = rank * per * cells_per_layer, (rank + 1) * per * cells_per_layer
    gid = np.arange(c0 * nsimp, c1 * nsimp)
    mp = structured_connectivity(ra, M, np.arange(c0, c1))      # (nloc, N_f) global linear ids
    nb_el = mp // N_f
    nb_owner = owner_of_cell(nb_el // nsimp)
    is_bnd = np.any(nb_owner != rank, axis=1)
    order = np.concatenate([np.nonzero(~is_bnd)[0], np.nonzero(is_bnd)[0]])
    gid = gid[order]
    mp, nb_el, nb_owner = mp[order], nb_el[order], nb_owner[order]
    nloc = gid.size
    g2l = -np.ones(int(np.prod(M)) * nsimp, dtype=np.int64)
    g2l[gid] = np.arange(nloc)
    own = gid[:, None] * N_f + np.arange(N_f)[None, :]           # our global facet-node ids
    # ghost slots: for each neighbour rank (ascending), the remote facet nodes we read,
    # sorted by global linear id so both sides agree on the order
    mapP = np.where(nb_owner == rank, g2l[np.where(nb_owner == rank, nb_el, gid[0])] * N_f + mp % N_f, -1)
    nbr_ranks, send_idx, recv_off = [], [], []
    ghost = 0
    for r in sorted(set(np.unique(nb_owner).tolist()) - {rank}):
        sel = nb_owner == r
        need, first = np.unique(mp[sel], return_index=True)      # remote global facet ids
        mapP[sel] = nloc * N_f + ghost + np.searchsorted(need, mp[sel])
        # mapP is an involution: what r needs from us are the partners of what we need
        # from r, and r orders its ghost slots by ascending global id
        mine_sorted = np.sort(own[sel][first])
        send_idx.append(g2l[mine_sorted // N_f] * N_f + mine_sorted % N_f)
        nbr_ranks.append(int(r))
        recv_off.append(int(ghost))
        ghost += need.size
    # geometry of the local elements only
    cells = gid // nsimp
    t = gid % nsimp
    EtoV_all, n1 = _cartesian_simplex_cells(d, M, np.unique(cells))
    ucell = np.unique(cells)
    pos = np.searchsorted(ucell, cells)
    EtoV = EtoV_all.reshape(ucell.size, nsimp, d + 1)[pos, t]
    EtoV = _orient(d, EtoV, n1, limits)
    VXfun = lambda ids: _vertex_coords(d, n1, limits, ids)
    geom = ra.geom
    # affine nodes without materialising the global vertex list
    if d == 2:
        r, s = geom.rst
        lam = [-0.5 * (r + s), 0.5 * (1 + r), 0.5 * (1 + s)]
    else:
        r, s, tt = geom.rst
        lam = [-0.5 * (1 + r + s + tt), 0.5 * (1 + r), 0.5 * (1 + s), 0.5 * (1 + tt)]
    vc = [VXfun(EtoV[:, v]) for v in range(d + 1)]
    xyz = [sum(vc[v][m][:, None] * lam[v][None, :] for v in range(d + 1)) for m in range(d)]
    xyz = apply_warp(xyz, warp)
    return Mesh(d, nloc, xyz, [x @ geom.Vq.T for x in xyz], [x @ geom.Vf.T for x in xyz],
                mapP, limits, n_ghost=ghost, send_idx=send_idx, recv_off=recv_off,
                nbr_ranks=nbr_ranks, n_boundary=int(is_bnd.sum()), elem_gid=gid)


# --------------------------------------------------------------------------
# geometric factors
# --------------------------------------------------------------------------
@dataclass
class GeometricFactors:
    """Element-first C-ordered images of the reference's arrays
    (SpatialDiscretizations.jl:283-289):
    J_q (N_e,N_q); Lambda_q (N_e,d[n],d[m],N_q) = Λ_q[i,m,n,k];
    J_f (N_e,N_f); nJf (N_e,N_f,d) = nJf[m,i,k]; nJq (N_e,N_q,N_fac,d) = nJq[n,f,i,k]."""
    J_q: np.ndarray
    Lambda_q: np.ndarray
    J_f: np.ndarray
    nJf: np.ndarray
    nJq: np.ndarray


def _metrics_pointwise(dxdr):
    """metrics() of mesh.jl:181-227 on arrays dxdr[..., m, n] = dx_m/dr_n.  Returns J, Λ[..., l, m]."""
    d = dxdr.shape[-1]
    if d == 1:
        return dxdr[..., 0, 0], np.ones_like(dxdr)
    if d == 2:
        J = dxdr[..., 0, 0] * dxdr[..., 1, 1] - dxdr[..., 0, 1] * dxdr[..., 1, 0]
        Lam = np.empty_like(dxdr)
        Lam[..., 0, 0] = dxdr[..., 1, 1]
        Lam[..., 0, 1] = -dxdr[..., 0, 1]
        Lam[..., 1, 0] = -dxdr[..., 1, 0]
        Lam[..., 1, 1] = dxdr[..., 0, 0]
        return J, Lam
    J = np.linalg.det(dxdr)
    return J, J[..., None, None] * np.linalg.inv(dxdr)


class _Acc:
    """Accumulates nodal metric components straight into the ABI layouts (no big temporaries)."""

    def __init__(self, ra, ne, need_nJq):
        d = ra.d
        self.ra, self.d = ra, d
        self.Lambda_q = np.empty((ne, d, d, ra.N_q))          # [k, n(phys), m(ref), i]
        self.nJf = np.zeros((ne, ra.N_f, d))                  # [k, i, m(phys)]
        self.nref = np.stack(ra.nrstJ, axis=1)                # (N_f, l)
        self.need_nJq = need_nJq

    def add(self, l, m, aq, af):
        """metric J d xi_l / d x_m at volume nodes (aq: ne x N_q) and facet nodes (af: ne x N_f)."""
        self.Lambda_q[:, m, l, :] = aq
        self.nJf[:, :, m] += af * self.nref[None, :, l]

    def finish(self, J_q):
        ra, d = self.ra, self.d
        J_f = np.sqrt(np.sum(self.nJf ** 2, axis=2))
        nJq = None
        if self.need_nJq:
            npf = ra.nodes_per_face
            nrf = np.array([[ra.nrstJ[l][npf * f] for l in range(d)] for f in range(ra.N_fac)])
            nJq = np.ascontiguousarray(np.einsum("knli,fl->kifn", self.Lambda_q, nrf))
        return GeometricFactors(np.ascontiguousarray(J_q), self.Lambda_q, J_f, self.nJf, nJq)


def geometric_factors(mesh: Mesh, ra: ReferenceApproximation, metric_type: str = "exact",
                      chunk: int = 32768, need_nJq: bool = True, device: Optional[int] = None) -> GeometricFactors:
    """metric_type: 'exact' (mesh.jl:229-282) or 'curl' (ConservativeCurl/ChanWilcox).  With `device` the metrics are
    computed by the CUDA library (sse_geometric_factors) instead of NumPy."""
    d, g = ra.d, ra.geom
    if device is not None:
        return _gf_device(mesh, ra, metric_type, need_nJq, device)
    outs = []
    for s in range(0, mesh.N_e, chunk):
        xyz = [x[s:s + chunk] for x in mesh.xyz]
        if metric_type == "exact" or d == 1:
            outs.append(_gf_exact(ra, xyz, need_nJq))
        elif d == 2:
            outs.append(_gf_curl_2d(ra, xyz, need_nJq))
        else:
            outs.append(_gf_curl_3d(ra, xyz, need_nJq))
    if len(outs) == 1:
        return outs[0]
    return GeometricFactors(*[None if getattr(outs[0], f) is None else
                              np.concatenate([getattr(o, f) for o in outs], axis=0)
                              for f in ("J_q", "Lambda_q", "J_f", "nJf", "nJq")])


def _curl_lift(ra):
    """Operators of the degree N+1 element on which the curl argument of the tetrahedral metrics lives
    (mesh.jl:417-433): (g1, N -> N+1 interpolation, (Vq N+1->N)^T, (Vf N+1->N)^T)."""
    g = ra.geom
    key = (id(ra), g.N)
    if key not in _CURL_CACHE:
        g1 = geometry_element(3, g.N + 1, g.rst, g.rst)        # Vq of g1 = interp (N+1 nodes -> N nodes)
        up = g.interp(g1.rst)                                  # N nodes -> N+1 nodes
        down = g1.Vq                                           # N+1 nodes -> N nodes
        _CURL_CACHE[key] = (g1, up, np.ascontiguousarray((g.Vq @ down).T), np.ascontiguousarray((g.Vf @ down).T))
    return _CURL_CACHE[key]


def _gf_device(mesh, ra, metric_type, need_nJq, device):
    """GeometricFactors through sse_geometric_factors (csrc/kernels_geometry.cuh); arrays come back in the ABI layouts."""
    import ctypes as C
    from . import _abi, _lib
    d, g = ra.d, ra.geom
    ne = mesh.N_e
    col = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)       # column-major image
    cfg = _abi.sse_geom_config()
    cfg.d, cfg.N_map, cfg.N_q, cfg.N_f, cfg.N_e = d, g.rst[0].size, ra.N_q, ra.N_f, ne
    cfg.metric = _abi.SSE_METRIC_CURL if (metric_type == "curl" and d > 1) else _abi.SSE_METRIC_EXACT
    cfg.N1 = cfg.N_map
    keep = {"Drst": [col(D) for D in g.Drst], "Vq": col(g.Vq), "Vf": col(g.Vf), "nrstJ": col(np.stack(ra.nrstJ, axis=1))}
    if cfg.metric == _abi.SSE_METRIC_CURL and d == 3:
        if ra.element == "Hex":
            keep.update(D1=keep["Drst"], Vq1=keep["Vq"], Vf1=keep["Vf"])
        else:
            g1, up, VqT, VfT = _curl_lift(ra)
            cfg.N1 = g1.rst[0].size
            keep.update(D1=[col(D) for D in g1.Drst], Vq1=col(VqT.T), Vf1=col(VfT.T), up=col(up))
    ops = _abi.sse_geom_ops()
    pd = C.POINTER(C.c_double)
    ptr = lambda a: a.ctypes.data_as(pd) if a is not None else C.cast(None, pd)
    ops.Drst = (pd * 3)(*[ptr(keep["Drst"][m]) if m < d else ptr(None) for m in range(3)])
    ops.D1 = (pd * 3)(*[ptr(keep["D1"][m]) if "D1" in keep else ptr(None) for m in range(3)])
    ops.Vq, ops.Vf, ops.nrstJ = ptr(keep["Vq"]), ptr(keep["Vf"]), ptr(keep["nrstJ"])
    ops.up, ops.Vq1, ops.Vf1 = ptr(keep.get("up")), ptr(keep.get("Vq1")), ptr(keep.get("Vf1"))
    xs = [np.ascontiguousarray(x, dtype=np.float64) for x in mesh.xyz]
    xyz = (pd * 3)(*[ptr(xs[m]) if m < d else ptr(None) for m in range(3)])
    J_q = np.empty((ne, ra.N_q))
    Lam = np.empty((ne, d, d, ra.N_q))
    J_f = np.empty((ne, ra.N_f))
    nJf = np.empty((ne, ra.N_f, d))
    _lib.check(_lib.load().sse_geometric_factors(C.byref(cfg), C.byref(ops), int(device), xyz, ptr(J_q), ptr(Lam), ptr(J_f), ptr(nJf)))
    nJq = None
    if need_nJq:
        npf = ra.nodes_per_face
        nrf = np.array([[ra.nrstJ[l][npf * f] for l in range(d)] for f in range(ra.N_fac)])
        nJq = np.ascontiguousarray(np.einsum("knli,fl->kifn", Lam, nrf))
    return GeometricFactors(J_q, Lam, J_f, nJf, nJq)


def _gf_exact(ra, xyz, need_nJq=True):
    d, g = ra.d, ra.geom
    ne = xyz[0].shape[0]
    dq = np.empty((ne, ra.N_q, d, d))
    df = np.empty((ne, ra.N_f, d, d))
    for m in range(d):
        for n in range(d):
            dx = xyz[m] @ g.Drst[n].T
            dq[:, :, m, n] = dx @ g.Vq.T
            df[:, :, m, n] = dx @ g.Vf.T
    J_q, Lq = _metrics_pointwise(dq)
    _, Lf = _metrics_pointwise(df)
    acc = _Acc(ra, ne, need_nJq)
    for l in range(d):
        for m in range(d):
            acc.add(l, m, Lq[:, :, l, m], Lf[:, :, l, m])
    return acc.finish(J_q)


def _gf_curl_2d(ra, xyz, need_nJq=True):
    # StartUpDG.geometric_factors(x, y, Dr, Ds) interpolated (mesh.jl:284-339)
    g = ra.geom
    x, y = xyz
    Dr, Ds = g.Drst
    xr, xs, yr, ys = x @ Dr.T, x @ Ds.T, y @ Dr.T, y @ Ds.T
    J = -xs * yr + xr * ys
    L = {(0, 0): ys, (1, 0): -yr, (0, 1): -xs, (1, 1): xr}      # (l, m): J d xi_l / d x_m
    acc = _Acc(ra, x.shape[0], need_nJq)
    for (l, m), a in L.items():
        acc.add(l, m, a @ g.Vq.T, a @ g.Vf.T)
    return acc.finish(J @ g.Vq.T)


_CURL_CACHE = {}


def _gf_curl_hex(ra, xyz, need_nJq=True):
    """Conservative-curl metrics on hexahedra (mesh.jl:341-408): StartUpDG's geometric_factors(x, y, z, Dr, Ds, Dt) at
    the degree-N tensor Lobatto mapping nodes (products collocated there), then interpolated by Vq / Vf."""
    g = ra.geom
    x, y, z = xyz
    Dr, Ds, Dt = (np.ascontiguousarray(a.T) for a in g.Drst)
    xr, xs, xt = x @ Dr, x @ Ds, x @ Dt
    yr, ys, yt = y @ Dr, y @ Ds, y @ Dt
    zr, zs, zt = z @ Dr, z @ Ds, z @ Dt
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)

    def curl(a, b):
        Fr, Fs, Ft = (a @ Dr) * b, (a @ Ds) * b, (a @ Dt) * b
        return Fs @ Dt - Ft @ Ds, Ft @ Dr - Fr @ Dt, Fr @ Ds - Fs @ Dr

    acc = _Acc(ra, x.shape[0], need_nJq)
    VqT, VfT = np.ascontiguousarray(g.Vq.T), np.ascontiguousarray(g.Vf.T)
    for m, (a, b, sgn) in enumerate(((y, z, 1.0), (x, z, -1.0), (y, x, -1.0))):
        for l, comp in enumerate(curl(a, b)):
            comp = sgn * comp
            acc.add(l, m, comp @ VqT, comp @ VfT)
    return acc.finish(J @ VqT)


def _gf_curl_3d(ra, xyz, need_nJq=True):
    """Conservative-curl metrics on tets (mesh.jl:410-506; Chan & Wilcox 2019): the
    curl argument is a degree N+1 polynomial, the metric itself degree N."""
    if ra.element == "Hex":
        return _gf_curl_hex(ra, xyz, need_nJq)
    g = ra.geom
    g1, up, VqT, VfT = _curl_lift(ra)
    x, y, z = xyz
    Dr, Ds, Dt = g.Drst
    xr, xs, xt = x @ Dr.T, x @ Ds.T, x @ Dt.T
    yr, ys, yt = y @ Dr.T, y @ Ds.T, y @ Dt.T
    zr, zs, zt = z @ Dr.T, z @ Ds.T, z @ Dt.T
    J = xr * (ys * zt - zs * yt) - yr * (xs * zt - zs * xt) + zr * (xs * yt - ys * xt)
    X, Y, Z = x @ up.T, y @ up.T, z @ up.T
    D1r, D1s, D1t = (np.ascontiguousarray(a.T) for a in g1.Drst)

    def curl(a, b):
        # components (r, s, t) of  curl_xi( b * grad_xi a )
        Fr, Fs, Ft = (a @ D1r) * b, (a @ D1s) * b, (a @ D1t) * b
        return Fs @ D1t - Ft @ D1s, Ft @ D1r - Fr @ D1t, Fr @ D1s - Fs @ D1r

    acc = _Acc(ra, x.shape[0], need_nJq)
    for m, (a, b, sgn) in enumerate(((Y, Z, 1.0), (X, Z, -1.0), (Y, X, -1.0))):
        for l, comp in enumerate(curl(a, b)):
            comp = sgn * comp
            acc.add(l, m, comp @ VqT, comp @ VfT)
    return acc.finish(J @ g.Vq.T)


def project_jacobian(J_q: np.ndarray, ra: ReferenceApproximation) -> np.ndarray:
    """project_jacobian! (SpatialDiscretizations.jl:310-317)."""
    V, W = ra.V, ra.W
    proj = V @ np.linalg.solve(V.T @ (W[:, None] * V), V.T * W[None, :])
    return J_q @ proj.T


# --------------------------------------------------------------------------
# self checks (SpatialDiscretizations.jl:424-470)
# --------------------------------------------------------------------------
def check_normals(mesh: Mesh, gf: GeometricFactors) -> float:
    if mesh.n_ghost:
        raise ValueError("check_normals needs an unpartitioned mesh")
    flat = gf.nJf.reshape(-1, gf.nJf.shape[2])
    return float(np.max(np.abs(gf.nJf + flat[mesh.mapP])))


def check_facet_nodes(mesh: Mesh) -> float:
    if mesh.n_ghost:
        raise ValueError("check_facet_nodes needs an unpartitioned mesh")
    err = 0.0
    for m in range(mesh.d):
        lo, hi = mesh.limits[m]
        dlt = mesh.xyzf[m] - mesh.xyzf[m].reshape(-1)[mesh.mapP]
        dlt = dlt - (hi - lo) * np.round(dlt / (hi - lo))
        err = max(err, float(np.max(np.abs(dlt))))
    return err


def check_sbp_property_physical(ra: ReferenceApproximation, gf: GeometricFactors, k: int = 0):
    d = ra.d
    W, B, R, D = ra.W, ra.B, ra.R, ra.D_xi()
    out = []
    for n in range(d):
        Q = sum(0.5 * D[m].T * (W * gf.Lambda_q[k, n, m])[None, :]
                - 0.5 * (gf.Lambda_q[k, n, m] * W)[:, None] * D[m] for m in range(d)) \
            + 0.5 * R.T @ ((B * gf.nJf[k, :, n])[:, None] * R)
        E = R.T @ ((B * gf.nJf[k, :, n])[:, None] * R)
        out.append(float(np.max(np.abs(Q + Q.T - E))))
    return out
