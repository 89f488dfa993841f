"""Host-side assembly of the reference `Solver` image (forms, discretisation, operators).

Mirrors the constructors of src/Solvers/Solvers.jl:287-376 and src/Solvers/operators.jl:
what they store in `ReferenceOperators` / `FluxDifferencingOperators` /
`PhysicalOperators` is gathered here into the flat column-major buffers of the C ABI
(include/sse_b200.h :: sse_arrays).  One-time setup; nothing here is on the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _abi
from .laws import (BR1, CentralNumericalFlux, ConservativeFlux, EntropyConservativeFlux,
                   EntropyConservativeNumericalFlux, LaxFriedrichsNumericalFlux)
from .mesh import GeometricFactors, Mesh, geometric_factors, project_jacobian
from .reference import ReferenceApproximation


# ---- residual forms / strategies (Solvers.jl:75-115) ---------------------------------------
@dataclass(frozen=True)
class StandardForm:
    mapping_form: str = "skew"            # SkewSymmetricMapping() | "standard"
    inviscid_numerical_flux: object = field(default_factory=LaxFriedrichsNumericalFlux)
    viscous_numerical_flux: object = field(default_factory=BR1)


@dataclass(frozen=True)
class FluxDifferencingForm:
    mapping_form: str = "skew"
    inviscid_numerical_flux: object = field(default_factory=LaxFriedrichsNumericalFlux)
    viscous_numerical_flux: object = field(default_factory=BR1)
    two_point_flux: object = field(default_factory=EntropyConservativeFlux)


REFERENCE_OPERATOR, PHYSICAL_OPERATOR = "ReferenceOperator", "PhysicalOperator"


@dataclass
class SpatialDiscretization:
    """SpatialDiscretization(mesh, reference_approximation, metric_type; project_jacobian)
    (SpatialDiscretizations.jl:333-392)."""
    mesh: Mesh
    reference_approximation: ReferenceApproximation
    geometric_factors: GeometricFactors
    N_e: int

    @staticmethod
    def build(mesh: Mesh, ra: ReferenceApproximation, metric_type: str = "exact",
              project_jacobian_flag: bool = True, need_nJq: bool = True, device=None) -> "SpatialDiscretization":
        gf = geometric_factors(mesh, ra, metric_type, need_nJq=need_nJq, device=device)
        if metric_type == "exact" and project_jacobian_flag:
            gf.J_q = np.ascontiguousarray(project_jacobian(gf.J_q, ra))
        return SpatialDiscretization(mesh, ra, gf, mesh.N_e)


def apply_reference_mapping(gf: GeometricFactors, ra: ReferenceApproximation) -> np.ndarray:
    """Composite collapsed metric Λ_η (SpatialDiscretizations.jl:398-411), C-layout (N_e,n,m,i)."""
    if ra.J_ref is None:
        return gf.Lambda_q
    coef = ra.L_ref / ra.J_ref[:, None, None]                 # (i, m, l)
    return np.ascontiguousarray(np.einsum("iml,knli->knmi", coef, gf.Lambda_q))


def default_mass_solver(ra: ReferenceApproximation) -> int:
    """default_mass_matrix_solver (mass_matrix.jl:19-24, 41-75): identity V -> DiagonalSolver."""
    return _abi.SSE_MASS_DIAGONAL if ra.V_is_identity else _abi.SSE_MASS_WEIGHT_ADJUSTED


def mass_matrix_inverse(ra, gf, mass_solver: int) -> np.ndarray:
    """mass_matrix_inverse (mass_matrix.jl:155-167) for every element: (N_e, N_p, N_p)."""
    if mass_solver == _abi.SSE_MASS_DIAGONAL:
        inv = 1.0 / (ra.W[None, :] * gf.J_q)
        return np.einsum("ki,ij->kij", inv, np.eye(ra.N_p))
    V = ra.V
    if mass_solver == _abi.SSE_MASS_CHOLESKY:                  # inv(V' WJ V)   mass_matrix.jl:155-158
        return np.linalg.inv(np.einsum("qa,kq,qb->kab", V, ra.W[None, :] * gf.J_q, V))
    return np.einsum("qa,kq,qb->kab", V, ra.W[None, :] / gf.J_q, V)


def physical_operators(sd: SpatialDiscretization, form: StandardForm, mass_solver: int):
    """PhysicalOperators (operators.jl:83-160).  Returns VOL (N_e,d,N_q,N_p) and FAC (N_e,N_f,N_p)
    C-ordered == column-major (N_p,N_q,d,N_e) / (N_p,N_f,N_e)."""
    ra, gf = sd.reference_approximation, sd.geometric_factors
    d, V, R, W, B, D = ra.d, ra.V, ra.R, ra.W, ra.B, ra.D
    Minv = mass_matrix_inverse(ra, gf, mass_solver)
    Ne = sd.N_e
    VOL = np.empty((Ne, d, ra.N_q, ra.N_p))
    if d == 1 and form.mapping_form == "standard":
        core = V.T @ D[0].T * W[None, :]                                  # V' D' W
        VOL[:, 0] = np.transpose(Minv @ core[None], (0, 2, 1))
        FAC = -(Minv @ (V.T @ R.T * B[None, :])[None])
        return VOL, np.ascontiguousarray(np.transpose(FAC, (0, 2, 1)))
    Lam = apply_reference_mapping(gf, ra)                                 # (k, n, m, i)
    for n in range(d):
        if form.mapping_form == "standard":
            inner = sum(np.einsum("ji,ki->kji", D[m].T, W[None, :] * Lam[:, n, m]) for m in range(d))
        else:
            inner = sum(np.einsum("ji,ki->kji", D[m].T, 0.5 * W[None, :] * Lam[:, n, m])
                        - np.einsum("kj,ji->kji", 0.5 * W[None, :] * Lam[:, n, m], D[m])
                        for m in range(d))
            inner = inner + np.einsum("fj,kf,fi->kji", R, 0.5 * B[None, :] * gf.nJf[:, :, n], R)
        VOL[:, n] = np.transpose(Minv @ (V.T[None] @ inner), (0, 2, 1))
    FAC = -(Minv @ np.einsum("aq,fq,kf->kaf", V.T, R, B[None, :] * gf.J_f))
    return VOL, np.ascontiguousarray(np.transpose(FAC, (0, 2, 1)))


def _F(a):
    """C-ordered numpy array -> flat float64 buffer (memory image unchanged)."""
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def _colmajor(a):
    """Julia-shaped (row, col, ...) numpy array -> flat column-major buffer."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).reshape(-1)


@dataclass
class SolverImage:
    """Everything `sse_create` (or the oracle) needs: config + host buffers."""
    cfg: _abi.sse_config
    arrays: dict
    law: object
    form: object
    sd: SpatialDiscretization

    def c_arrays(self) -> _abi.sse_arrays:
        return _abi.fill_arrays(self.arrays)

    @property
    def state_shape(self):
        return (int(self.cfg.N_e), int(self.cfg.N_c), int(self.cfg.N_p))


def assemble(law, sd: SpatialDiscretization, form, strategy: str = REFERENCE_OPERATOR,
             mass_solver: Optional[int] = None, pass_nJq: bool = False) -> SolverImage:
    """Solver(conservation_law, spatial_discretization, form, strategy, alg, mass_solver,
    parallelism) (Solvers.jl:287-376) -> ABI image."""
    ra, gf, mesh = sd.reference_approximation, sd.geometric_factors, sd.mesh
    d = ra.d
    if law.d != d:
        raise ValueError("dimension mismatch between conservation law and discretization")
    if mass_solver is None:
        mass_solver = default_mass_solver(ra)
    if mass_solver == _abi.SSE_MASS_CHOLESKY and ra.V_is_identity:      # CholeskySolver(J_q, V::UniformScalingMap, W)
        mass_solver = _abi.SSE_MASS_DIAGONAL                             # is the DiagonalSolver (mass_matrix.jl:26-28)
    cfg = _abi.sse_config()
    cfg.abi_version = _abi.SSE_ABI_VERSION
    cfg.d, cfg.N_c, cfg.N_p, cfg.N_q, cfg.N_f, cfg.N_fac = d, law.N_c, ra.N_p, ra.N_q, ra.N_f, ra.N_fac
    cfg.p = ra.p
    cfg.N_e, cfg.N_ghost = sd.N_e, mesh.n_ghost
    cfg.pde = law.pde_id
    cfg.inviscid_flux = form.inviscid_numerical_flux.flux_id
    cfg.half_lambda = form.inviscid_numerical_flux.half_lambda
    cfg.viscous_flux = _abi.SSE_VISCOUS_BR1 if law.second_order else _abi.SSE_VISCOUS_NONE
    cfg.two_point_flux = (form.two_point_flux.two_point_id if isinstance(form, FluxDifferencingForm)
                          else _abi.SSE_TWO_POINT_CONSERVATIVE)
    cfg.mass_solver = mass_solver
    for m in range(d):
        cfg.a[m] = getattr(law, "a", (0.0,) * 3)[m]
    cfg.b = getattr(law, "b", 0.0)
    cfg.gamma = getattr(law, "gamma", 1.4)
    arrays = {}
    # V
    if ra.V_is_identity:
        cfg.v_kind = _abi.SSE_V_IDENTITY
    elif ra.V_warped is not None:
        cfg.v_kind = _abi.SSE_V_WARPED
        w = ra.V_warped
        arrays["A"], arrays["B"] = _colmajor(w.A), _colmajor(w.B)
        if d == 3:
            arrays["C"] = _colmajor(w.C)
        arrays["sigma_i"] = np.ascontiguousarray((w.sigma_i + 1).T.reshape(-1), dtype=np.int64)
        arrays["sigma_o"] = np.ascontiguousarray((w.sigma_o + 1).T.reshape(-1), dtype=np.int64)
        for m in range(d):
            cfg.M1d[m] = w.sigma_o.shape[m]
        arrays["V"] = _colmajor(ra.V)      # dense image as well (functionals / checks)
    else:
        cfg.v_kind = _abi.SSE_V_DENSE
        arrays["V"] = _colmajor(ra.V)
    arrays["R"] = _colmajor(ra.R)
    arrays["W"], arrays["Bf"] = _F(ra.W), _F(ra.B)
    arrays["J_q"], arrays["J_f"], arrays["nJf"] = _F(gf.J_q), _F(gf.J_f), _F(gf.nJf)
    arrays["mapP"] = np.ascontiguousarray(mesh.mapP.reshape(-1) + 1, dtype=np.int64)
    npf = ra.nodes_per_face
    arrays["nref"] = _F(np.array([[ra.nrstJ[m][npf * f] for m in range(d)] for f in range(ra.N_fac)]))

    if law.second_order or strategy == PHYSICAL_OPERATOR:
        if isinstance(form, FluxDifferencingForm):
            raise NotImplementedError("no physical-operator flux-differencing form (Solvers.jl:161-163)")
        cfg.form = _abi.SSE_FORM_STANDARD_PHYSICAL
        VOL, FAC = physical_operators(sd, form, mass_solver)
        arrays["VOL"], arrays["FAC"] = _F(VOL), _F(FAC)
        arrays["Lambda_q"] = _F(gf.Lambda_q)
    elif isinstance(form, FluxDifferencingForm):
        cfg.form = _abi.SSE_FORM_FLUX_DIFFERENCING
        S, Cfd = ra.flux_differencing_operators()
        arrays["S"] = [_colmajor(s) for s in S]
        if Cfd is not None:
            arrays["Cfd"] = _colmajor(Cfd)
        arrays["Lambda_q"] = _F(gf.Lambda_q)
        if pass_nJq:
            if gf.nJq is None:
                raise ValueError("geometric factors were built without nJq")
            arrays["nJq"] = _F(gf.nJq)
    else:
        cfg.form = _abi.SSE_FORM_STANDARD_REFERENCE
        arrays["D"] = [_colmajor(Dm) for Dm in ra.D]
        arrays["Lambda_q"] = _F(apply_reference_mapping(gf, ra))
    return SolverImage(cfg, arrays, law, form, sd)
