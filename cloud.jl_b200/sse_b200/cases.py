"""The BASELINE.json configurations as concrete synthetic inputs.

Each builder returns a `Case` (law, discretisation, form, strategy, initial-data callable);
`Case.image()` assembles the Solver image, `Case.u0(seed)` the L2-projected initial state
with an optional smooth-random modal perturbation (numpy default_rng(seed)).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .assembly import (FluxDifferencingForm, PHYSICAL_OPERATOR, REFERENCE_OPERATOR, SpatialDiscretization,
                       StandardForm, assemble)
from .laws import (CentralNumericalFlux, EntropyConservativeNumericalFlux, EulerEquations, euler_periodic_test,
                   InviscidBurgersEquation, ViscousBurgersEquation, initial_data_gassner,
                   LaxFriedrichsNumericalFlux, LinearAdvectionDiffusionEquation, LinearAdvectionEquation,
                   initial_data_cosine, initial_data_sine, isentropic_vortex, project_function,
                   taylor_green_vortex)
from .mesh import ChanWarping, DelReyWarping, uniform_periodic_mesh
from .reference import ModalMulti, ModalTensor, NodalMulti, NodalTensor, reference_approximation


@dataclass
class Case:
    name: str
    law: object
    sd: SpatialDiscretization
    form: object
    strategy: str
    ic: object

    def image(self, **kw):
        return assemble(self.law, self.sd, self.form, self.strategy, **kw)

    def u0(self, seed: Optional[int] = None, eps: float = 2e-3) -> np.ndarray:
        sd = self.sd
        u = project_function(self.ic, sd.reference_approximation, sd.geometric_factors.J_q, sd.mesh.xyzq)
        if seed is not None:
            rng = np.random.default_rng(seed)
            scale = np.abs(u).max(axis=(0, 2), keepdims=True)
            scale = np.where(scale > 0, scale, 1.0)
            u = u + eps * rng.standard_normal(u.shape) * scale
        return np.ascontiguousarray(u)

    @property
    def dof(self) -> int:
        ra = self.sd.reference_approximation
        return ra.N_p * self.law.N_c * self.sd.N_e


def _flux(name: str):
    return {"lf": LaxFriedrichsNumericalFlux(1.0), "lf0": LaxFriedrichsNumericalFlux(0.0),
            "central": CentralNumericalFlux(), "ec": EntropyConservativeNumericalFlux()}[name]


def _approx(kind: str, p: int):
    """"modal" / "nodal": the sum-factorised tensor-product schemes (tensor_simplex.jl); "modal_multi" / "nodal_multi": the
    multidimensional schemes with dense operators (multidimensional.jl:1-75)."""
    return {"modal": ModalTensor, "nodal": NodalTensor, "modal_multi": ModalMulti, "nodal_multi": NodalMulti}[kind](p)


def advection_2d(M=4, p=4, flux="lf", kind="modal", warp=0.1, part=None) -> Case:
    """config 1: examples/advection_2d.ipynb / runtests.jl:38-60."""
    ra = reference_approximation(_approx(kind, p), "Tri", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M), DelReyWarping(warp, (1.0, 1.0)), part)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("advection_2d", LinearAdvectionEquation((1.0, 1.0)), sd,
                StandardForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                initial_data_sine(1.0, (2 * np.pi, 2 * np.pi)))


def euler_vortex_2d(M=4, p=4, flux="lf", kind="modal", part=None) -> Case:
    """config 2: test/euler_vortex_2d_modal.jl."""
    g = 1.4
    ra = reference_approximation(_approx(kind, p), "Tri", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M), ChanWarping(1.0 / 16.0, (1.0, 1.0)), part)
    sd = SpatialDiscretization.build(mesh, ra, "curl")
    ic = isentropic_vortex(g, 0.4, 0.0, 0.1, np.sqrt(2 / (g - 1) * (1 - 0.75 ** (g - 1))), 1.0, (0.5, 0.5))
    return Case("euler_vortex_2d", EulerEquations(2, g), sd,
                FluxDifferencingForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR, ic)


def advection_diffusion_2d(M=4, p=4, kind="modal", part=None) -> Case:
    """config 3: 2-D advection-diffusion, BR1 (PhysicalOperators are the only 2nd-order strategy)."""
    ra = reference_approximation(_approx(kind, p), "Tri", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M), DelReyWarping(0.1, (1.0, 1.0)), part)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("advection_diffusion_2d", LinearAdvectionDiffusionEquation((1.0, 1.0), 5e-2), sd,
                StandardForm(inviscid_numerical_flux=_flux("lf")), PHYSICAL_OPERATOR,
                initial_data_sine(1.0, (2 * np.pi, 2 * np.pi)))


def advection_3d(M=2, p=4, flux="central", kind="modal", part=None) -> Case:
    """config 4: test/advection_3d.jl."""
    ra = reference_approximation(_approx(kind, p), "Tet", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (M,) * 3, DelReyWarping(0.1, (1.0,) * 3), part)
    sd = SpatialDiscretization.build(mesh, ra, "curl")
    return Case("advection_3d", LinearAdvectionEquation((1.0, 1.0, 1.0)), sd,
                StandardForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                initial_data_cosine(1.0, (2 * np.pi,) * 3))


def euler_tgv_3d(M=2, p=4, flux="lf", kind="modal", part=None) -> Case:
    """config 5 (headline): 3-D Euler Taylor-Green vortex on curved tets, flux differencing."""
    L = 2 * np.pi
    ra = reference_approximation(_approx(kind, p), "Tet", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, L),) * 3, (M,) * 3, ChanWarping(1.0 / 16.0, (L,) * 3), part)
    sd = SpatialDiscretization.build(mesh, ra, "curl")
    return Case("euler_tgv_3d", EulerEquations(3, 1.4), sd,
                FluxDifferencingForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                taylor_green_vortex(1.4, 0.1))


def euler_vortex_2d_standard(M=4, p=4, flux="lf", kind="modal", strategy=REFERENCE_OPERATOR, part=None) -> Case:
    """2-D Euler vortex in StandardForm (standard_form_first_order.jl:16-94 with the Euler physical flux,
    euler_navierstokes.jl:58-68, 85-91): the weak-form residual the reference runs when no entropy-stable form is asked for."""
    c = euler_vortex_2d(M, p, flux, kind, part)
    return Case("euler_vortex_2d_standard", c.law, c.sd, StandardForm(inviscid_numerical_flux=_flux(flux)), strategy, c.ic)


def euler_tgv_3d_standard(M=2, p=4, flux="lf", kind="modal", strategy=REFERENCE_OPERATOR, part=None) -> Case:
    """3-D Euler Taylor-Green vortex on curved tets in StandardForm (ReferenceOperator or PhysicalOperator strategy)."""
    c = euler_tgv_3d(M, p, flux, kind, part)
    return Case("euler_tgv_3d_standard", c.law, c.sd, StandardForm(inviscid_numerical_flux=_flux(flux)), strategy, c.ic)


def euler_periodic_3d_hex(M=2, p=4, flux="ec") -> Case:
    """test/euler_3d.jl (runtests.jl:131-144): 3-D Euler density wave on warped hexahedra, NodalTensor Lobatto
    collocation (diagonal-E), flux differencing, conservative-curl metrics."""
    L = 2.0
    ra = reference_approximation(NodalTensor(p), "Hex", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, L),) * 3, (M,) * 3, ChanWarping(1.0 / 16.0, (L,) * 3))
    sd = SpatialDiscretization.build(mesh, ra, "curl")
    return Case("euler_periodic_3d_hex", EulerEquations(3, 1.4), sd,
                FluxDifferencingForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                euler_periodic_test(3, 1.4, 0.2, L))


def burgers_1d(M=20, p=7, flux="ec") -> Case:
    """test/burgers_fluxdiff_1d.jl (runtests.jl:82-87): inviscid Burgers, flux differencing on Lobatto NodalTensor lines."""
    ra = reference_approximation(NodalTensor(p), "Line")
    mesh = uniform_periodic_mesh(ra, (0.0, 2.0), M)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("burgers_1d", InviscidBurgersEquation(), sd,
                FluxDifferencingForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                initial_data_gassner(np.pi, 0.01))


def viscous_burgers_1d(M=8, p=5, b=5e-2, flux="lf") -> Case:
    """ViscousBurgersEquation(b) (burgers.jl:23-49) with BR1 on Lobatto NodalTensor lines, PhysicalOperators."""
    ra = reference_approximation(NodalTensor(p), "Line")
    mesh = uniform_periodic_mesh(ra, (0.0, 2.0), M)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("viscous_burgers_1d", ViscousBurgersEquation((1.0,), b), sd,
                StandardForm(inviscid_numerical_flux=_flux(flux)), PHYSICAL_OPERATOR, initial_data_gassner(np.pi, 0.01))


def viscous_burgers_2d(M=3, p=4, b=5e-2, kind="modal") -> Case:
    """ViscousBurgersEquation((1, 1), b) with BR1 on curved triangles, PhysicalOperators."""
    ra = reference_approximation(_approx(kind, p), "Tri", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M), DelReyWarping(0.1, (1.0, 1.0)))
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("viscous_burgers_2d", ViscousBurgersEquation((1.0, 1.0), b), sd,
                StandardForm(inviscid_numerical_flux=_flux("lf")), PHYSICAL_OPERATOR,
                initial_data_sine(1.0, (2 * np.pi, 2 * np.pi)))


def advection_2d_quad(M=2, p=4, flux="lf", warp=0.1) -> Case:
    """runtests.jl:62-80: 2-D advection, flux-differencing form on warped quadrilaterals (NodalTensor Lobatto)."""
    ra = reference_approximation(NodalTensor(p), "Quad", mapping_degree=p)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 2, (M, M), DelReyWarping(warp, (1.0, 1.0)))
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    return Case("advection_2d_quad", LinearAdvectionEquation((1.0, 1.0)), sd,
                FluxDifferencingForm(inviscid_numerical_flux=_flux(flux)), REFERENCE_OPERATOR,
                initial_data_sine(1.0, (2 * np.pi, 2 * np.pi)))


BUILDERS = {"advection_2d": advection_2d, "euler_vortex_2d": euler_vortex_2d,
            "advection_diffusion_2d": advection_diffusion_2d, "advection_3d": advection_3d,
            "euler_tgv_3d": euler_tgv_3d, "euler_vortex_2d_standard": euler_vortex_2d_standard,
            "euler_tgv_3d_standard": euler_tgv_3d_standard,
            "euler_periodic_3d_hex": euler_periodic_3d_hex, "burgers_1d": burgers_1d, "advection_2d_quad": advection_2d_quad}
