"""Pins the oracle against the reference's own golden vectors (test/runtests.jl).

Only the 1-D testsets can be reproduced without the un-vendored StartUpDG mesh/node data; they
exercise the same residual code (flux_differencing_form.jl, standard_form_second_order.jl,
ConservationLaws) with d = 1.  Time integration is CarpenterKennedy2N54 with the reference's dt."""
import numpy as np

import oracle
from sse_b200 import analysis
from sse_b200.assembly import (FluxDifferencingForm, PHYSICAL_OPERATOR, SpatialDiscretization, StandardForm,
                               assemble)
from sse_b200.laws import (EntropyConservativeNumericalFlux, EulerEquations, LaxFriedrichsNumericalFlux,
                           LinearAdvectionDiffusionEquation, initial_data_sine, project_function)
from sse_b200.mesh import uniform_periodic_mesh
from sse_b200.quadrature import LGQuadrature
from sse_b200.reference import ModalMulti, NodalTensor, reference_approximation
from sse_b200.solver import CK54_A, CK54_B

TOL = 1.0e-10   # test/runtests.jl:12


def ck54(img, u, dt, n):
    tmp = np.zeros_like(u)
    for _ in range(n):
        for s in range(5):
            du = oracle.rhs(img, u)
            tmp = CK54_A[s] * tmp + dt * du
            u = u + CK54_B[s] * tmp
    return u


def euler_1d_setup():
    g = 1.4
    ra = reference_approximation(NodalTensor(5), "Line", volume_quadrature_rule=LGQuadrature(5))
    mesh = uniform_periodic_mesh(ra, (0.0, 2.0), 4)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    img = assemble(EulerEquations(1, g), sd, FluxDifferencingForm(inviscid_numerical_flux=EntropyConservativeNumericalFlux()))

    def exact(x):
        rho = 1.0 + 0.2 * np.sin(np.pi * x[0])
        return np.stack([rho, rho, 1.0 / (g - 1) + 0.5 * rho], axis=-1)
    return ra, mesh, sd, img, exact


EULER_1D_GOLDEN = [3.5808560177567635e-5, 5.2129828619609155e-5, 1.2637647535378534e-4]   # runtests.jl:92


def test_euler_1d_gauss_collocation_golden():
    """test/euler_1d_gauss.jl, runtests.jl:89-96."""
    ra, mesh, sd, img, exact = euler_1d_setup()
    u0 = project_function(exact, ra, sd.geometric_factors.J_q, mesh.xyzq)
    u = ck54(img, u0, 2.0 / 1000, 1000)
    ue = np.transpose(exact(mesh.xyzq), (0, 2, 1))
    l2 = np.sqrt(np.einsum("kei,ki,kei->e", ue - u, ra.W[None, :] * sd.geometric_factors.J_q, ue - u))
    assert np.allclose(l2, EULER_1D_GOLDEN, rtol=0, atol=TOL)
    du = oracle.rhs(img, u)
    assert np.abs(analysis.conservation_residual(img, du)).max() < TOL      # runtests.jl:94
    assert abs(analysis.entropy_residual(img, u, du)) < TOL                 # runtests.jl:95


def test_advection_diffusion_1d_br1_golden():
    """runtests.jl:14-36 (ModalMulti(4) Line, StandardMapping, LF + BR1, PhysicalOperator)."""
    ra = reference_approximation(ModalMulti(4), "Line")
    mesh = uniform_periodic_mesh(ra, (0.0, 1.0), 4)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    law = LinearAdvectionDiffusionEquation((1.0,), 5.0e-2)
    img = assemble(law, sd, StandardForm(mapping_form="standard", inviscid_numerical_flux=LaxFriedrichsNumericalFlux()),
                   PHYSICAL_OPERATOR)
    u0 = project_function(initial_data_sine(1.0, 2 * np.pi), ra, sd.geometric_factors.J_q, mesh.xyzq)
    u = ck54(img, u0, 1.0 / 100, 100)
    k = 2 * np.pi
    ue = (np.sin(k * (mesh.xyzq[0] - 1.0)) * np.exp(-5.0e-2 * k * k))[:, None, :]
    uq = np.einsum("qa,kea->keq", ra.V, u)
    l2 = np.sqrt(np.einsum("kei,ki,kei->e", ue - uq, ra.W[None, :] * sd.geometric_factors.J_q, ue - uq))
    assert abs(l2[0] - 6.988216111882884e-6) < TOL                          # runtests.jl:34
    assert np.abs(analysis.conservation_residual(img, oracle.rhs(img, u))).max() < TOL
