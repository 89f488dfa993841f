"""Pins the oracle against the reference's own golden vectors (test/runtests.jl).

Reproduced to round-off: the 1-D testsets (runtests.jl:14-36, 89-96), the 2-D triangle testsets of the collapsed
ModalTensor operators (advection :38-60; Euler vortex with flux differencing, entropy projection, facet correction and
the weight-adjusted mass solve :111-121), the quadrilateral flux-differencing testset (:62-80) and the 3-D Euler
hexahedral testset (:131-144).  The triangle mesh split and mapping nodes of the un-vendored StartUpDG / NodesAndModes
were identified through these goldens (sse_b200/mesh.py, sse_b200/reference.py).  The tetrahedral advection golden
(:123-129) is reproduced to the accuracy its error quadrature allows (4e-5: the Jaskowiec-Sukumar rule is tabulated data of the
un-vendored StartUpDG); it identifies StartUpDG's cube-to-tet split.  Not reproducible here: the NodalMultiDiagE golden
(:98-109, tabulated SBP nodes).  Time integration is CarpenterKennedy2N54 with the reference's dt, except for the Hex testset,
whose DP8 tableau (OrdinaryDiffEq, un-vendored) is replaced by CK54 at a time step where the ODE is solved to 2e-11."""
import numpy as np

import oracle
from sse_b200 import analysis, cases
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


def _l2_error(c, u):
    """ErrorAnalysis with the scheme's own quadrature (Analysis/error.jl:14-24, 60-80)."""
    sd = c.sd
    ra = sd.reference_approximation
    uq = np.einsum("qa,kea->keq", ra.V, u)
    ue = np.transpose(c.ic(sd.mesh.xyzq), (0, 2, 1))
    return np.sqrt(np.einsum("kei,ki,kei->e", ue - uq, ra.W[None, :] * sd.geometric_factors.J_q, ue - uq))


ADVECTION_2D_TRI_GOLDEN = 0.2660013939427627                                   # runtests.jl:57
ADVECTION_2D_QUAD_GOLDEN = 0.04790536605026519                                 # runtests.jl:78
EULER_VORTEX_2D_MODAL_GOLDEN = [0.015568197027072704, 0.040539693811761104,
                                0.04060141777050208, 0.043960971468832745]     # runtests.jl:114-119
EULER_3D_HEX_GOLDEN = [0.18342164491797003, 0.1834216449179776, 0.18342164491796725,
                       0.18342164491796784, 0.2751324673769553]                # runtests.jl:134-140


def test_advection_2d_modal_tri_golden():
    """runtests.jl:38-60: StandardForm, skew-symmetric mapping, LF(0) = central flux, ModalTensor(4) on 2 x 2 x 2 warped
    triangles, one period with dt = 1/100."""
    c = cases.advection_2d(M=2, p=4, flux="lf0", warp=0.1)
    img = c.image()
    u = ck54(img, c.u0(), 1.0 / 100, 100)
    assert abs(_l2_error(c, u)[0] - ADVECTION_2D_TRI_GOLDEN) < TOL
    du = oracle.rhs(img, u)
    assert np.abs(analysis.conservation_residual(img, du)).max() < TOL        # runtests.jl:58
    assert abs(analysis.energy_residual(img, u, du)) < TOL                     # runtests.jl:59


def test_advection_2d_quad_fluxdiff_golden():
    """runtests.jl:62-80: FluxDifferencingForm() on NodalTensor(4) Lobatto quadrilaterals (diagonal-E, no correction)."""
    c = cases.advection_2d_quad(M=2, p=4, flux="lf", warp=0.1)
    img = c.image()
    u = ck54(img, c.u0(), 1.0 / 100, 100)
    assert abs(_l2_error(c, u)[0] - ADVECTION_2D_QUAD_GOLDEN) < TOL
    assert np.abs(analysis.conservation_residual(img, oracle.rhs(img, u))).max() < TOL   # runtests.jl:79


def test_euler_vortex_2d_modal_tri_golden():
    """test/euler_vortex_2d_modal.jl, runtests.jl:111-121: the 2-D instance of the headline path — ModalTensor(3)
    collapsed triangles, flux differencing with the Ranocha flux, entropy projection, facet correction, weight-adjusted
    mass solve, Lax-Friedrichs interface flux, ChanWilcox metrics on a ChanWarping(1/16) mesh; 1000 CK54 steps."""
    c = cases.euler_vortex_2d(M=4, p=3, flux="lf")
    img = c.image()
    T = 1.0 / 0.4
    u = ck54(img, c.u0(), T / 1000, 1000)
    assert np.allclose(_l2_error(c, u), EULER_VORTEX_2D_MODAL_GOLDEN, rtol=0, atol=TOL)
    assert np.abs(analysis.conservation_residual(img, oracle.rhs(img, u))).max() < TOL   # runtests.jl:120


def test_euler_3d_hex_golden():
    """test/euler_3d.jl, runtests.jl:131-144: 3-D Euler, EC two-point and interface flux, NodalTensor(4) Lobatto
    hexahedra, conservative-curl metrics on a ChanWarping(1/16) mesh, one period T = 2.  The reference integrates with
    DP8 (250 steps); CK54 with 2500 steps solves the same ODE to 2e-11 (the error falls 16x per halving of dt:
    1.4e-8, 5.5e-11, 3.4e-12 at 500, 2000, 4000 steps)."""
    c = cases.euler_periodic_3d_hex(M=2, p=4, flux="ec")
    img = c.image()
    u0 = c.u0()
    du0 = oracle.rhs(img, u0)
    assert np.abs(analysis.conservation_residual(img, du0)).max() < TOL       # runtests.jl:141
    assert abs(analysis.entropy_residual(img, u0, du0)) < TOL                 # runtests.jl:142
    u = ck54(img, u0, 2.0 / 2500, 2500)
    assert np.allclose(_l2_error(c, u), EULER_3D_HEX_GOLDEN, rtol=0, atol=TOL)


ADVECTION_3D_GOLDEN = 0.1876141674772107      # runtests.jl:126


def test_advection_3d_tet_golden_to_quadrature_accuracy():
    """test/advection_3d.jl, runtests.jl:123-129: ModalTensor(4) tets, M = 2, DelRey warping 0.1, conservative-curl metrics,
    skew-symmetric StandardForm with the central flux, CarpenterKennedy2N54 at the reference's dt up to T = 1.

    The reference measures the L2 error with JaskowiecSukumarQuadrature(2p + 3) after projecting node positions and J onto
    P_p (Analysis/error.jl:22-46); that rule is tabulated in StartUpDG and not available here, so the same functional is
    evaluated with collapsed Gauss rules instead: 0.187578 once converged (q >= 8), i.e. 3.6e-5 below the golden, with a
    spread of -1.8e-4 ... +1.6e-5 over rules of comparable degree (q = 5, 6, 7).  The golden therefore pins the path to
    ~2e-4 relative, which is decisive for the mesh: the three other body-diagonal splits of the cube give 0.1766, 0.1557
    and 0.1595 (sse_b200/mesh.py:_cartesian_simplex_cells).  Conservation and energy are pinned to round-off."""
    from sse_b200 import reference as R
    from sse_b200.laws import initial_data_cosine
    from sse_b200.quadrature import GaussQuadrature
    c = cases.advection_3d(M=2, p=4, flux="central")
    ra, sd = c.sd.reference_approximation, c.sd
    img = c.image()
    exact = initial_data_cosine(1.0, (2 * np.pi,) * 3)
    u = project_function(exact, ra, sd.geometric_factors.J_q, sd.mesh.xyzq)
    h = 1.0 / (ra.N_p * sd.N_e) ** (1 / 3)
    dt = 0.1 * h / np.sqrt(3.0)
    n = int(np.floor(1.0 / dt + 1e-12))
    du0 = oracle.rhs(img, u)
    assert abs(analysis.conservation_residual(img, du0)[0]) < TOL and abs(analysis.energy_residual(img, u, du0)[0]) < TOL
    u = ck54(img, u, dt, n)
    u = ck54(img, u, 1.0 - n * dt, 1)
    du1 = oracle.rhs(img, u)
    assert abs(analysis.conservation_residual(img, du1)[0]) < TOL and abs(analysis.energy_residual(img, u, du1)[0]) < TOL
    (re, se, te), we = R.simplex_quadrature("Tet", (LGQuadrature(8), LGQuadrature(8), GaussQuadrature(8, 1, 0)))
    Vm_err, Vm_q = R._poly_basis(3, 4, [re, se, te]), R._poly_basis(3, 4, ra.rstq)
    P = np.linalg.solve(Vm_q.T @ (ra.W[:, None] * Vm_q), Vm_q.T * ra.W[None, :])         # error.jl:33-38
    v2e = Vm_err @ P
    err = 0.0
    for k in range(sd.N_e):
        e = exact([v2e @ sd.mesh.xyzq[m][k] for m in range(3)])[..., 0] - v2e @ ra.V @ u[k, 0]
        err += np.dot(e, (we * (v2e @ sd.geometric_factors.J_q[k])) * e)
    assert abs(np.sqrt(err) - ADVECTION_3D_GOLDEN) < 1.0e-4
