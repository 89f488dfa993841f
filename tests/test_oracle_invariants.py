"""The reference's invariant assertions (runtests.jl:35-142; Analysis/conservation.jl:145-189) on
curved periodic 2-D/3-D meshes, plus the cross-variant identities of SURVEY.md §8c(iv)."""
import numpy as np
import pytest

import oracle
from sse_b200 import analysis, cases
from sse_b200.assembly import PHYSICAL_OPERATOR, assemble


@pytest.mark.parametrize("case", [
    lambda: cases.advection_2d(M=2, flux="lf0"), lambda: cases.advection_2d(M=3, flux="lf"),
    lambda: cases.advection_3d(M=2, flux="central"), lambda: cases.advection_3d(M=2, flux="lf"),
    lambda: cases.euler_vortex_2d(M=4, p=3, flux="ec"), lambda: cases.euler_vortex_2d(M=4, p=4, flux="lf"),
    lambda: cases.euler_vortex_2d(M=3, p=4, flux="ec", kind="nodal"),
    lambda: cases.euler_tgv_3d(M=2, flux="ec"), lambda: cases.euler_tgv_3d(M=2, flux="lf"),
    lambda: cases.euler_tgv_3d(M=2, p=3, flux="ec", kind="nodal"),
    lambda: cases.advection_diffusion_2d(M=3),
    # the degrees the compile-time kernels were extended to in round 2 (p = 7: examples/advection_3d.ipynb of the reference)
    lambda: cases.advection_3d(M=2, p=7, flux="central"), lambda: cases.euler_tgv_3d(M=2, p=6, flux="ec"),
    lambda: cases.euler_tgv_3d(M=2, p=7, flux="ec"),
    lambda: cases.euler_vortex_2d_standard(M=3, p=4, flux="lf"),
    lambda: cases.euler_vortex_2d_standard(M=3, p=3, flux="central", strategy=PHYSICAL_OPERATOR),
    lambda: cases.euler_tgv_3d_standard(M=2, p=3, flux="lf"),
    lambda: cases.euler_tgv_3d_standard(M=2, p=4, flux="central"),
    # multidimensional schemes with dense D, S, R (and dense V or V = I): multidimensional.jl:1-75, runtests.jl:98-109
    lambda: cases.advection_2d(M=3, flux="lf0", kind="modal_multi"),
    lambda: cases.advection_3d(M=2, p=3, flux="central", kind="nodal_multi"),
    lambda: cases.euler_vortex_2d(M=3, p=3, flux="ec", kind="modal_multi"),
    lambda: cases.euler_vortex_2d(M=3, p=3, flux="lf", kind="nodal_multi"),
    lambda: cases.euler_tgv_3d(M=2, p=2, flux="ec", kind="modal_multi")])
def test_invariants(case):
    c = case()
    img, u = c.image(), c.u0(seed=0)
    du = oracle.rhs(img, u)
    scale = max(1.0, np.abs(du).max())
    # round-off of the collapsed-coordinate modal basis grows about tenfold per degree (3-D Euler EC on 48 curved tets, this
    # restatement: conservation 1.2e-11, 2.6e-10, 1.8e-9, 1.5e-8 and entropy 6.9e-13, 1.2e-12, 7.8e-11, 6.2e-10 at p = 4, 5, 6, 7);
    # the reference's own assertions stop at p = 4, so the tolerances written for it are widened by that factor above it
    scale *= 10.0 ** max(0, int(img.cfg.p) - 4)
    assert np.all(np.isfinite(du))
    assert np.abs(analysis.conservation_residual(img, du)).max() < 1e-12 * scale * 100
    flux = c.form.inviscid_numerical_flux
    if c.law.pde_id == 0 and flux.half_lambda == 0.0:          # energy conservation, central flux
        assert np.abs(analysis.energy_residual(img, u, du)).max() < 1e-12 * scale
    if c.law.pde_id == 2 and int(img.cfg.form) == 2:            # entropy statements belong to the flux-differencing form
        ds = analysis.entropy_residual(img, u, du)
        if flux.flux_id == 2:
            assert abs(ds) < 1e-11 * scale                      # entropy conservation, EC interface flux
        else:
            assert ds < 1e-11 * scale                           # entropy dissipation with LF
    if c.law.pde_id == 1:
        assert analysis.energy_residual(img, u, du)[0] < 0      # BR1 dissipates


@pytest.mark.parametrize("case", [lambda: cases.advection_2d(M=3, flux="lf"), lambda: cases.advection_3d(M=2, flux="lf"),
                                  lambda: cases.euler_vortex_2d_standard(M=3, p=4, flux="lf"),
                                  lambda: cases.euler_tgv_3d_standard(M=2, p=3, flux="lf")])
def test_physical_operator_equals_reference_operator(case):
    """PhysicalOperators fold M^-1 into VOL/FAC (operators.jl:132-160): same residual."""
    c = case()
    u = c.u0(seed=1)
    a = oracle.rhs(c.image(), u)
    b = oracle.rhs(assemble(c.law, c.sd, c.form, PHYSICAL_OPERATOR), u)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


def test_multidimensional_operators():
    """ModalMulti / NodalMulti (multidimensional.jl:1-75): D_m = grad(VDM)_m P, R = V_f P with P the W-orthogonal projection.
    check_sbp_property (SpatialDiscretizations.jl:441-455), exact differentiation of P_p and D_m V = grad V on the nodes."""
    from sse_b200.reference import ModalMulti, ModalTensor, NodalMulti, reference_approximation
    for el, p in (("Tri", 4), ("Tet", 3)):
        t = reference_approximation(ModalTensor(p), el, mapping_degree=p)
        for T in (ModalMulti, NodalMulti):
            ra = reference_approximation(T(p), el, mapping_degree=p)
            assert ra.V_warped is None and ra.J_ref is None and not ra.is_tensor
            assert (ra.N_p == ra.N_q) == (T is NodalMulti)
            assert max(ra.check_sbp_property()) < 1e-13
            for m, Dm in enumerate(ra.D):
                assert np.abs(Dm @ t.V - t.D_xi()[m] @ t.V).max() < 1e-11          # both differentiate the basis exactly
                assert np.count_nonzero(np.abs(Dm) > 1e-14) > 0.9 * Dm.size         # dense, unlike the tensor-product operators
            S, C = ra.flux_differencing_operators()
            assert all(np.abs(Sm + Sm.T).max() < 1e-14 for Sm in S) and C is not None


def test_euler_standard_form_physical_flux():
    """StandardForm with the Euler physical flux (euler_navierstokes.jl:58-68, 85-91; standard_form_first_order.jl:16-63): on an
    affine mesh with a uniform state the weak-form residual vanishes (free-stream preservation: sum_m D_m applied to constant
    fluxes cancels against the facet terms), and the conservative two-point flux used at the interfaces
    (ConservationLaws.jl:83, euler_navierstokes.jl:152-158) is the arithmetic mean of the two physical fluxes."""
    from sse_b200.assembly import SpatialDiscretization, StandardForm, REFERENCE_OPERATOR
    from sse_b200.laws import EulerEquations
    from sse_b200.mesh import uniform_periodic_mesh
    from sse_b200.reference import ModalTensor, reference_approximation
    ra = reference_approximation(ModalTensor(3), "Tet", mapping_degree=1)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (2,) * 3, None)
    sd = SpatialDiscretization.build(mesh, ra, "exact", True)
    law = EulerEquations(3, 1.4)
    ic = lambda x: np.stack([c0 + 0 * x[0] for c0 in (1.1, 0.3, -0.2, 0.5, 2.7)], axis=-1)
    c = cases.Case("uniform", law, sd, StandardForm(inviscid_numerical_flux=cases._flux("lf")), REFERENCE_OPERATOR, ic)
    img = c.image()
    du = oracle.rhs(img, c.u0())
    assert np.abs(du).max() < 1e-11
    import ctypes as C
    L = oracle.lib()
    uL, uR = np.array([1.0, 0.2, -0.1, 0.3, 2.5]), np.array([0.8, -0.1, 0.25, 0.1, 2.1])
    F = np.zeros(15)
    pd = C.POINTER(C.c_double)
    L.sse_oracle_two_point_flux(C.byref(img.cfg), 0, uL.ctypes.data_as(pd), uR.ctypes.data_as(pd), F.ctypes.data_as(pd))
    def phys(u):
        rho, V, E = u[0], u[1:4] / u[0], u[4]
        p = 0.4 * (E - 0.5 * rho * V @ V)
        return np.array([[u[1 + n] for n in range(3)]] + [[u[1 + m] * V[n] + (p if m == n else 0.0) for n in range(3)] for m in range(3)]
                        + [[(E + p) * V[n] for n in range(3)]])
    ref = 0.5 * (phys(uL) + phys(uR))
    assert np.abs(F.reshape(3, 5).T - ref).max() < 1e-14


def test_recomputed_nJq_equals_stored():
    c = cases.euler_tgv_3d(M=2, flux="ec")
    u = c.u0(seed=2)
    a, b = oracle.rhs(c.image(), u), oracle.rhs(c.image(pass_nJq=True), u)
    assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()


def test_pointwise_physics():
    L = oracle.lib()
    # logmean: Taylor branch and log branch agree with the definition (ConservationLaws.jl:132-156)
    for x, y in [(1.0, 1.0 + 1e-9), (1.0, 1.01), (0.7, 1.3), (2.0, 0.5), (1e-3, 2e-3)]:
        ref = (y - x) / np.log(y / x)
        assert abs(L.sse_oracle_logmean(x, y) - ref) <= 2e-13 * ref
        assert abs(L.sse_oracle_inv_logmean(x, y) - 1 / ref) <= 2e-13 / ref
    assert L.sse_oracle_logmean(1.5, 1.5) == 1.5


def test_burgers_1d_fluxdiff_reference_testset():
    """test/burgers_fluxdiff_1d.jl, runtests.jl:82-87: inviscid Burgers with the EC two-point and interface flux on
    NodalTensor(7) Lobatto lines conserves the primary variable and the energy u^2/2 to round-off along the whole run
    (T = 0.3, CFL 0.1, CarpenterKennedy2N54)."""
    from sse_b200.solver import CK54_A, CK54_B
    c = cases.burgers_1d(M=20, p=7, flux="ec")
    img, u = c.image(), c.u0()
    ra = c.sd.reference_approximation
    dt = 0.1 * (2.0 / (ra.N_p * c.sd.N_e))
    tmp = np.zeros_like(u)
    for it in range(int(round(0.3 / dt))):
        for s in range(5):
            tmp = CK54_A[s] * tmp + dt * oracle.rhs(img, u)
            u = u + CK54_B[s] * tmp
        if it % 24 == 0:
            du = oracle.rhs(img, u)
            assert np.abs(analysis.conservation_residual(img, du)).max() < 1e-10      # runtests.jl:85
            assert abs(analysis.energy_residual(img, u, du)) < 1e-10                  # runtests.jl:86
    assert np.all(np.isfinite(u))


def test_oracle_is_independent_of_its_thread_count():
    """The OpenMP element loops of the oracle (Threads.@threads in the reference, Solvers.jl:505-511) only partition
    independent elements: one thread and several threads must return the same bits — the CPU baseline of bench.py and the
    checker of the parity tests are the same function."""
    for c in (cases.euler_tgv_3d(M=2, flux="lf"), cases.advection_diffusion_2d(M=3)):
        img, u = c.image(), c.u0(seed=4)
        a, b = oracle.rhs(img, u, nthreads=1), oracle.rhs(img, u, nthreads=4)
        assert np.array_equal(a, b)
