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
    lambda: cases.advection_diffusion_2d(M=3)])
def test_invariants(case):
    c = case()
    img, u = c.image(), c.u0(seed=0)
    du = oracle.rhs(img, u)
    scale = max(1.0, np.abs(du).max())
    assert np.all(np.isfinite(du))
    assert np.abs(analysis.conservation_residual(img, du)).max() < 1e-12 * scale * 100
    flux = c.form.inviscid_numerical_flux
    if c.law.pde_id == 0 and flux.half_lambda == 0.0:          # energy conservation, central flux
        assert np.abs(analysis.energy_residual(img, u, du)).max() < 1e-12 * scale
    if c.law.pde_id == 2:
        ds = analysis.entropy_residual(img, u, du)
        if flux.flux_id == 2:
            assert abs(ds) < 1e-11 * scale                      # entropy conservation, EC interface flux
        else:
            assert ds < 1e-11 * scale                           # entropy dissipation with LF
    if c.law.pde_id == 1:
        assert analysis.energy_residual(img, u, du)[0] < 0      # BR1 dissipates


@pytest.mark.parametrize("case", [lambda: cases.advection_2d(M=3, flux="lf"), lambda: cases.advection_3d(M=2, flux="lf")])
def test_physical_operator_equals_reference_operator(case):
    """PhysicalOperators fold M^-1 into VOL/FAC (operators.jl:132-160): same residual."""
    c = case()
    u = c.u0(seed=1)
    a = oracle.rhs(c.image(), u)
    b = oracle.rhs(assemble(c.law, c.sd, c.form, PHYSICAL_OPERATOR), u)
    assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max()


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
