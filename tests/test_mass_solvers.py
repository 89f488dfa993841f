"""CholeskySolver (mass_matrix.jl:1-39, 117-122, 169-175): ldiv!(cholesky(Symmetric(V' WJ_k V)), rhs).

CPU: the oracle's factor-and-substitute restatement against an independent NumPy solve, the reference's invariant
assertions with the exact mass matrix, and the V = I collapse to the DiagonalSolver.  GPU (marked): parity of the
device factorisation + substitution with the oracle through the C ABI."""
import numpy as np
import pytest

import oracle
from sse_b200 import _abi, analysis, cases
from sse_b200.assembly import PHYSICAL_OPERATOR, assemble

CHOL = _abi.SSE_MASS_CHOLESKY


def _mats(c):
    ra, gf = c.sd.reference_approximation, c.sd.geometric_factors
    V = ra.V
    M = np.einsum("qa,kq,qb->kab", V, ra.W[None, :] * gf.J_q, V)            # V' WJ V
    Minv_wa = np.einsum("qa,kq,qb->kab", V, ra.W[None, :] / gf.J_q, V)      # weight-adjusted inverse
    return M, Minv_wa


@pytest.mark.parametrize("case", [lambda: cases.advection_2d(M=3, flux="lf"), lambda: cases.advection_3d(M=2, flux="lf")])
def test_oracle_cholesky_against_numpy(case):
    """StandardForm: dudt = M^-1 r with the same r for every mass solver, so
    dudt_chol = (V' WJ V)^-1 (V' (W/J) V)^-1 dudt_wa."""
    c = case()
    u = c.u0(seed=3)
    du_wa = oracle.rhs(c.image(), u)
    du_ch = oracle.rhs(c.image(mass_solver=CHOL), u)
    M, Minv_wa = _mats(c)
    r = np.linalg.solve(Minv_wa, du_wa.transpose(0, 2, 1))                  # (k, a, e)
    want = np.linalg.solve(M, r).transpose(0, 2, 1)
    assert np.abs(du_ch - want).max() <= 1e-11 * np.abs(want).max()
    assert np.abs(du_ch - du_wa).max() > 1e-8 * np.abs(want).max()          # curved mesh: the two solvers do differ


@pytest.mark.parametrize("case", [lambda: cases.euler_vortex_2d(M=3, p=4, flux="ec"), lambda: cases.euler_tgv_3d(M=2, p=3, flux="ec"),
                                  lambda: cases.advection_3d(M=2, flux="central")])
def test_invariants_with_the_exact_mass_matrix(case):
    """conservation / energy / entropy residuals vanish with CholeskySolver as they do with the default
    (runtests.jl:35-142 assertions, Analysis/conservation.jl:145-189 with mass_matrix(::CholeskySolver, k))."""
    c = case()
    img, u = c.image(mass_solver=CHOL), c.u0(seed=0)
    du = oracle.rhs(img, u)
    scale = max(1.0, np.abs(du).max())
    assert np.all(np.isfinite(du))
    assert np.abs(analysis.conservation_residual(img, du)).max() < 1e-10 * scale
    if c.law.pde_id == 0:
        assert np.abs(analysis.energy_residual(img, u, du)).max() < 1e-11 * scale
    if c.law.pde_id == 2:
        assert abs(analysis.entropy_residual(img, u, du)) < 1e-10 * scale


def test_cholesky_with_identity_V_is_the_diagonal_solver():
    c = cases.euler_periodic_3d_hex(M=2, p=3, flux="ec")                    # NodalTensor Lobatto: V = I
    u = c.u0(seed=1)
    img = c.image(mass_solver=CHOL)
    assert int(img.cfg.mass_solver) == _abi.SSE_MASS_DIAGONAL
    assert np.array_equal(oracle.rhs(img, u), oracle.rhs(c.image(), u))


def test_physical_operators_fold_the_cholesky_inverse():
    c = cases.advection_2d(M=3, flux="lf")
    u = c.u0(seed=1)
    a = oracle.rhs(c.image(mass_solver=CHOL), u)
    b = oracle.rhs(assemble(c.law, c.sd, c.form, PHYSICAL_OPERATOR, CHOL), u)
    assert np.abs(a - b).max() <= 1e-12 * np.abs(a).max()


GPU_CASES = {
    "advection_2d": lambda: cases.advection_2d(M=4, flux="lf"),
    "advection_3d": lambda: cases.advection_3d(M=2, flux="lf"),
    "euler_vortex_2d_ec": lambda: cases.euler_vortex_2d(M=4, p=4, flux="ec"),
    "euler_tgv_3d_lf": lambda: cases.euler_tgv_3d(M=2, flux="lf"),
    "euler_tgv_3d_p3": lambda: cases.euler_tgv_3d(M=2, p=3, flux="ec"),
    "advection_diffusion_2d": lambda: cases.advection_diffusion_2d(M=3),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GPU_CASES))
def test_gpu_cholesky_parity(name):
    import torch
    from sse_b200.solver import Solver
    c = GPU_CASES[name]()
    strategy = PHYSICAL_OPERATOR if c.law.second_order else c.strategy
    img = assemble(c.law, c.sd, c.form, strategy, CHOL)
    u = c.u0(seed=0)
    ref = oracle.rhs(img, u)
    s = Solver(img, 0)
    du = s.new_state()
    ud = torch.from_numpy(u).cuda()
    s.rhs(du, ud)
    s.synchronize()
    got = du.cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    if not c.law.second_order:
        f = np.asarray(s.functionals(ud, du))
        nc = c.law.N_c
        want_e = analysis.energy_residual(img, u, ref).sum()
        assert abs(f[nc] - want_e) <= 1e-9 * max(1.0, abs(want_e), np.abs(ref).max())
    s.close()


# ---- ViscousBurgersEquation (burgers.jl:23-49, 59-99): Burgers flux + the BR1 terms of the advection-diffusion law
def test_viscous_burgers_oracle_terms():
    """F(u, q) = a u^2/2 - b q (burgers.jl:60-70): with b -> 0 the residual is the inviscid Burgers StandardForm residual,
    and the viscous part is linear in b; the BR1 law dissipates the energy u^2/2 (conservation.jl:154-167)."""
    from sse_b200.laws import InviscidBurgersEquation, ViscousBurgersEquation
    c = cases.viscous_burgers_2d(M=3, p=3, b=5e-2)
    u = c.u0(seed=0)

    def rhs(law, strategy=PHYSICAL_OPERATOR):
        return oracle.rhs(assemble(law, c.sd, c.form, strategy), u)
    r0 = rhs(ViscousBurgersEquation((1.0, 1.0), 0.0))
    ri = rhs(InviscidBurgersEquation((1.0, 1.0)))
    assert np.abs(r0 - ri).max() <= 1e-12 * np.abs(ri).max()
    r1, r2 = rhs(ViscousBurgersEquation((1.0, 1.0), 5e-2)), rhs(ViscousBurgersEquation((1.0, 1.0), 1e-1))
    assert np.abs((r2 - r0) - 2 * (r1 - r0)).max() <= 1e-11 * np.abs(r2).max()
    assert np.abs(r1 - r0).max() > 1e-3 * np.abs(r0).max()
    img = c.image()
    du = oracle.rhs(img, u)
    assert np.abs(analysis.conservation_residual(img, du)).max() < 1e-11 * max(1.0, np.abs(du).max())
    # the viscous terms alone dissipate: u' M (r(b) - r(0)) < 0
    assert analysis.energy_residual(img, u, r1 - r0)[0] < 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", [lambda: cases.viscous_burgers_1d(M=8, p=5), lambda: cases.viscous_burgers_2d(M=3, p=4),
                                  lambda: cases.viscous_burgers_2d(M=3, p=3, kind="nodal")])
def test_gpu_viscous_burgers_parity(case):
    import torch
    from sse_b200.solver import Solver
    c = case()
    img, u = c.image(), c.u0(seed=0)
    ref = oracle.rhs(img, u)
    s = Solver(img, 0)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    s.synchronize()
    assert np.abs(du.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    s.close()
