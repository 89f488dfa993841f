"""Parity of the CUDA library (through the C ABI) with the CPU oracle on identical meshes and
states: max|Δ| <= 1e-12 * max|ref| (BASELINE.json north_star: 1e-12 relative, FP64), for every
BASELINE config, both kernel variants, plus golden vectors and size-independent invariants."""
import numpy as np
import pytest
import torch

import oracle
from sse_b200 import analysis, cases
from sse_b200.assembly import PHYSICAL_OPERATOR, assemble
from sse_b200.solver import Solver, solve_ck54, ODEProblem, semi_discrete_residual

pytestmark = pytest.mark.gpu
RTOL = 1.0e-12


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def gpu_rhs(img, u, variant=1, check_scratch=False):
    s = Solver(img, 0)
    s.set_kernel_variant(variant)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    s.synchronize()
    out = du.cpu().numpy()
    extra = None
    if check_scratch:
        uq, uf = s.debug_views()
        extra = (uq.cpu().numpy(), uf.cpu().numpy(), s.kernel_variant())
    s.close()
    return out, extra


CASES = {
    "advection_2d_lf": lambda: cases.advection_2d(M=4, flux="lf"),
    "advection_2d_central": lambda: cases.advection_2d(M=2, flux="lf0"),
    "advection_2d_nodal": lambda: cases.advection_2d(M=3, flux="lf", kind="nodal"),
    "euler_vortex_2d_p3_lf": lambda: cases.euler_vortex_2d(M=4, p=3, flux="lf"),
    "euler_vortex_2d_p4_ec": lambda: cases.euler_vortex_2d(M=4, p=4, flux="ec"),
    "euler_vortex_2d_nodal": lambda: cases.euler_vortex_2d(M=3, p=4, flux="ec", kind="nodal"),
    "advection_diffusion_2d": lambda: cases.advection_diffusion_2d(M=4),
    "advection_3d_central": lambda: cases.advection_3d(M=2, flux="central"),
    "advection_3d_lf": lambda: cases.advection_3d(M=2, flux="lf"),
    "advection_3d_p3": lambda: cases.advection_3d(M=2, p=3, flux="lf"),
    "euler_tgv_3d_lf": lambda: cases.euler_tgv_3d(M=2, flux="lf"),
    "euler_tgv_3d_ec": lambda: cases.euler_tgv_3d(M=2, flux="ec"),
    "euler_tgv_3d_p3": lambda: cases.euler_tgv_3d(M=2, p=3, flux="lf"),
    "euler_tgv_3d_nodal": lambda: cases.euler_tgv_3d(M=2, p=3, flux="ec", kind="nodal"),
    "euler_tgv_3d_M4": lambda: cases.euler_tgv_3d(M=4, flux="lf"),
    # the compile-time kernels are instantiated for N = p + 1 = 3 .. 8 (p = 7: examples/advection_3d.ipynb of the reference)
    "euler_tgv_3d_p6": lambda: cases.euler_tgv_3d(M=2, p=6, flux="lf"),
    "euler_tgv_3d_p7": lambda: cases.euler_tgv_3d(M=2, p=7, flux="ec"),
    "advection_3d_p6": lambda: cases.advection_3d(M=2, p=6, flux="central"),
    "advection_3d_p7": lambda: cases.advection_3d(M=2, p=7, flux="lf"),
    "euler_tgv_3d_p2": lambda: cases.euler_tgv_3d(M=2, p=2, flux="lf"),
    "euler_tgv_3d_p5": lambda: cases.euler_tgv_3d(M=2, p=5, flux="ec"),
    "advection_3d_p2": lambda: cases.advection_3d(M=2, p=2, flux="lf"),
    "advection_3d_p5": lambda: cases.advection_3d(M=2, p=5, flux="central"),
    "euler_vortex_2d_standard_lf": lambda: cases.euler_vortex_2d_standard(M=4, p=4, flux="lf"),
    "euler_vortex_2d_standard_nodal": lambda: cases.euler_vortex_2d_standard(M=3, p=3, flux="central", kind="nodal"),
    "euler_vortex_2d_standard_physical": lambda: cases.euler_vortex_2d_standard(M=3, p=4, flux="lf", strategy=PHYSICAL_OPERATOR),
    "euler_tgv_3d_standard_lf": lambda: cases.euler_tgv_3d_standard(M=2, p=4, flux="lf"),
    "euler_tgv_3d_standard_p3_central": lambda: cases.euler_tgv_3d_standard(M=2, p=3, flux="central"),
    "euler_tgv_3d_standard_physical": lambda: cases.euler_tgv_3d_standard(M=2, p=3, flux="lf", strategy=PHYSICAL_OPERATOR),
    # dense multidimensional operators (multidimensional.jl:1-75) through the generic kernels
    "advection_2d_modal_multi": lambda: cases.advection_2d(M=3, flux="lf", kind="modal_multi"),
    "advection_3d_nodal_multi": lambda: cases.advection_3d(M=2, p=3, flux="lf", kind="nodal_multi"),
    "euler_vortex_2d_modal_multi": lambda: cases.euler_vortex_2d(M=3, p=3, flux="ec", kind="modal_multi"),
    "euler_vortex_2d_nodal_multi": lambda: cases.euler_vortex_2d(M=3, p=4, flux="lf", kind="nodal_multi"),
    "euler_tgv_3d_modal_multi": lambda: cases.euler_tgv_3d(M=2, p=2, flux="ec", kind="modal_multi"),
    "euler_tgv_3d_nodal_multi": lambda: cases.euler_tgv_3d(M=2, p=2, flux="lf", kind="nodal_multi"),
    "burgers_1d_ec": lambda: cases.burgers_1d(M=8, p=7, flux="ec"),
    "burgers_1d_lf": lambda: cases.burgers_1d(M=8, p=5, flux="lf"),
    "advection_2d_quad": lambda: cases.advection_2d_quad(M=3, p=4, flux="lf"),
    "euler_3d_hex_ec": lambda: cases.euler_periodic_3d_hex(M=2, p=3, flux="ec"),
    "euler_3d_hex_lf": lambda: cases.euler_periodic_3d_hex(M=2, p=4, flux="lf"),
}


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("name", sorted(CASES))
def test_rhs_matches_oracle(name, variant):
    c = CASES[name]()
    img, u = c.image(), c.u0(seed=0)
    ref, uq_ref, uf_ref = oracle.rhs(img, u, return_scratch=True)
    got, (uq, uf, used) = gpu_rhs(img, u, variant, check_scratch=True)
    assert np.all(np.isfinite(got))
    assert relerr(uf[:, :uf_ref.shape[1] * uf_ref.shape[2]].reshape(uf_ref.shape), uf_ref) <= RTOL, "facet states (pass A)"
    assert relerr(got, ref) <= RTOL, f"dudt, variant requested {variant} used {used}"


def test_physical_first_order_and_stored_nJq():
    c = cases.advection_3d(M=2, flux="lf")
    u = c.u0(seed=3)
    img = assemble(c.law, c.sd, c.form, PHYSICAL_OPERATOR)
    assert relerr(gpu_rhs(img, u)[0], oracle.rhs(img, u)) <= RTOL
    c = cases.euler_tgv_3d(M=2, flux="ec")
    u = c.u0(seed=3)
    img = c.image(pass_nJq=True)
    assert relerr(gpu_rhs(img, u, 0)[0], oracle.rhs(img, u)) <= RTOL


def test_logmean_branches_are_both_exercised():
    """A state with large density jumps drives logmean through the log branch (f^2 >= 1e-4)."""
    c = cases.euler_vortex_2d(M=4, p=4, flux="lf")
    u = c.u0(seed=5, eps=0.01)
    img = c.image()
    ref = oracle.rhs(img, u)
    assert np.all(np.isfinite(ref))
    for v in (0, 1):
        assert relerr(gpu_rhs(img, u, v)[0], ref) <= RTOL


def _strong_gradient_state(c, amp):
    """L2 projection of a smooth state whose density and pressure vary by a factor (1 + amp) / (1 - amp) over the box."""
    from sse_b200.laws import project_function
    sd = c.sd

    def f(xyz):
        x, y, z = xyz[0], xyz[1], xyz[2]
        rho = 1.0 + amp * np.sin(x) * np.cos(y + 0.3) * np.cos(z)
        p = 10.0 * (1.0 + amp * np.cos(x - 0.2) * np.sin(y) * np.sin(z + 0.5))
        v = np.stack([np.sin(x) * np.cos(y), -np.cos(x) * np.sin(y), 0.3 * np.sin(z)], axis=-1)
        e = p / (c.law.gamma - 1.0) + 0.5 * rho * (v ** 2).sum(-1)
        return np.concatenate([rho[..., None], rho[..., None] * v, e[..., None]], axis=-1)

    return np.ascontiguousarray(project_function(f, sd.reference_approximation, sd.geometric_factors.J_q, sd.mesh.xyzq))


@pytest.mark.parametrize("amp,p", [(0.3, 4), (0.6, 4), (0.5, 3)])
def test_logmean_tiers_on_compile_time_path(amp, p):
    """Under-resolved 3-D Euler states drive the pair kernel's log-mean through all three tiers (degree-3 series,
    degree-10 series, log formula: physics.cuh logmean_pair_scaled*); the oracle takes the reference's two branches."""
    c = cases.euler_tgv_3d(M=2, p=p, flux="lf")
    img, u = c.image(), _strong_gradient_state(c, amp)
    ref = oracle.rhs(img, u)
    assert np.all(np.isfinite(ref))
    out, extra = gpu_rhs(img, u, 1, check_scratch=True)
    assert extra[2] == 2                       # compile-time kernels
    assert relerr(out, ref) <= RTOL


def test_golden_fixture():
    """tests/golden/euler_tgv_3d_M2.npz (made by tests/golden/make_golden.py from the pinned oracle)."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "euler_tgv_3d_M2.npz"))
    c = cases.euler_tgv_3d(M=2, flux="lf")
    got, _ = gpu_rhs(c.image(), z["u"])
    assert relerr(got, z["dudt"]) <= RTOL


def test_euler_1d_reference_golden_on_gpu():
    """The reference's own golden L2 errors (runtests.jl:89-96) reproduced by the CUDA path with the
    fused device-resident CarpenterKennedy2N54 integrator."""
    from test_oracle_goldens import EULER_1D_GOLDEN, euler_1d_setup
    from sse_b200.laws import project_function
    ra, mesh, sd, img, exact = euler_1d_setup()
    u0 = project_function(exact, ra, sd.geometric_factors.J_q, mesh.xyzq)
    s = Solver(img, 0)
    u = solve_ck54(ODEProblem(semi_discrete_residual, u0, (0.0, 2.0), s), 2.0 / 1000, 1000)
    ue = np.transpose(exact(mesh.xyzq), (0, 2, 1))
    l2 = np.sqrt(np.einsum("kei,ki,kei->e", ue - u, ra.W[None, :] * sd.geometric_factors.J_q, ue - u))
    assert np.allclose(l2, EULER_1D_GOLDEN, rtol=0, atol=1e-10)
    s.close()


GOLDEN_RUNS = {
    # name: (case builder, end time, CK54 steps, golden attribute in test_oracle_goldens, expected kernel variant)
    "advection_2d_tri": (lambda: cases.advection_2d(M=2, p=4, flux="lf0", warp=0.1), 1.0, 100, "ADVECTION_2D_TRI_GOLDEN", 2),
    "advection_2d_quad": (lambda: cases.advection_2d_quad(M=2, p=4, flux="lf", warp=0.1), 1.0, 100, "ADVECTION_2D_QUAD_GOLDEN", None),
    "euler_vortex_2d_modal_tri": (lambda: cases.euler_vortex_2d(M=4, p=3, flux="lf"), 2.5, 1000,
                                  "EULER_VORTEX_2D_MODAL_GOLDEN", 2),
    "euler_3d_hex": (lambda: cases.euler_periodic_3d_hex(M=2, p=4, flux="ec"), 2.0, 2500, "EULER_3D_HEX_GOLDEN", None),
}


@pytest.mark.parametrize("name", sorted(GOLDEN_RUNS))
def test_reference_goldens_2d_3d_on_gpu(name):
    """The reference's own 2-D / 3-D golden L2 errors (runtests.jl:38-80, 111-144; see test_oracle_goldens.py) reproduced
    by the CUDA path alone: device-resident fused CarpenterKennedy2N54 steps through the C ABI, no oracle involved."""
    import test_oracle_goldens as og
    build, T, n, gold, variant = GOLDEN_RUNS[name]
    c = build()
    s = Solver(c.image(), 0)
    assert variant is None or s.kernel_variant() == variant
    u = solve_ck54(ODEProblem(semi_discrete_residual, c.u0(), (0.0, T), s), T / n, n)
    assert np.allclose(og._l2_error(c, u), getattr(og, gold), rtol=0, atol=1e-10)
    s.close()


def test_invariants_at_scale_and_functionals():
    """Size-independent properties on a mesh the oracle is not run on: conservation and entropy
    conservation (EC interface flux) to roundoff, via the device functionals."""
    c = cases.euler_tgv_3d(M=8, flux="ec")
    img, u = c.image(), c.u0(seed=7)
    s = Solver(img, 0)
    du = s.new_state()
    ud = torch.from_numpy(u).cuda()
    s.rhs(du, ud)
    f = s.functionals(ud, du)
    d = du.cpu().numpy()
    scale = np.abs(d).max() * (2 * np.pi) ** 3
    assert np.abs(f[:5]).max() < 1e-12 * scale
    assert abs(f[6]) < 1e-12 * scale
    assert np.allclose(f[:5], analysis.conservation_residual(img, d), rtol=0, atol=1e-12 * scale)
    assert abs(f[6] - analysis.entropy_residual(img, u, d)) < 1e-12 * scale
    s.close()


@pytest.mark.parametrize("name", ["advection_2d_lf", "advection_2d_central", "advection_3d_lf", "euler_tgv_3d_lf",
                                  "burgers_1d_ec", "euler_3d_hex_lf"])
def test_energy_functional_matches_host_analysis(name):
    """out[N_c] of sse_functionals = sum_e u' M dudt (Analysis/conservation.jl:154-167) with the mass matrix of the
    solver: diag(W J) for collocated schemes, the inverse of the weight-adjusted inverse for modal ones (solved by
    conjugate gradients on the device)."""
    c = CASES[name]()
    img, u = c.image(), c.u0(seed=9)
    s = Solver(img, 0)
    du = s.new_state()
    ud = torch.from_numpy(u).cuda()
    s.rhs(du, ud)
    f = s.functionals(ud, du)
    d = du.cpu().numpy()
    ref = analysis.energy_residual(img, u, d)
    scale = np.abs(np.einsum("kea,kea->", np.abs(u), np.abs(d))) + 1e-300
    assert np.isfinite(f[img.cfg.N_c])
    assert abs(f[img.cfg.N_c] - ref.sum()) < 1e-11 * max(scale, abs(ref.sum()))
    if name == "advection_2d_central":
        assert abs(f[1]) < 1e-10                  # energy conservation with the central flux (runtests.jl:59)
    s.close()


def test_host_buffer_api_and_linearity():
    c = cases.advection_3d(M=2, flux="lf")
    img = c.image()
    s = Solver(img, 0)
    u1, u2 = c.u0(seed=1), c.u0(seed=2)
    r = [semi_discrete_residual(np.empty_like(u1), x, s) for x in (u1, u2, 2.0 * u1 - 3.0 * u2)]
    assert relerr(r[2], 2.0 * r[0] - 3.0 * r[1]) <= 1e-12          # the advection residual is linear
    assert relerr(r[0], oracle.rhs(img, u1)) <= RTOL
    s.close()


@pytest.mark.parametrize("chunks", [3, 8, 16])
def test_pipelined_host_residual_is_bitwise_the_device_residual(chunks):
    """rhs_host overlaps uploads, pass A / pass B of element ranges and downloads on three streams, ordering pass B by
    the face-neighbour dependencies read from mapP; the result must be bit-identical to the plain device residual."""
    c = cases.euler_tgv_3d(M=4, flux="lf")
    img, u = c.image(), c.u0(seed=11)
    s = Solver(img, 0)
    bounds, up, after = s._chunk_plan(chunks)
    assert sorted(up) == list(range(chunks)) and sorted(k for a in after for k in a) == list(range(chunks))
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    s.synchronize()
    for _ in range(2):
        out = s.rhs_host(np.empty_like(u), u, chunks=chunks)
        assert np.array_equal(out, du.cpu().numpy())
    hu = torch.from_numpy(u).pin_memory()
    hdu = torch.empty_like(hu).pin_memory()
    s.rhs_host(hdu, hu, chunks=chunks)
    assert np.array_equal(hdu.numpy(), du.cpu().numpy())
    s.close()


@pytest.mark.parametrize("case", [lambda: cases.euler_tgv_3d(M=2, flux="lf"), lambda: cases.euler_tgv_3d(M=2, p=3, flux="ec"),
                                  lambda: cases.advection_3d(M=2, flux="lf"), lambda: cases.euler_vortex_2d(M=4, p=4, flux="ec")])
def test_ragged_and_empty_element_ranges(case):
    """The range entry points (the pipelined host call and the multi-GPU driver use them) on ragged splits: counts that
    are not multiples of the 6 / 8 / 24 elements a projection CTA batches, single elements and empty ranges must
    reproduce the whole-mesh residual bit for bit."""
    c = case()
    img, u = c.image(), c.u0(seed=5)
    s = Solver(img, 0)
    ud = torch.from_numpy(u).cuda()
    ref = s.new_state()
    s.rhs(ref, ud)
    s.synchronize()
    ne = c.sd.N_e
    cuts = sorted({0, 1, 1, 6, 13, 13, 20, 37, ne - 7, ne - 1, ne})          # includes empty ranges
    cuts = [x for x in cuts if 0 <= x <= ne]
    du = s.new_state()
    du.fill_(float("nan"))
    s.pass_a_range(ud, 0, 0)
    for a, b in zip(cuts[:-1], cuts[1:]):
        s.pass_a_range(ud, a, b - a)
    s.pass_b(du, 0, 0)
    for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):
        s.pass_b(du, a, b - a)
    s.synchronize()
    assert torch.equal(du, ref)
    s.close()


def test_bad_arguments_fail_loudly():
    from sse_b200._lib import SSEError
    c = cases.advection_2d(M=2)
    img = c.image()
    s = Solver(img, 0)
    with pytest.raises(ValueError):
        s.rhs(s.new_state(), torch.zeros(3, dtype=torch.float64, device="cuda"))
    img.arrays["mapP"] = img.arrays["mapP"].copy()
    img.arrays["mapP"][0] = 10 ** 9
    with pytest.raises(SSEError):
        Solver(img, 0)
    s.close()


@pytest.mark.parametrize("name", ["euler_tgv_3d_lf", "advection_3d_lf", "euler_vortex_2d_p4_ec"])
def test_fused_rk_step_equals_unfused(name):
    """sse_step_ck54 (stage update fused into the projection epilogue on the compile-time path) against the
    explicit rhs + lsrk_stage sequence: same arithmetic, so identical to roundoff."""
    from sse_b200.solver import CK54_A, CK54_B, CK54_C
    c = CASES[name]()
    img, u0 = c.image(), c.u0(seed=4)
    s = Solver(img, 0)
    dt = 1e-4
    ua, ub = torch.from_numpy(u0).cuda(), torch.from_numpy(u0).cuda()
    ta, tb, du = s.new_state(), s.new_state(), s.new_state()
    for _ in range(2):
        s.step_ck54(ua, ta, du, 0.0, dt)
        for st in range(5):
            s.rhs(du, ub, CK54_C[st] * dt)
            s.lsrk_stage(ub, tb, du, CK54_A[st], CK54_B[st], dt)
    s.synchronize()
    a, b = ua.cpu().numpy(), ub.cpu().numpy()
    assert np.all(np.isfinite(a))
    assert relerr(a, b) <= 1e-14
    assert relerr(a, u0) > 1e-9          # the state really moved
    s.close()


@pytest.mark.parametrize("name,variant", [("euler_tgv_3d", 1), ("euler_tgv_3d", 0), ("euler_vortex_2d", 1)])
def test_nonphysical_state_is_reported(name, variant):
    """A state outside the physical domain (negative pressure) makes the reference throw a DomainError from log / sqrt
    (SURVEY.md §8b); the library reports SSE_ERR_NONFINITE at the next blocking call, once, and keeps working afterwards."""
    from sse_b200._lib import SSEError
    c = cases.BUILDERS[name](M=2 if name == "euler_tgv_3d" else 3)
    img, good = c.image(), c.u0(seed=0)
    bad = good.copy()
    bad[1, -1, :] = -np.abs(bad[1, -1, :]) - 1.0          # total energy of one element negative -> p < 0
    s = Solver(img, 0)
    s.set_kernel_variant(variant)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(bad).cuda())
    with pytest.raises(SSEError) as e:
        s.synchronize()
    assert e.value.code == 4 and "DomainError" in str(e.value)
    s.synchronize()                                         # reported once
    s.rhs(du, torch.from_numpy(good).cuda())
    s.synchronize()
    assert relerr(du.cpu().numpy(), oracle.rhs(img, good)) <= RTOL
    hb = torch.from_numpy(bad).pin_memory()
    hd = torch.empty_like(hb).pin_memory()
    with pytest.raises(SSEError) as e:                      # the synchronous host-buffer residual reports it itself
        s.rhs_host(hd, hb)
    assert e.value.code == 4
    s.close()


def test_config5_at_24576_elements_matches_oracle():
    """BASELINE config 5 at M = 16 (24 576 curved tets, 4.3 M DOF: beyond the L2, 166 elements per SM) against the oracle on
    the same mesh and state -- the largest size the oracle finishes in about a second; larger sizes are covered by the
    size-independent invariants (test_invariants_at_scale_and_functionals)."""
    c = cases.euler_tgv_3d(M=16, flux="lf")
    img, u = c.image(), c.u0(seed=0)
    ref = oracle.rhs(img, u)
    got, _ = gpu_rhs(img, u, 1)
    assert relerr(got, ref) <= RTOL
    hu = torch.from_numpy(u).pin_memory()
    hd = torch.empty_like(hu).pin_memory()
    s = Solver(img, 0)
    s.rhs_host(hd, hu, chunks=7)                      # the pipelined host-buffer residual returns the same bits
    s.close()
    assert np.array_equal(hd.numpy(), got)


@pytest.mark.parametrize("first,count", [(0, 384), (5, 100), (7, 1), (101, 283), (380, 4)])
def test_config4_fused_path_on_element_ranges(first, count):
    """The fused advection kernels work on tasks of 32/N elements; a pass over an arbitrary element range (the multi-GPU
    interior / halo split, the ranges of the pipelined host-buffer residual) must touch exactly that range -- including
    ranges that start or end inside a task and an odd number of tasks (spare warp of the last CTA)."""
    c = cases.advection_3d(M=4, flux="lf")
    img, u = c.image(), c.u0(seed=1)
    ref = oracle.rhs(img, u)
    s = Solver(img, 0)
    assert s.kernel_variant() == 2
    du = s.new_state()
    du.fill_(777.0)
    ud = torch.from_numpy(u).cuda()
    s.pass_a(ud)
    s.pass_b(du, first, count)
    s.synchronize()
    got = du.cpu().numpy()
    assert relerr(got[first:first + count], ref[first:first + count]) <= RTOL
    mask = np.ones(got.shape[0], dtype=bool)
    mask[first:first + count] = False
    assert np.all(got[mask] == 777.0)
    hu = torch.from_numpy(u).pin_memory()
    hd = torch.empty_like(hu).pin_memory()
    s.rhs_host(hd, hu, chunks=7)
    s.close()
    assert relerr(hd.numpy(), ref) <= RTOL


def test_ck54_stage_fused_kernels_match_the_unfused_sequence():
    """sse_step_ck54 on the compile-time Euler path runs each stage as pair kernel + ONE kernel that finishes the stage and starts
    the next (k_nodal_ct<FUSED>); the result must equal residual + sse_lsrk_stage per stage (solve_ck54(fused=False)) to
    round-off, on a mesh with a partial last CTA (384 elements = 64 CTAs of 6) and one without (48 elements)."""
    for M in (2, 4):
        c = cases.euler_tgv_3d(M=M, flux="lf")
        img, u = c.image(), c.u0(seed=0)
        s = Solver(img, 0)
        a = solve_ck54(ODEProblem(semi_discrete_residual, u, (0.0, 1.0), s), 1e-3, 3, fused=True)
        b = solve_ck54(ODEProblem(semi_discrete_residual, u, (0.0, 1.0), s), 1e-3, 3, fused=False)
        s.close()
        assert np.all(np.isfinite(a))
        assert relerr(a, b) <= 1e-13


@pytest.mark.parametrize("p", [2, 3, 4, 5, 6, 7])
def test_compile_time_kernels_cover_p2_to_p7(p):
    """ModalTensor(p) tets run the compile-time kernels for p = 2 .. 7 (kernel variant 2), for the Euler flux-differencing path
    and for the fused advection path, and the device-resident CarpenterKennedy2N54 step agrees with the unfused sequence."""
    c = cases.euler_tgv_3d(M=2, p=p, flux="lf")
    img, u = c.image(), c.u0(seed=0)
    s = Solver(img, 0)
    assert s.kernel_variant() == 2
    a = solve_ck54(ODEProblem(semi_discrete_residual, u, (0.0, 1.0), s), 1e-3, 2, fused=True)
    b = solve_ck54(ODEProblem(semi_discrete_residual, u, (0.0, 1.0), s), 1e-3, 2, fused=False)
    s.close()
    assert relerr(a, b) <= 1e-13
    c = cases.advection_3d(M=2, p=p, flux="lf")
    s = Solver(c.image(), 0)
    assert s.kernel_variant() == 2
    s.close()


@pytest.mark.parametrize("name", ["euler_vortex_2d", "euler_tgv_3d", "advection_diffusion_2d", "advection_3d"])
def test_graph_replay_of_the_ck54_step_is_bit_identical(name):
    """sse_set_graph_mode: the launches of one CarpenterKennedy2N54 step captured into a CUDA graph and replayed (re-captured
    when dt or a buffer changes) give the same bits as launching them one by one -- for the compile-time, tensor-line and
    generic (BR1) kernel families."""
    c = cases.BUILDERS[name](M=4 if name.endswith("2d") else 2)
    img, u0 = c.image(), c.u0(seed=0)
    outs = []
    for graph in (False, True):
        s = Solver(img, 0)
        s.use_current_stream()
        s.set_graph_mode(graph)
        u, tmp, du = torch.from_numpy(u0).cuda(), s.new_state(), s.new_state()
        t = 0.0
        for dt in (1e-3, 1e-3, 1e-3, 5e-4, 5e-4):
            s.step_ck54(u, tmp, du, t, dt)
            t += dt
        u2 = u.clone()                                     # another buffer: the graph is re-captured
        tmp.zero_()
        s.step_ck54(u2, tmp, du, t, 1e-3)
        s.synchronize()
        outs.append((u.cpu().numpy(), u2.cpu().numpy(), s.launches))
        s.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]                        # the replay accounts for the launches it contains
    assert np.all(np.isfinite(outs[1][1]))


def test_element_packing_is_bitwise_neutral():
    """Small elements run several per CTA, one warp each (common.cuh: sse_element / sse_row_smem / sse_sync).  Rows are
    independent, so the residual must not depend on the launch shape: one element per CTA, 64 threads per element (the
    round-1 shape) and the packed default give the same bits; element counts and ranges leave one-row remainder launches."""
    import json
    import os
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "packing_worker.py")
    outs = []
    for extra in ({}, {"SSE_PACK_ROWS": "1"}, {"SSE_PACK_ROWS": "1", "SSE_THREADS_MIN": "64"}, {"SSE_PACK_ROWS": "3"}):
        env = dict(os.environ, **extra)
        p = subprocess.run([sys.executable, worker], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1] == outs[2] == outs[3], outs
