"""Warp-per-element kernels of the 2-D Euler flux-differencing path on triangles (kernels_tri.cuh, BASELINE config 2:
test/euler_vortex_2d_modal.jl): parity with the CPU oracle (1e-12 relative) of the volume and facet states of pass A and of
dudt, for every compiled degree, both interface fluxes, stored normals, arbitrary element ranges, meshes that give a warp
several elements (grid-stride loop with the next element's loads in flight), the fused 2N-storage stage, and bitwise
agreement between launch shapes."""
import numpy as np
import pytest
import torch

import oracle
from sse_b200 import cases
from sse_b200.solver import Solver

pytestmark = pytest.mark.gpu
RTOL = 1.0e-12


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run(img, u, variant=1):
    s = Solver(img, 0)
    s.set_kernel_variant(variant)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    s.synchronize()
    uq, uf = s.debug_views()
    out = du.cpu().numpy(), uq.cpu().numpy().copy(), uf.cpu().numpy().copy(), s.kernel_variant()
    s.close()
    return out


@pytest.mark.parametrize("p", [2, 3, 4])
@pytest.mark.parametrize("flux", ["lf", "ec"])
def test_triangle_kernels_match_oracle(p, flux):
    c = cases.euler_vortex_2d(M=4, p=p, flux=flux)
    img, u = c.image(), c.u0(seed=p)
    ref, uq_ref, uf_ref = oracle.rhs(img, u, return_scratch=True)
    got, uq, uf, used = run(img, u)
    assert used == 2, "the compile-time (warp-per-element) path was not selected"
    assert np.all(np.isfinite(got))
    # pass B of this path leaves the volume states of pass A in the scratch (nothing is handed over through it)
    assert relerr(uq.reshape(uq_ref.shape), uq_ref) <= RTOL, "volume states (pass A)"
    assert relerr(uf[:, :uf_ref.shape[1] * uf_ref.shape[2]].reshape(uf_ref.shape), uf_ref) <= RTOL, "facet states (pass A)"
    assert relerr(got, ref) <= RTOL
    # the runtime tensor-line kernels on the same input
    got0, _, _, used0 = run(img, u, 0)
    assert used0 == 0 and relerr(got0, ref) <= RTOL


def test_triangle_kernels_nodal_and_p5_fall_back():
    for c in (cases.euler_vortex_2d(M=3, p=4, kind="nodal"), cases.euler_vortex_2d(M=2, p=5)):
        img, u = c.image(), c.u0(seed=1)
        got, _, _, used = run(img, u)
        assert used == 1
        assert relerr(got, oracle.rhs(img, u)) <= RTOL


def test_triangle_kernels_with_stored_normals_and_strong_gradients():
    c = cases.euler_vortex_2d(M=4, p=4, flux="lf")
    u = c.u0(seed=5, eps=0.01)                     # drives the log-mean through its log branch
    img = c.image(pass_nJq=True)
    got, _, _, used = run(img, u)
    assert used == 2
    assert relerr(got, oracle.rhs(img, u)) <= RTOL


def test_triangle_kernels_many_elements_per_warp():
    """5 000 elements on at most 592 CTAs of 4 warps: every warp walks over two or three elements, the last ones over a
    clamped prefetch."""
    c = cases.euler_vortex_2d(M=50, p=4, flux="lf")
    img, u = c.image(), c.u0(seed=2)
    ref = oracle.rhs(img, u)
    got, _, _, used = run(img, u)
    assert used == 2
    assert relerr(got, ref) <= RTOL


@pytest.mark.parametrize("first,count", [(0, 1), (5, 3), (7, 120), (100, 28), (0, 128)])
def test_triangle_kernels_on_element_ranges(first, count):
    c = cases.euler_vortex_2d(M=8, p=4, flux="lf")
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
    # pass B is a pure function of the scratch of pass A: a second call gives the same bits
    du2 = s.new_state()
    s.pass_b(du2, first, count)
    s.synchronize()
    assert np.array_equal(du2.cpu().numpy()[first:first + count], got[first:first + count])
    # pass A on a range writes that range of the scratch only, with the same bits as the full pass
    uq, uf = s.debug_views()
    uq_full, uf_full = uq.cpu().numpy().copy(), uf.cpu().numpy().copy()
    uq.fill_(0.0)
    s.pass_a_range(ud, first, count)
    s.synchronize()
    uq2 = uq.cpu().numpy().reshape(ref.shape[0], -1)
    assert np.array_equal(uq2[first:first + count], uq_full.reshape(ref.shape[0], -1)[first:first + count])
    assert np.all(uq2[mask] == 0.0)
    hu = torch.from_numpy(u).pin_memory()
    hd = torch.empty_like(hu).pin_memory()
    s.rhs_host(hd, hu, chunks=5)
    s.close()
    assert relerr(hd.numpy(), ref) <= RTOL


def test_triangle_kernels_fused_rk_stage_and_step():
    """sse_rhs_lsrk on this path carries the 2N-storage update in the epilogue of the fused pass-B kernel; five stages must
    reproduce residual + sse_lsrk_stage of the generic kernels."""
    c = cases.euler_vortex_2d(M=6, p=4, flux="lf")
    img, u0 = c.image(), c.u0(seed=4)
    outs = []
    for variant in (1, 0):
        s = Solver(img, 0)
        s.set_kernel_variant(variant)
        u = torch.from_numpy(u0.copy()).cuda()
        tmp, du = s.new_state(), s.new_state()
        tmp.zero_()
        for _ in range(3):
            s.step_ck54(u, tmp, du, 0.0, 1.0e-3)
        s.synchronize()
        outs.append(u.cpu().numpy())
        s.close()
    assert relerr(outs[0], outs[1]) <= RTOL
    assert np.abs(outs[0] - u0).max() > 1e-6


# ---- BASELINE config 1: 2-D linear advection, StandardForm + ReferenceOperators (k_tri_adv_facets / k_tri_adv)

@pytest.mark.parametrize("p", [2, 3, 4])
@pytest.mark.parametrize("flux", ["lf", "lf0", "central"])
def test_triangle_advection_kernels_match_oracle(p, flux):
    c = cases.advection_2d(M=4, p=p, flux=flux)
    img, u = c.image(), c.u0(seed=p)
    ref, uq_ref, uf_ref = oracle.rhs(img, u, return_scratch=True)
    got, _, uf, used = run(img, u)
    assert used == 2, "the warp-per-element path was not selected"
    assert relerr(uf[:, :uf_ref.shape[1] * uf_ref.shape[2]].reshape(uf_ref.shape), uf_ref) <= RTOL, "facet states (pass A)"
    assert relerr(got, ref) <= RTOL
    got0, _, _, used0 = run(img, u, 0)
    assert used0 == 0 and relerr(got0, ref) <= RTOL


def test_triangle_advection_kernels_many_elements_and_ranges():
    c = cases.advection_2d(M=50, p=4, flux="lf")                  # 5 000 elements: several per warp
    img, u = c.image(), c.u0(seed=2)
    ref = oracle.rhs(img, u)
    s = Solver(img, 0)
    assert s.kernel_variant() == 2
    ud = torch.from_numpy(u).cuda()
    du = s.new_state()
    s.rhs(du, ud)
    s.synchronize()
    assert relerr(du.cpu().numpy(), ref) <= RTOL
    first, count = 1234, 777
    du.fill_(777.0)
    s.pass_a(ud)
    ud.fill_(float("nan"))                                        # pass B must not read the caller's state again
    s.pass_b(du, first, count)
    s.synchronize()
    got = du.cpu().numpy()
    assert relerr(got[first:first + count], ref[first:first + count]) <= RTOL
    mask = np.ones(got.shape[0], dtype=bool)
    mask[first:first + count] = False
    assert np.all(got[mask] == 777.0)
    hu = torch.from_numpy(u).pin_memory()
    hd = torch.empty_like(hu).pin_memory()
    s.rhs_host(hd, hu, chunks=5)
    s.close()
    assert relerr(hd.numpy(), ref) <= RTOL


def test_triangle_advection_kernels_fused_rk_step():
    c = cases.advection_2d(M=6, p=4, flux="lf")
    img, u0 = c.image(), c.u0(seed=4)
    outs = []
    for variant in (1, 0):
        s = Solver(img, 0)
        s.set_kernel_variant(variant)
        u = torch.from_numpy(u0.copy()).cuda()
        tmp, du = s.new_state(), s.new_state()
        tmp.zero_()
        for _ in range(3):
            s.step_ck54(u, tmp, du, 0.0, 1.0e-3)
        s.synchronize()
        outs.append(u.cpu().numpy())
        s.close()
    assert relerr(outs[0], outs[1]) <= RTOL
    assert np.abs(outs[0] - u0).max() > 1e-6
