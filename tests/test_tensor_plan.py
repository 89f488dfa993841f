"""Host-only replay of the tensor-line pair schedule used by the specialised flux-differencing kernel:
every non-zero of S_m and C is visited exactly once with the right weight (no GPU needed)."""
import ctypes as C

import numpy as np
import pytest

from sse_b200 import _lib, cases


def selfcheck(img):
    info = (C.c_int32 * 8)()
    err = C.c_double(0.0)
    arr = img.c_arrays()
    _lib.check(_lib.load().sse_plan_selfcheck(C.byref(img.cfg), C.byref(arr), info, C.byref(err)))
    return list(info), err.value


@pytest.mark.parametrize("case,evals", [
    (lambda: cases.euler_tgv_3d(M=2, p=4), 750 + 1000 + 100), (lambda: cases.euler_tgv_3d(M=2, p=3), 288 + 448 + 64),
    (lambda: cases.euler_tgv_3d(M=2, p=2), None), (lambda: cases.euler_tgv_3d(M=2, p=3, kind="nodal"), None),
    (lambda: cases.euler_vortex_2d(M=2, p=4), 100 + 75 + 15), (lambda: cases.euler_vortex_2d(M=2, p=3), None),
    (lambda: cases.euler_vortex_2d(M=2, p=4, kind="nodal"), None)])
def test_schedule_replays_operators(case, evals):
    img = case().image()
    info, err = selfcheck(img)
    assert info[0] in (1, 2, 4), "tensor-line structure not detected"
    assert err == 0.0
    if evals is not None:
        assert info[7] == evals          # unique two-point fluxes per element


def test_p4_tet_schedule_shape():
    info, err = selfcheck(cases.euler_tgv_3d(M=2, p=4).image())
    assert info[1:6] == [128, 6, 8, 25, 5] and info[6] <= 48 * 1024


def test_kernel_family_selection():
    """Which kernels sse_create will pick (host-side decision, checked without a GPU)."""
    assert selfcheck(cases.euler_tgv_3d(M=2, p=4).image())[0][0] == 2       # headline: compile-time kernels
    assert selfcheck(cases.euler_tgv_3d(M=2, p=3).image())[0][0] == 2
    assert selfcheck(cases.euler_tgv_3d(M=2, p=2).image())[0][0] == 2       # compile-time kernels exist for p = 2 .. 7
    assert selfcheck(cases.euler_tgv_3d(M=2, p=5).image())[0][0] == 2
    assert selfcheck(cases.euler_tgv_3d(M=2, p=7).image())[0][0] == 2       # ... and up to p = 7
    assert selfcheck(cases.advection_3d(M=2, p=7).image())[0][0] == 3       # examples/advection_3d.ipynb of the reference
    assert selfcheck(cases.euler_tgv_3d(M=2, p=1).image())[0][0] == 1       # other degrees: runtime tensor-line kernel
    assert selfcheck(cases.euler_tgv_3d(M=2, p=3, kind="nodal").image())[0][0] == 1
    for p in (2, 3, 4):                                                     # config 2: warp-per-element triangle kernels
        assert selfcheck(cases.euler_vortex_2d(M=2, p=p).image())[0][0] == 4
    assert selfcheck(cases.euler_vortex_2d(M=2, p=5).image())[0][0] == 1       # 36 nodes do not fit a warp
    assert selfcheck(cases.euler_vortex_2d(M=2, p=4, kind="nodal").image())[0][0] == 1
    assert selfcheck(cases.advection_3d(M=2).image())[0][0] == 3            # config 4: compile-time StandardForm path
    assert selfcheck(cases.advection_3d(M=2, p=2).image())[0][0] == 3
    assert selfcheck(cases.advection_3d(M=2, p=1).image())[0][0] == 0       # generic
    for p in (2, 3, 4):                                                     # config 1: warp-per-element triangle kernels
        assert selfcheck(cases.advection_2d(M=2, p=p).image())[0][0] == 5
    assert selfcheck(cases.advection_2d(M=2, p=5).image())[0][0] == 0
    assert selfcheck(cases.advection_2d(M=2, kind="nodal").image())[0][0] == 0
    assert selfcheck(cases.advection_diffusion_2d(M=2).image())[0][0] == 0


def test_ct_nmax_switch_moves_high_degrees_to_the_runtime_kernels(monkeypatch):
    """SSE_CT_NMAX (the A/B switch of tools/bench_highp.py): degrees above it leave the compile-time families."""
    img_e, img_a = cases.euler_tgv_3d(M=2, p=7).image(), cases.advection_3d(M=2, p=7).image()
    assert selfcheck(img_e)[0][0] == 2 and selfcheck(img_a)[0][0] == 3
    monkeypatch.setenv("SSE_CT_NMAX", "6")
    assert selfcheck(img_e)[0][0] in (0, 1) and selfcheck(img_a)[0][0] == 0
    assert selfcheck(cases.euler_tgv_3d(M=2, p=4).image())[0][0] == 2       # the headline degree is not affected


def test_host_pipeline_chunk_plan():
    """Schedule of Solver.rhs_host: every range is uploaded once, its pass B is released exactly when pass A has covered
    all ranges holding one of its face neighbours, and on the slab-ordered periodic mesh only three ranges wait for the
    last upload."""
    from sse_b200.solver import chunk_plan
    c = cases.euler_tgv_3d(M=8, p=2)
    img = c.image()
    ne, nf = int(img.cfg.N_e), int(img.cfg.N_f)
    mapP = np.asarray(img.arrays["mapP"])
    nb = (mapP.reshape(ne, nf) - 1) // nf
    for chunks in (3, 8, 16):
        bounds, up, after = chunk_plan(mapP, ne, nf, chunks)
        assert sorted(up) == list(range(chunks)) and bounds[0] == 0 and bounds[-1] == ne
        released = [k for a in after for k in a]
        assert sorted(released) == list(range(chunks))
        done = set()
        for i, cth in enumerate(up):
            done |= set(range(bounds[cth], bounds[cth + 1]))
            for k in after[i]:
                need = set(nb[bounds[k]:bounds[k + 1]].reshape(-1).tolist()) | set(range(bounds[k], bounds[k + 1]))
                assert need <= done, "pass B released before pass A covered its neighbours"
        if chunks == 8:                      # one cube layer per range: a range depends on its two neighbours
            assert len(after[-1]) == 3 and all(len(a) <= 1 for a in after[:-1])


@pytest.mark.parametrize("chunks", [3, 8, 16])
def test_library_range_plan_of_the_host_buffer_residual(chunks):
    """sse_host_range_plan is the schedule sse_rhs_host follows (no device needed): the wrap-around range is uploaded first,
    every range is uploaded once, and pass B of a range is released exactly when pass A has covered the ranges holding its
    face neighbours — the same answer as the NumPy restatement `range_plan` for the same upload order."""
    import ctypes as C
    from sse_b200 import _lib
    from sse_b200.solver import range_plan
    for c in (cases.euler_tgv_3d(M=4, flux="lf"), cases.euler_vortex_2d(M=8, p=3, flux="lf")):
        img = c.image()
        mapP = np.ascontiguousarray(np.asarray(img.arrays["mapP"]).reshape(-1), dtype=np.int64)
        ne, nf, nfac = c.sd.N_e, int(img.cfg.N_f), int(img.cfg.N_fac)
        order, ready = (C.c_int32 * chunks)(), (C.c_int32 * chunks)()
        rc = _lib.load().sse_host_range_plan(mapP.ctypes.data_as(C.POINTER(C.c_int64)), ne, nf, nfac, chunks, order, ready)
        assert rc == 0
        order, ready = list(order), list(ready)
        assert sorted(order) == list(range(chunks)) and order[0] == chunks - 1
        bounds = [ne * k // chunks for k in range(chunks + 1)]
        after = range_plan(mapP, ne, nf, [(bounds[k], bounds[k + 1]) for k in order])
        want = [None] * chunks
        for i, ks in enumerate(after):
            for k in ks:
                want[order[k]] = i
        assert ready == want


def test_compile_time_path_is_refused_for_operators_without_its_structure():
    """The compile-time kernels hard-code structure that sse_create verifies on the host before selecting them (ct_eligible,
    ct_facet_factors, c_tensor_symmetric): R must be the Kronecker product of its 1-D factors entry by entry, and the C tensor
    of the collapsed tet must depend on (i + j, k) only.  An operator image that breaks either falls back to a general path."""
    def perturbed(case, name, where, delta):
        img = case.image()
        a = np.array(img.arrays[name], dtype=np.float64, copy=True)
        flat = a.reshape(-1)
        flat[where(flat)] += delta
        img.arrays[name] = a
        return img
    nonzero = lambda f: int(np.flatnonzero(f)[7])
    zero = lambda f: int(np.flatnonzero(f == 0.0)[11])
    euler, adv = cases.euler_tgv_3d(M=2, p=4), cases.advection_3d(M=2)
    assert selfcheck(euler.image())[0][0] == 2 and selfcheck(adv.image())[0][0] == 3
    assert selfcheck(perturbed(euler, "R", nonzero, 1e-9))[0][0] != 2          # an entry off its structured value
    assert selfcheck(perturbed(euler, "R", zero, 1e-13))[0][0] != 2            # a structural zero that is not zero
    assert selfcheck(perturbed(adv, "R", nonzero, 1e-9))[0][0] != 3
    assert selfcheck(perturbed(euler, "C", nonzero, 1e-15))[0][0] != 2         # C[a3, i, j, k] no longer a function of (i + j, k)
