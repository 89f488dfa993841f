"""Device-side GeometricFactors (sse_geometric_factors, csrc/kernels_geometry.cuh) against the NumPy restatement of
mesh.jl:229-506 in sse_b200/mesh.py, for every element type and metric the host mirror builds."""
import numpy as np
import pytest

from sse_b200 import mesh as M
from sse_b200.reference import ModalTensor, NodalTensor, reference_approximation

pytestmark = pytest.mark.gpu

CASES = {
    "line_exact": ("Line", NodalTensor(5), 5, None, "exact", 6, 2.0),
    "tri_exact": ("Tri", ModalTensor(4), 4, "delrey", "exact", 3, 1.0),
    "tri_curl": ("Tri", ModalTensor(3), 3, "chan", "curl", 4, 1.0),
    "quad_exact": ("Quad", NodalTensor(4), 4, "delrey", "exact", 3, 1.0),
    "quad_curl": ("Quad", NodalTensor(3), 3, "chan", "curl", 3, 1.0),
    "tet_exact": ("Tet", ModalTensor(3), 3, "delrey", "exact", 2, 1.0),
    "tet_curl": ("Tet", ModalTensor(4), 4, "chan", "curl", 2, 2 * np.pi),
    "hex_exact": ("Hex", NodalTensor(3), 3, "chan", "exact", 2, 2.0),
    "hex_curl": ("Hex", NodalTensor(4), 4, "chan", "curl", 2, 2.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_geometry_matches_host(name):
    elem, approx, pmap, warp, metric, cells, L = CASES[name]
    ra = reference_approximation(approx, elem, mapping_degree=pmap)
    d = ra.d
    w = None if warp is None else (M.DelReyWarping(0.1, (L,) * d) if warp == "delrey" else M.ChanWarping(1.0 / 16.0, (L,) * d))
    if d == 1:
        mesh = M.uniform_periodic_mesh(ra, (0.0, L), cells)
    else:
        mesh = M.uniform_periodic_mesh(ra, ((0.0, L),) * d, (cells,) * d, w)
    host = M.geometric_factors(mesh, ra, metric)
    dev = M.geometric_factors(mesh, ra, metric, device=0)
    for f in ("J_q", "Lambda_q", "J_f", "nJf", "nJq"):
        a, b = getattr(host, f), getattr(dev, f)
        assert a.shape == b.shape, f
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(a).max(), 1.0), f
    assert M.check_normals(mesh, dev) < 1e-10          # check_normals (SpatialDiscretizations.jl:457-470)


def test_device_geometry_feeds_the_solver():
    """A solver image assembled from device-computed metrics gives the oracle's residual."""
    import torch
    import oracle
    from sse_b200 import cases
    from sse_b200.assembly import SpatialDiscretization, assemble
    from sse_b200.solver import Solver
    c = cases.euler_tgv_3d(M=2, flux="lf")
    sd = SpatialDiscretization.build(c.sd.mesh, c.sd.reference_approximation, "curl", device=0)
    img = assemble(c.law, sd, c.form)
    u = c.u0(seed=2)
    s = Solver(img, 0)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    ref = oracle.rhs(c.image(), u)                     # image with host-computed metrics
    assert np.abs(du.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    s.close()
