"""Host setup restated from the reference: structural self-checks the reference ships
(check_sbp_property / check_normals / check_facet_nodes, SpatialDiscretizations.jl:424-470)
and the sparsity facts of SURVEY.md Appendix A."""
import numpy as np
import pytest

from sse_b200.mesh import (ChanWarping, DelReyWarping, check_facet_nodes, check_normals,
                           check_sbp_property_physical, geometric_factors, structured_connectivity,
                           uniform_periodic_mesh)
from sse_b200.quadrature import GaussQuadrature, LGLQuadrature, LGQuadrature, quadrature_line
from sse_b200.reference import ModalMulti, ModalTensor, NodalTensor, reference_approximation


def test_quadrature_exactness():
    x, w = quadrature_line(GaussQuadrature(4, 1, 0))
    for k in range(10):   # exact for (1 - x) x^k, k <= 2*5 - 1
        exact = (1 - (-1) ** (k + 1)) / (k + 1) - (1 - (-1) ** (k + 2)) / (k + 2)
        assert abs(np.sum(w * x ** k) - exact) < 1e-14
    x, w = quadrature_line(LGLQuadrature(4))
    assert x[0] == -1 and x[-1] == 1 and abs(w.sum() - 2) < 1e-14
    assert abs(np.sum(w * x ** 6) - 2 / 7) < 1e-14


@pytest.mark.parametrize("elem,approx,npq", [("Tri", ModalTensor(4), (15, 25, 15)), ("Tet", ModalTensor(4), (35, 125, 100)),
                                             ("Tri", NodalTensor(3), (16, 16, 12)), ("Tet", NodalTensor(3), (64, 64, 64)),
                                             ("Tet", ModalTensor(3), (20, 64, 64))])
def test_reference_sbp_and_orthonormality(elem, approx, npq):
    ra = reference_approximation(approx, elem, mapping_degree=approx.p)
    assert (ra.N_p, ra.N_q, ra.N_f) == npq
    assert max(ra.check_sbp_property()) < 1e-13
    assert abs(ra.W.sum() - (2.0 if elem == "Tri" else 4.0 / 3.0)) < 1e-13
    if not ra.V_is_identity:
        assert np.abs(ra.V.T @ (ra.W[:, None] * ra.V) - np.eye(ra.N_p)).max() < 1e-13   # M̂ = I (mass_matrix.jl:59-75)
    S, C = ra.flux_differencing_operators()
    assert max(np.abs(s + s.T).max() for s in S) == 0.0


def test_tet_p4_sparsity_facts():
    ra = reference_approximation(ModalTensor(4), "Tet", mapping_degree=4)
    S, C = ra.flux_differencing_operators()
    assert [int((s != 0).sum()) for s in S] == [500, 1000, 1500]
    assert int((C != 0).sum()) == 1000
    npf = ra.nodes_per_face
    assert [int((C[:, f * npf:(f + 1) * npf] != 0).sum()) for f in range(4)] == [125, 125, 125, 625]
    ra = reference_approximation(ModalTensor(4), "Tri", mapping_degree=4)
    S, C = ra.flux_differencing_operators()
    assert [int((s != 0).sum()) for s in S] == [100, 200] and int((C != 0).sum()) == 75


def test_line_operators():
    ra = reference_approximation(NodalTensor(5), "Line", volume_quadrature_rule=LGQuadrature(5))
    assert max(ra.check_sbp_property()) < 1e-13
    ra = reference_approximation(ModalMulti(4), "Line")
    assert ra.N_q == 5 and max(ra.check_sbp_property()) < 1e-13


@pytest.mark.parametrize("elem,d,M,warp,metric", [
    ("Tri", 2, 2, DelReyWarping(0.1, (1.0, 1.0)), "exact"), ("Tri", 2, 3, ChanWarping(1 / 16, (1.0, 1.0)), "curl"),
    ("Tet", 3, 2, DelReyWarping(0.1, (1.0,) * 3), "curl"), ("Tet", 3, 3, ChanWarping(1 / 16, (1.0,) * 3), "curl"),
    ("Tet", 3, 2, ChanWarping(1 / 16, (1.0,) * 3), "exact")])
def test_curved_mesh_watertight(elem, d, M, warp, metric):
    ra = reference_approximation(ModalTensor(4), elem, mapping_degree=4)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, (M,) * d, warp)
    gf = geometric_factors(mesh, ra, metric)
    assert check_facet_nodes(mesh) < 1e-13
    assert check_normals(mesh, gf) < 1e-11
    assert max(check_sbp_property_physical(ra, gf, 1)) < 1e-13
    assert gf.J_q.min() > 0
    # mapP is an involution without fixed points
    flat = mesh.mapP.reshape(-1)
    assert np.array_equal(flat[flat], np.arange(flat.size)) and not np.any(flat == np.arange(flat.size))


@pytest.mark.parametrize("elem,d", [("Tri", 2), ("Tet", 3)])
def test_structured_connectivity_matches_coordinate_matching(elem, d):
    ra = reference_approximation(ModalTensor(3), elem, mapping_degree=2)
    a = uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, (4,) * d, structured=False)
    assert np.array_equal(a.mapP, structured_connectivity(ra, (4,) * d))
    b = uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, (4,) * d, structured=True)
    assert np.array_equal(a.mapP, b.mapP) and np.abs(a.xyz[0] - b.xyz[0]).max() == 0


def test_affine_curl_metrics_equal_exact():
    ra = reference_approximation(ModalTensor(4), "Tet", mapping_degree=4)
    mesh = uniform_periodic_mesh(ra, ((0.0, 1.0),) * 3, (2,) * 3)
    a, b = geometric_factors(mesh, ra, "exact"), geometric_factors(mesh, ra, "curl")
    assert np.abs(a.Lambda_q - b.Lambda_q).max() < 1e-12 and np.abs(a.nJf - b.nJf).max() < 1e-12


def test_tet_mapping_nodes_are_symmetric_with_lobatto_edges_and_triangle_faces():
    """nodes(Tet(), N) of the mapping element (NodesAndModes, un-vendored): warp-and-blend without alpha optimisation —
    invariant under the 24 vertex permutations, Gauss-Lobatto nodes on every edge, the triangle node set on every face."""
    import itertools
    from sse_b200 import reference as R
    from sse_b200.quadrature import GaussLobattoQuadrature, quadrature_line
    for N in (2, 3, 4, 5):
        r, s, t = R._warp_blend_nodes_tet(N)
        assert r.size == (N + 1) * (N + 2) * (N + 3) // 6
        lam = np.stack([-(1 + r + s + t) / 2, (1 + r) / 2, (1 + s) / 2, (1 + t) / 2], axis=1)
        for perm in itertools.permutations(range(4)):
            for q in lam[:, perm]:
                assert np.min(np.linalg.norm(lam - q, axis=1)) < 1e-12
        gll, _ = quadrature_line(GaussLobattoQuadrature(N))
        edge = lam[(np.abs(lam[:, 2]) < 1e-12) & (np.abs(lam[:, 3]) < 1e-12)]
        assert np.abs(np.sort(2 * edge[:, 1] - 1) - np.sort(gll)).max() < 1e-13
        tri = R._warp_blend_nodes_tri(N)
        tl = np.stack([-(tri[0] + tri[1]) / 2, (1 + tri[0]) / 2, (1 + tri[1]) / 2], axis=1)
        for q in lam[np.abs(lam[:, 3]) < 1e-12][:, :3]:
            assert np.min(np.linalg.norm(tl - q, axis=1)) < 1e-13
