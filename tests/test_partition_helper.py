"""sse_partition_* (host-only C ABI): the local view of one rank of an element partition from the global mapP and an owner
array.  Checked (a) against the structured slab partitioner of the Python mirror (same conventions -> identical arrays), and
(b) on a random, non-slab owner assignment through the defining property: gathering neighbour values through the local mapP,
after a simulated halo exchange driven by the send lists, equals the gather through the global mapP."""
import numpy as np
import pytest

from sse_b200.dist import partition
from sse_b200.mesh import ChanWarping, uniform_periodic_mesh
from sse_b200.reference import ModalTensor, reference_approximation


def _meshes(elem, d, M, world):
    ra = reference_approximation(ModalTensor(2), elem, mapping_degree=2)
    warp = ChanWarping(1 / 16, (1.0,) * d)
    full = uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, (M,) * d, warp, part=(0, 1))
    parts = [uniform_periodic_mesh(ra, ((0.0, 1.0),) * d, (M,) * d, warp, part=(r, world)) for r in range(world)]
    return ra, full, parts


@pytest.mark.parametrize("elem,d,M,world", [("Tri", 2, 4, 2), ("Tet", 3, 4, 2), ("Tet", 3, 4, 4), ("Tri", 2, 6, 3)])
def test_matches_the_structured_slab_partitioner(elem, d, M, world):
    ra, full, parts = _meshes(elem, d, M, world)
    nsimp = 2 if d == 2 else 6
    owner = (np.arange(full.N_e) // nsimp // (M ** (d - 1))) // (M // world)
    assert np.array_equal(full.elem_gid, np.arange(full.N_e))
    for r, pm in enumerate(parts):
        got = partition(full.mapP.reshape(-1) + 1, ra.N_f, owner, world, r)
        assert np.array_equal(got["elem_gid"], pm.elem_gid)
        assert got["n_interior"] == pm.N_e - pm.n_boundary and got["n_ghost"] == pm.n_ghost
        assert got["nbr_ranks"] == pm.nbr_ranks
        assert got["send_count"] == [int(s.size) for s in pm.send_idx]
        assert np.array_equal(got["send_idx"], np.concatenate(pm.send_idx))
        assert np.array_equal(got["mapP"], pm.mapP)


def test_random_owner_assignment_reproduces_the_global_gather():
    ra, full, _ = _meshes("Tet", 3, 3, 1)
    rng = np.random.default_rng(5)
    nparts = 3
    owner = rng.integers(0, nparts, size=full.N_e).astype(np.int32)
    nf = ra.N_f
    val = rng.standard_normal(full.N_e * nf)                         # one number per global facet node
    want = val[full.mapP.reshape(-1)].reshape(full.N_e, nf)          # what every element gathers from its neighbours
    views = [partition(full.mapP.reshape(-1) + 1, nf, owner, nparts, r) for r in range(nparts)]
    for r, v in enumerate(views):
        gid, nl = v["elem_gid"], v["elem_gid"].size
        assert np.all(owner[gid] == r) and nl == int((owner == r).sum())
        owned = val.reshape(full.N_e, nf)[gid].reshape(-1)
        ghost = np.full(v["n_ghost"], np.nan)
        off = 0
        for nb, rc in zip(v["nbr_ranks"], v["recv_count"]):
            w = views[nb]
            i = w["nbr_ranks"].index(r)                              # the neighbour's segment for us
            so = int(np.sum(w["send_count"][:i]))
            seg = w["send_idx"][so:so + w["send_count"][i]]
            assert seg.size == rc
            ghost[off:off + rc] = val.reshape(full.N_e, nf)[w["elem_gid"]].reshape(-1)[seg]
            off += rc
        facet = np.concatenate([owned, ghost])
        got = facet[v["mapP"].reshape(-1)].reshape(nl, nf)
        assert np.array_equal(got, want[gid])
        # interior elements read no ghost slot, halo-adjacent ones do
        assert np.all(v["mapP"][:v["n_interior"]] < nl * nf)
        assert np.all(np.any(v["mapP"][v["n_interior"]:] >= nl * nf, axis=1))


def test_bad_arguments_are_refused():
    from sse_b200._lib import SSEError
    ra, full, _ = _meshes("Tri", 2, 3, 1)
    mp = full.mapP.reshape(-1) + 1
    with pytest.raises(SSEError):
        partition(mp, ra.N_f, np.zeros(full.N_e, dtype=np.int32), 2, 1)          # rank 1 owns nothing
    bad = mp.copy()
    bad[3] = 0
    with pytest.raises(SSEError):
        partition(bad, ra.N_f, np.zeros(full.N_e, dtype=np.int32), 1, 0)         # BoundsError
