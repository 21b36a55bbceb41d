"""CPU replay of the tree kernels' wavefront schedulers (hare_b200/csrc/oct_wave.cuh, kd_wave.cuh) against the oracle.

The per-slot phase functions and scheduling policies of the Octree and KDTree kernels are __host__ __device__;
tests/emu/oct_emu.cu and kd_emu.cu compile them for the host, build the device arrays with the library's own pack_octree /
pack_kdtree (hare_b200/csrc/pack.hpp) and drive them with a sequential copy of the kernels' trip loops.  Every output must be
bit-identical to the oracle's Octree.Shoot / KDTree.Shoot ("Octree - alt.cs":159-306, KDTree.cs:198-361).  The kernels
themselves are covered by the -m gpu tests.
"""
import shutil

import numpy as np
import pytest

from hare_b200.harness import meshes, rays_from_sources
from oracle import hare_oracle as ho

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="the emulators are compiled with nvcc (host code only)")

FIELDS = ("poly_id", "t", "xyz", "uv")


@pytest.fixture(scope="module")
def hall10k():
    T = ho.Topology.from_mesh(meshes.hall("10k"))
    return T, T.arrays()


@pytest.mark.parametrize("level,args,slots,nmax,warps", [("tiny", (3, 4), 64, 4, 3), ("2k", (5, 8), 48, 4, 1), ("10k", (6, 16), 64, 1, 2),
                                                         ("10k", (8, 4), 32, 4, 5), ("10k", (6, 16), 96, 4, 2), ("2k", (0, 8), 64, 4, 2)])
def test_octree_wave_matches_oracle(level, args, slots, nmax, warps):
    from tests.emu import oct_emu
    T = ho.Topology.from_mesh(meshes.hall(level))
    oc = ho.Octree(T, *args)
    o, d = rays_from_sources(4000, meshes.sources(8), stream=3)
    o[::9] += np.array([90.0, -7.0, 3.0])            # some rays start outside the root cube
    ref0 = oc.Shoot(o, d, nthreads=4)
    o1 = np.full(len(o), -1, np.int32); o1[::3] = ref0["poly_id"][::3]     # poly_origin1 = the polygon the plain Shoot hits
    o2 = np.full(len(o), -1, np.int32); o2[::5] = np.roll(ref0["poly_id"], 1)[::5]
    ref = oc.Shoot(o, d, origin1=o1, origin2=o2, nthreads=4)
    got = oct_emu.run(T.arrays(), oc.arrays(), o, d, origin1=o1, origin2=o2, slots=slots, nmax=nmax, n_warps=warps)
    for k in FIELDS:
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(got["o"], o)                # the Octree never moves a ray
    # the content / chunk / entry culls only ever skip work: the GPU walk enters no more nodes and runs no more exact tests
    assert got["counters"][0] <= ref["counters"][0] and got["counters"][2] <= ref["counters"][2]


def test_octree_wave_irregular_tree_skips_the_shared_child_filter(hall10k):
    """oct_child_filter() derives the eight child intervals from the parent's box; a tree whose child boxes are not the reference's
    function of the parent box (a foreign upload) must not use it.  Same results either way."""
    from tests.emu import oct_emu
    T, ta = hall10k
    oc = ho.Octree(T, 6, 16)
    o, d = rays_from_sources(3000, meshes.sources(8), stream=5)
    ref = oc.Shoot(o, d, nthreads=4)
    a = oct_emu.run(ta, oc.arrays(), o, d, regular=True)
    b = oct_emu.run(ta, oc.arrays(), o, d, regular=False)
    for k in FIELDS:
        assert np.array_equal(a[k], ref[k]) and np.array_equal(b[k], ref[k]), k
    assert a["stats"]["exec"]["N"] < b["stats"]["exec"]["N"]      # the filter saves node steps


def test_octree_wave_chain_matches_oracle(hall10k):
    from tests.emu import oct_emu
    T, ta = hall10k
    oc = ho.Octree(T, 6, 16)
    o, d = rays_from_sources(1200, meshes.sources(4), stream=3)
    ref = oc.reflect_chain(o, d, 10, nthreads=4, points=True)
    got = oct_emu.run(ta, oc.arrays(), o, d, chain=True, order=10, slots=64, nmax=4, n_warps=2)
    for k in ("ev_poly_id", "ev_t", "ev_xyz", "ev_uv", "o", "d", "nshots"):     # ev_xyz / ev_uv: per-bounce X_Point and u, v rows
        assert np.array_equal(got[k], ref[k]), k
    assert int(got["total"][0]) == int(ref["nshots"].sum())


def test_octree_wave_returns_non_closest_hits_like_the_reference(hall10k):
    """Quirk Q11 ("Octree - alt.cs":233-237): children are popped far -> near and a leaf returns as soon as closestT <= its own
    nodeTmin, so the event is not always the closest hit.  The hall rays contain such cases (the KDTree's exhaustive walk gives the
    true closest hit) and the replay reproduces every one of them."""
    from tests.emu import oct_emu
    T, ta = hall10k
    oc = ho.Octree(T, 6, 16)
    o, d = rays_from_sources(20000, meshes.sources(8), stream=3)
    ref = oc.Shoot(o, d, nthreads=8)
    closest = ho.KDTree(T, 18, 16).Shoot(o, d, nthreads=8)
    farther = ref["t"] > closest["t"]
    assert farther.sum() >= 50 and (ref["poly_id"] != closest["poly_id"])[farther].all()
    got = oct_emu.run(ta, oc.arrays(), o[farther], d[farther])
    for k in FIELDS:
        assert np.array_equal(got[k], ref[k][farther]), k


def test_octree_wave_degenerate_rays(hall10k):
    from tests.emu import oct_emu
    T, ta = hall10k
    oc = ho.Octree(T, 6, 16)
    v = ta[0].reshape(-1, 3)
    o = np.array([[15.0, 6.0, 5.0]] * 6 + [v[10], v[500], v[1500], [15.0, 6.0, 0.0], [1e6, 1e6, 1e6], [-1e3, 20.0, 8.0], [15.0, 6.0, 5.0], [15.0, 6.0, 5.0],
                                          [-5.0, 6.0, 5.0], [15.0, 6.0, 5.0]], dtype=np.float64)
    d = np.array([[0, 0, 0], [1e-30, 0, 0], [1e30, 2e30, -1e30], [1e-12, 1e-12, 1], [3, -4, 12], [1e-300, 1e-300, 1e-300],
                  [0.3, 0.4, 0.5], [-0.3, 0.4, 0.5], [0.0, 0.0, 1.0], [0.6, 0.0, 0.8], [-1, -1, -1], [1, 0, 0], [0, -0.0, -1], [1, 1, 0],
                  [-1, 0, 0], [-0.0, 1, 0]], dtype=np.float64)
    ref = oc.Shoot(o, d)
    for slots, nmax in ((64, 4), (32, 4)):
        got = oct_emu.run(ta, oc.arrays(), o, d, slots=slots, nmax=nmax, n_warps=1)
        for k in FIELDS:
            assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("level,args,slots,nmax,warps", [("tiny", (8, 4), 64, 4, 3), ("2k", (14, 8), 48, 4, 1), ("10k", (18, 16), 64, 1, 2), ("2k", (0, 8), 32, 4, 2)])
def test_kdtree_wave_matches_oracle(level, args, slots, nmax, warps):
    from tests.emu import kd_emu
    T = ho.Topology.from_mesh(meshes.hall(level))
    kd = ho.KDTree(T, *args)
    n = 1500
    o, d = rays_from_sources(n, meshes.sources(8), stream=4)
    o[::9] += np.array([90.0, -7.0, 3.0])
    rid = np.arange(1, n + 1, dtype=np.int32); rid[5::11] = 0      # Ray_ID == 0: blind rays (fresh-mailbox case)
    blind = rid == 0
    ref0 = kd.Shoot(o, d, nthreads=8)
    o1 = np.full(n, -1, np.int32); o1[::3] = ref0["poly_id"][::3]
    ref = kd.Shoot(o, d, origin1=o1, ray_id=np.where(blind, 1 << 30, rid).astype(np.int32), nthreads=8)
    got = kd_emu.run(T.arrays(), kd.arrays(), o, d, origin1=o1, ray_id=rid, slots=slots, nmax=nmax, n_warps=warps)
    for k in FIELDS:
        assert np.array_equal(got[k][~blind], ref[k][~blind]), k
    assert (got["poly_id"][blind] == -1).all() and not got["t"][blind].any()
    assert np.array_equal(got["o"], o)


def test_kdtree_wave_chain_matches_oracle():
    from tests.emu import kd_emu
    T = ho.Topology.from_mesh(meshes.hall("2k"))
    kd = ho.KDTree(T, 14, 8)
    o, d = rays_from_sources(600, meshes.sources(4), stream=3)
    ref = kd.reflect_chain(o, d, 8, nthreads=8, points=True)
    got = kd_emu.run(T.arrays(), kd.arrays(), o, d, chain=True, order=8, n_warps=2)
    for k in ("ev_poly_id", "ev_t", "ev_xyz", "ev_uv", "o", "d", "nshots"):     # ev_xyz / ev_uv: per-bounce X_Point and u, v rows
        assert np.array_equal(got[k], ref[k]), k


def lattice_tie_rays():
    """Origins on the 0.5 m lattice of meshes.lattice_room(), 3-4-5 / 45-degree / skew directions with small-integer components:
    every ray meets the floor, the ceiling or a wall exactly on a shared edge or vertex, so two to six polygons are hit at
    bit-identical t."""
    xs = np.arange(1.0, 12.0, 0.5)
    O, D = [], []
    for x in xs:
        for y in xs:
            for dd in ((3, 0, -4), (0, 3, -4), (1, 0, -1), (1, 1, -1), (3, 4, -4), (-1, 0.5, -1), (0.5, 0.5, -4), (3, 0, 4), (1, 1, 1)):
                O.append((x, y, 4.0)); D.append(dd)
    return np.array(O, dtype=np.float64), np.array(D, dtype=np.float64)


@pytest.mark.parametrize("args", [(20, 2), (12, 8)])
def test_kdtree_exact_ties_use_the_reference_node_boxes(args):
    """Among polygons hit at exactly equal t the reference keeps the first one its exhaustive DFS meets; its first/second rule
    (KDTree.cs:249-353) tests the split-plane crossing against the node's own box.  The device nodes hold that box intersected
    with the content box (thin slabs of wall in boxes of air), so the rule must read the untightened side table: with oblique
    rays the crossing point lies inside the node box but outside the content box, `inside` flips and the other polygon wins.
    The what-if replay on the tightened boxes (the round-1 bug) must FAIL this test; the shipped rule must pass it."""
    from tests.emu import kd_emu
    T = ho.Topology.from_mesh(meshes.lattice_room())
    o, d = lattice_tie_rays()
    kd = ho.KDTree(T, *args)
    ref = kd.Shoot(o, d, nthreads=8)
    second = kd.Shoot(o, d, origin1=ref["poly_id"], nthreads=8)
    ties = (ref["poly_id"] >= 0) & (second["poly_id"] >= 0) & (second["t"] == ref["t"])
    assert ties.sum() >= 100
    got = kd_emu.run(T.arrays(), kd.arrays(), o, d)
    for k in FIELDS:
        assert np.array_equal(got[k], ref[k]), k
    bug = kd_emu.run(T.arrays(), kd.arrays(), o, d, tie_rule_on_tight_boxes=True)
    assert (bug["poly_id"] != ref["poly_id"]).sum() >= 20, "the test no longer distinguishes the two box tables"
    assert np.array_equal(bug["t"], ref["t"])          # only WHICH of the tied polygons is reported differs
