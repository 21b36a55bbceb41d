"""GPU-vs-oracle differential tests: the CUDA path called through
the C ABI against the CPU restatement on the same seeded inputs.  Integer outputs (hit,
Poly_id) must be bit-exact; t, X_Point, u, v are FP64 and are also required to be bit-exact
(north_star allows 1e-9 relative; -fmad=false makes 0 ulp attainable, so that is the bar).
"""
import numpy as np
import pytest

from hare_b200.harness import meshes, rays_from_sources
from hare_b200.harness.meshes import Mesh
from oracle import hare_oracle as ho
from tests.util import assert_events_equal

pytestmark = pytest.mark.gpu

SRC = np.array([[5.0, 3.5, 1.5]])


def _pair(gpu, mesh):
    return gpu.Topology.from_mesh(mesh), ho.Topology.from_mesh(mesh)


# ---------------------------------------------------------------- C1: shoebox, Voxel_Grid 10^3, 100k rays
def test_c1_shoebox_voxelgrid_full(gpu):
    mesh = meshes.shoebox()
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(100_000, SRC, stream=1)
    g = gpu.Voxel_Grid([T], 10)
    ref = ho.Voxel_Grid(To, 10, mode="flat").Shoot(o, d, nthreads=4)
    got = g.Shoot_Batch(o, d, counters=True)
    assert_events_equal(got, ref, what="C1")
    assert got["hit"].all()
    # analytic known answer: nearest of the six planes
    tt = np.full(len(d), np.inf); pid = np.full(len(d), -1)
    for ax, val, idx in [(2, 0.0, 0), (2, 3.0, 1), (0, 0.0, 2), (0, 10.0, 3), (1, 0.0, 4), (1, 7.0, 5)]:
        with np.errstate(divide="ignore", invalid="ignore"):
            tc = (val - o[:, ax]) / d[:, ax]
        ok = (tc > 1e-10) & (tc < tt); tt[ok] = tc[ok]; pid[ok] = idx
    assert np.array_equal(got["poly_id"], pid)
    assert np.allclose(got["t"], tt, rtol=1e-12, atol=0)
    assert int(got["counters"][0]) == int(ref["counters"][0])      # cells visited: same DDA walk
    assert int(got["counters"][3]) == 100_000


@pytest.mark.parametrize("kind,args", [("Octree", (3, 2)), ("KDTree", (4, 1))])
def test_shoebox_trees(gpu, kind, args):
    mesh = meshes.shoebox()
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(50_000, SRC, stream=1)
    got = getattr(gpu, kind)([T], *args).Shoot_Batch(o, d)
    ref = getattr(ho, kind)(To, *args).Shoot(o, d, nthreads=4)
    assert_events_equal(got, ref, what=kind)


# ---------------------------------------------------------------- committed golden vectors (second restatement)
@pytest.mark.parametrize("name", ["shoebox", "tiny"])
def test_golden_vectors(gpu, name):
    """tests/golden/hare_golden.npz was produced by the pure-Python restatement (make_golden.py);
    the CUDA path must reproduce it bit for bit, inputs taken from the fixture itself."""
    import os
    from hare_b200.harness.meshes import Mesh
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hare_golden.npz"))
    mesh = Mesh(G[f"{name}_verts"], G[f"{name}_vcount"], G[f"{name}_minpt"], G[f"{name}_maxpt"], name)
    T = gpu.Topology.from_mesh(mesh)
    a = [int(x) for x in G[f"{name}_args"]]
    o, d = G[f"{name}_o"], G[f"{name}_d"]
    g = gpu.Voxel_Grid([T], a[0])
    off, pol = g.csr()
    assert np.array_equal(off, G[f"{name}_vg_offset"]) and np.array_equal(pol, G[f"{name}_vg_polys"])
    first = G[f"{name}_vg_poly_id"]
    for kind, part, o1, o2, uv in (("vg", g, None, None, False), ("vgo", g, first, np.roll(first, 1), False),
                                   ("oct", gpu.Octree([T], a[1], a[2]), None, None, True), ("kd", gpu.KDTree([T], a[3], a[4]), None, None, True)):
        r = part.Shoot_Batch(o, d, o1, o2, moved=True)
        ref = {k: G[f"{name}_{kind}_{k}"] for k in ("poly_id", "t", "xyz", "uv")}
        assert_events_equal(r, ref, uv=uv, what=f"golden {name} {kind}")
        assert np.array_equal(r["o"], G[f"{name}_{kind}_o_moved"])
        assert (r["poly_id"] == -2).sum() == (ref["poly_id"] == -2).sum()


# ---------------------------------------------------------------- procedural halls
# 100^3: the occupancy bitmap only fits shared memory next to fewer ray pools (12 warps); 128^3: it does not fit at all (L1 path)
@pytest.mark.parametrize("level,domain,nrays", [("tiny", 8, 20_000), ("2k", 16, 50_000), ("10k", 32, 100_000), ("50k", 64, 200_000),
                                                ("50k", 100, 100_000), ("10k", 128, 100_000)])
def test_hall_voxelgrid(gpu, level, domain, nrays):
    mesh = meshes.hall(level)
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(nrays, meshes.sources(4), stream=2)
    g = gpu.Voxel_Grid([T], domain)
    og = ho.Voxel_Grid(To, domain, mode="fast")
    # K2: GPU-built cell lists == oracle lists (exact CSR equality, ascending lists)
    off, pol = g.csr(); ooff, opol = og.csr()
    assert np.array_equal(off, ooff) and np.array_equal(pol, opol)
    ref = og.Shoot(o, d, nthreads=8)
    got = g.Shoot_Batch(o, d)
    assert_events_equal(got, ref, uv=False, what=f"hall-{level} VG{domain}")
    assert (got["uv"] == 0).all()
    assert got["hit"].mean() > 0.99


@pytest.mark.parametrize("level,args,nrays", [("tiny", (3, 4), 20_000), ("2k", (5, 8), 50_000), ("10k", (6, 16), 50_000), ("50k", (7, 32), 100_000)])
def test_hall_octree(gpu, level, args, nrays):
    mesh = meshes.hall(level)
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(nrays, meshes.sources(8), stream=3)
    got = gpu.Octree([T], *args).Shoot_Batch(o, d)
    ref = ho.Octree(To, *args).Shoot(o, d, nthreads=8)
    assert_events_equal(got, ref, what=f"hall-{level} Octree{args}")


@pytest.mark.parametrize("level,args,nrays", [("tiny", (8, 4), 5_000), ("2k", (14, 8), 5_000), ("10k", (18, 16), 2_000)])
def test_hall_kdtree(gpu, level, args, nrays):
    """KDTree.Shoot is exhaustive on the CPU (O(P) per ray); the GPU walk is pruned, and exact-t ties between
    different polygons are resolved in the reference's DFS order (kd_dfs_before), so everything is bit-exact."""
    mesh = meshes.hall(level)
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(nrays, meshes.sources(8), stream=4)
    got = gpu.KDTree([T], *args).Shoot_Batch(o, d)
    ref = ho.KDTree(To, *args).Shoot(o, d, nthreads=8)
    assert_events_equal(got, ref, what=f"hall-{level} KDTree{args}")


def test_kdtree_exact_ties_follow_reference_dfs_order(gpu):
    """Vertical rays through mesh vertices hit up to four polygons at bit-identical t: the reference keeps the
    first one its exhaustive DFS meets."""
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    v = np.unique(T.verts.reshape(-1, 3), axis=0)
    v = v[(v[:, 2] < 3.0)][:400]                      # floor vertices
    o = v + np.array([0.0, 0.0, 0.5]); d = np.tile(np.array([[0.0, 0.0, -1.0]]), (len(v), 1))
    for args in ((12, 8), (6, 40)):
        got = gpu.KDTree([T], *args).Shoot_Batch(o, d)
        ref = ho.KDTree(To, *args).Shoot(o, d)
        assert_events_equal(got, ref, what=f"KDTree{args} vertex rays")
    assert (got["hit"]).mean() > 0.5


# ---------------------------------------------------------------- origins, Ray_ID quirk, outside starts, edge cases
def test_poly_origin_and_rayid(gpu):
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(20_000, meshes.sources(4), stream=5)
    g = gpu.Voxel_Grid([T], 16); og = ho.Voxel_Grid(To, 16, mode="fast")
    first = og.Shoot(o, d)
    o1 = first["poly_id"].copy(); o2 = np.roll(o1, 1)
    rid = np.arange(1, 20_001, dtype=np.int32); rid[::7] = 0      # Q1: Ray_ID == 0 against a fresh mailbox never hits
    blind = rid == 0

    def oracle(part, n):
        # the reference's Ray_ID == 0 behaviour depends on what earlier rays left in the mailbox slot;
        # the batched API defines the fresh-mailbox case, so the blind rays get their own (fresh) oracle call
        ref = part.Shoot(o[:n], d[:n], origin1=o1[:n], origin2=o2[:n], ray_id=np.where(blind[:n], 1 << 30, rid[:n]).astype(np.int32))
        rb = part.Shoot(o[:n][blind[:n]], d[:n][blind[:n]], origin1=o1[:n][blind[:n]], origin2=o2[:n][blind[:n]], ray_id=np.zeros(int(blind[:n].sum()), np.int32))
        for k in ("t", "xyz", "poly_id", "uv"):
            ref[k][blind[:n]] = rb[k]
        return ref
    ref = oracle(og, 20_000)
    got = g.Shoot_Batch(o, d, o1, o2, rid)
    assert_events_equal(got, ref, uv=False, what="origins+rayid")
    assert (got["poly_id"][::7] == -1).all()
    assert (got["poly_id"] != o1).all()
    ref = oracle(ho.Octree(To, 5, 8), 3000)           # the Octree has no mailbox: Ray_ID == 0 is an ordinary ray
    got = gpu.Octree([T], 5, 8).Shoot_Batch(o[:3000], d[:3000], o1[:3000], o2[:3000], rid[:3000])
    assert_events_equal(got, ref, what="octree origins+rayid")
    assert (got["poly_id"][::7] >= 0).any()
    ref = oracle(ho.KDTree(To, 12, 8), 3000)
    got = gpu.KDTree([T], 12, 8).Shoot_Batch(o[:3000], d[:3000], o1[:3000], o2[:3000], rid[:3000])
    assert_events_equal(got, ref, what="kdtree origins+rayid")
    assert (got["poly_id"][::7] == -1).all()


def test_outside_starts_move_the_ray(gpu):
    """Q6: a ray starting outside OBox is moved to its entry point and t includes t_start."""
    mesh = meshes.shoebox()
    T, To = _pair(gpu, mesh)
    n = 20_000
    _, d = rays_from_sources(n, SRC, stream=6)
    o = np.array([5.0, 3.5, 1.5]) - 30.0 * d + 5.0 * np.roll(d, 1, axis=0)   # far outside, aimed roughly at the room
    g = gpu.Voxel_Grid([T], 10); og = ho.Voxel_Grid(To, 10, mode="flat")
    ref = og.Shoot(o, d)
    got = g.Shoot_Batch(o, d, moved=True)
    assert_events_equal(got, ref, uv=False, what="outside")
    assert np.array_equal(got["o"], ref["o"])
    assert (got["o"] != o).any() and got["hit"].any() and (~got["hit"]).any()


def test_negative_zero_direction_and_axis_aligned(gpu):
    """Q7 / H2: -0.0 components, exact axis directions, zero components."""
    mesh = meshes.shoebox()
    T, To = _pair(gpu, mesh)
    d = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                  [1, -0.0, 0], [-0.0, 1, 0], [0.6, 0.8, -0.0], [-0.0, -0.0, 1], [0.6, 0, 0.8], [0, 0.6, -0.8]], dtype=np.float64)
    o = np.repeat(SRC, len(d), axis=0)
    for kind, args, okind, oargs in (("Voxel_Grid", (10,), "Voxel_Grid", (10, "flat")), ("Octree", (3, 2), "Octree", (3, 2)), ("KDTree", (4, 1), "KDTree", (4, 1))):
        got = getattr(gpu, kind)([T], *args).Shoot_Batch(o, d)
        ref = getattr(ho, okind)(To, *oargs).Shoot(o, d)
        assert_events_equal(got, ref, uv=kind != "Voxel_Grid", what=kind + " special directions")


def test_empty_and_single(gpu):
    mesh = meshes.shoebox()
    T = gpu.Topology.from_mesh(mesh)
    g = gpu.Voxel_Grid([T], 10)
    r = g.Shoot_Batch(np.zeros((0, 3)), np.zeros((0, 3)))
    assert r["poly_id"].shape == (0,)
    hit, ev = g.Shoot(gpu.Ray(5, 3.5, 1.5, 0, 0, 1))
    assert hit and ev.Poly_id == 1 and ev.t == 1.5 and ev.X_Point.z == 3.0
    hit, ev = g.Shoot(gpu.Ray(5, 3.5, 1.5, 0, 0, 1, Ray_ID=0))
    assert not hit and ev.Poly_id == -1 and ev.X_Point is None
    hit, ev = g.Shoot(gpu.Ray(5, 3.5, 1.5, 0, 0, 1), 0, 1)       # poly_origin1 = ceiling -> nothing else above
    assert not hit


def test_upload_matches_build(gpu):
    """hare_voxelgrid_upload (host-built lists, here the oracle's hierarchical ctor) and
    hare_octree_upload / hare_kdtree_upload give the same Shoot results as the library's own builds."""
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(20_000, meshes.sources(4), stream=7)
    og = ho.Voxel_Grid(To, 4, mode="hier", avg_polys=0, nthreads=4)      # 2^4 = 16 per axis
    obox, vd, ct, _ = og.info(); off, pol = og.csr()
    g = gpu.Voxel_Grid.from_lists([T], obox, ct, off, pol)
    assert_events_equal(g.Shoot_Batch(o, d), og.Shoot(o, d), uv=False, what="uploaded grid")
    oo = ho.Octree(To, 5, 8)
    t = gpu.Octree.from_nodes([T], *oo.arrays())
    assert_events_equal(t.Shoot_Batch(o, d), oo.Shoot(o, d), what="uploaded octree")
    ko = ho.KDTree(To, 12, 8)
    box, sp, ax, le, ri, lo, lc, pl = ko.arrays()
    assert np.array_equal(ri[le >= 0], le[le >= 0] + 1)
    k = gpu.KDTree.from_nodes([T], box, sp, ax, le, lo, lc, pl)
    assert_events_equal(k.Shoot_Batch(o[:2000], d[:2000]), ko.Shoot(o[:2000], d[:2000]), what="uploaded kd-tree")


# ---------------------------------------------------------------- reflection chains (C2 in small)
@pytest.mark.parametrize("level,domain,nrays,order", [("2k", 16, 4_000, 20), ("10k", 32, 10_000, 50), ("10k", 128, 4_000, 30)])
def test_reflect_chain_voxelgrid(gpu, level, domain, nrays, order):
    mesh = meshes.hall(level)
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(nrays, meshes.sources(4), stream=8)
    got = gpu.Voxel_Grid([T], domain).Reflect_Chain(o, d, order, counters=True)
    ref = ho.Voxel_Grid(To, domain, mode="fast").reflect_chain(o, d, order, nthreads=8)
    assert np.array_equal(got["ev_poly_id"], ref["ev_poly_id"])
    assert np.array_equal(got["ev_t"], ref["ev_t"])
    assert np.array_equal(got["nshots"], ref["nshots"])
    assert np.array_equal(got["o"], ref["o"]) and np.array_equal(got["d"], ref["d"])
    assert got["total_shots"] == int(ref["nshots"].sum())
    assert int(got["counters"][0]) == int(ref["counters"][0])


def test_reflect_chain_octree(gpu):
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(2_000, meshes.sources(4), stream=9)
    got = gpu.Octree([T], 5, 8).Reflect_Chain(o, d, 10)
    ref = ho.Octree(To, 5, 8).reflect_chain(o, d, 10, nthreads=8)
    assert np.array_equal(got["ev_poly_id"], ref["ev_poly_id"]) and np.array_equal(got["ev_t"], ref["ev_t"])


@pytest.mark.parametrize("kind,args,level,nchains", [("Voxel_Grid", (24,), "10k", 120_000), ("Octree", (6, 16), "10k", 40_000), ("KDTree", (14, 8), "2k", 8_000)])
def test_reflect_chain_event_streams_carry_the_whole_x_event(gpu, kind, args, level, nchains):
    """hare_reflect_chain_events: per-bounce X_Point and u, v rows next to Poly_id and t -- bit-equal to the oracle's chain, rows of
    misses and of Shoots that never happened zero, and consistent with each other (the next segment starts at X_Point; t is its
    length along the segment's direction)."""
    mesh = meshes.hall(level)
    T, To = _pair(gpu, mesh)
    order = 12
    o, d = rays_from_sources(nchains, meshes.sources(4), stream=12)     # Voxel_Grid: two pipelined chunks of 87 381 chains
    o[::11] += np.array([60.0, -9.0, 2.0])                       # some chains start outside the model
    part = getattr(gpu, kind)([T], *args)
    ref = (ho.Voxel_Grid(To, args[0], mode="fast") if kind == "Voxel_Grid" else getattr(ho, kind)(To, *args)).reflect_chain(o, d, order, nthreads=8, points=True)
    got = part.Reflect_Chain(o, d, order, points=True)
    for k in ("ev_poly_id", "ev_t", "ev_xyz", "ev_uv", "nshots", "o", "d"):
        assert np.array_equal(got[k], ref[k]), (kind, k)
    hit = got["ev_poly_id"] >= 0
    assert not got["ev_xyz"][~hit].any() and not got["ev_uv"][~hit].any()
    if kind == "Voxel_Grid":
        assert not got["ev_uv"].any()                               # Voxel_Grid.cs:487-488
    else:
        assert got["ev_uv"][hit].any()
    # a chain that made all its Shoots ends where its last X_Point is
    full = hit[:, -1]
    assert full.any() and np.array_equal(got["o"][full], got["ev_xyz"][full, -1])
    # only the streams asked for are touched: the plain call returns the same Poly_id / t
    plain = part.Reflect_Chain(o, d, order)
    assert np.array_equal(plain["ev_poly_id"], got["ev_poly_id"]) and np.array_equal(plain["ev_t"], got["ev_t"])


# ---------------------------------------------------------------- size-independent properties at larger sizes
def test_large_batch_properties(gpu):
    """2M rays on the 50k hall: closed mesh => (almost) every interior ray hits; the hit point lies on the
    reported polygon's plane; results are independent of batch chunking (idempotence across chunk seams)."""
    mesh = meshes.hall("50k")
    T = gpu.Topology.from_mesh(mesh)
    g = gpu.Voxel_Grid([T], 64)
    n = 2_200_000                                   # eight pipelined chunks of 275 k rays on three streams, each through the coherence pre-pass
    o, d = rays_from_sources(n, meshes.sources(4), stream=10)
    r = g.Shoot_Batch(o, d)
    assert r["hit"].mean() > 0.995
    h = r["hit"]
    p = r["poly_id"][h]
    nrm = T.normals[p]; v0 = T.verts[p, 0]
    dist = np.einsum("ij,ij->i", r["xyz"][h] - v0, nrm)
    quad = T.vcount[p] == 4
    assert np.abs(dist[~quad]).max() < 1e-9          # triangles are planar
    assert np.allclose(r["xyz"][h], o[h] + d[h] * r["t"][h][:, None], rtol=0, atol=1e-12)
    lo, hi = (1 << 20) - 1000, (1 << 20) + 1000
    r2 = g.Shoot_Batch(o[lo:hi], d[lo:hi])
    assert np.array_equal(r2["poly_id"], r["poly_id"][lo:hi]) and np.array_equal(r2["t"], r["t"][lo:hi])


def test_adaptive_ctor_matches_hierarchical_reference_ctor(gpu):
    """Voxel_Grid(Model, MaxDomain, Avg_polys) (Voxel_Grid.cs:128-254): same lists, same resolution, same Shoot results
    as the oracle's level-by-level build, including the stop rule (:252)."""
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(20_000, meshes.sources(4), stream=15)
    for maxdom, avg in ((4, 0), (6, 3), (6, 1000)):
        g = gpu.Voxel_Grid([T], maxdom, avg)
        og = ho.Voxel_Grid(To, maxdom, mode="hier", avg_polys=avg, nthreads=4)
        assert np.array_equal(g.info()[2], og.info()[2]), (maxdom, avg)
        a, b = g.csr(), og.csr()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert_events_equal(g.Shoot_Batch(o, d), og.Shoot(o, d), uv=False, what=f"adaptive {maxdom},{avg}")
    X, Y, Z = g.PointInVoxel(gpu.Point(15.0, 6.0, 5.0))
    # the reference's VoxelDecode returns X and Y swapped with respect to VoxelCode (Voxel_Grid.cs:256-267); kept as is
    assert g.VoxelDecode(g.VoxelCode(X, Y, Z)) == (Y, X, Z)


# ---------------------------------------------------------------- in-process multi-device sharding (the C# drop-in's multi-GPU path)
def test_in_process_multi_device_sharding(gpu):
    """hare_init(ids, n): handles are replicated on n devices and hare_shoot_batch block-shards the batch over
    them; results land in the caller's arrays in ray order.  Needs >= 2 GPUs (skipped on a 1-GPU box)."""
    n = gpu.lib().hare_device_count()
    if n < 2:
        pytest.skip("needs at least 2 CUDA devices")
    mesh = meshes.hall("10k")
    o, d = rays_from_sources(300_001, meshes.sources(4), stream=14)
    gpu.init([0])
    T1 = gpu.Topology.from_mesh(mesh)
    ref = gpu.Voxel_Grid([T1], 32).Shoot_Batch(o, d)
    refo = gpu.Octree([T1], 6, 16).Shoot_Batch(o, d)
    refc = gpu.Voxel_Grid([T1], 32).Reflect_Chain(o[:50_000], d[:50_000], 10)
    try:
        gpu.init(list(range(min(n, 8))))
        T = gpu.Topology.from_mesh(mesh)
        got = gpu.Voxel_Grid([T], 32).Shoot_Batch(o, d)
        assert_events_equal(got, ref, uv=False, what="multi-device Voxel_Grid")
        goto = gpu.Octree([T], 6, 16).Shoot_Batch(o, d)
        assert_events_equal(goto, refo, what="multi-device Octree")
        gotc = gpu.Voxel_Grid([T], 32).Reflect_Chain(o[:50_000], d[:50_000], 10)
        assert np.array_equal(gotc["ev_poly_id"], refc["ev_poly_id"]) and np.array_equal(gotc["ev_t"], refc["ev_t"])
        assert gotc["total_shots"] == refc["total_shots"]
    finally:
        gpu.init([0])


def test_degenerate_rays(gpu):
    """Zero, tiny and huge directions, origins on vertices / on walls / far outside: literal IEEE behaviour, same as the oracle."""
    mesh = meshes.hall("2k")
    T, To = _pair(gpu, mesh)
    v = T.verts.reshape(-1, 3)
    o = np.array([[15.0, 6.0, 5.0]] * 6 + [v[10], v[500], v[1500], [15.0, 6.0, 0.0], [1e6, 1e6, 1e6], [-1e3, 20.0, 8.0], [15.0, 6.0, 5.0], [15.0, 6.0, 5.0]], dtype=np.float64)
    d = np.array([[0, 0, 0], [1e-30, 0, 0], [1e30, 2e30, -1e30], [1e-12, 1e-12, 1], [3, -4, 12], [1e-300, 1e-300, 1e-300],
                  [0.3, 0.4, 0.5], [-0.3, 0.4, 0.5], [0.0, 0.0, 1.0], [0.6, 0.0, 0.8], [-1, -1, -1], [1, 0, 0], [0, -0.0, -1], [1, 1, 0]], dtype=np.float64)
    assert o.shape == d.shape
    for kind, args, oargs in (("Voxel_Grid", (16,), (16, "fast")), ("Octree", (5, 8), (5, 8)), ("KDTree", (12, 8), (12, 8))):
        got = getattr(gpu, kind)([T], *args).Shoot_Batch(o, d, moved=True)
        ref = getattr(ho, kind)(To, *oargs).Shoot(o, d)
        assert_events_equal(got, ref, uv=kind != "Voxel_Grid", what=kind + " degenerate rays")
        assert np.array_equal(got["o"], ref["o"]), kind


@pytest.mark.parametrize("level,args", [("shoebox", (3, 2)), ("tiny", (3, 4)), ("2k", (5, 8)), ("10k", (6, 16)), ("50k", (7, 32))])
def test_gpu_octree_build_matches_host_build_and_oracle(gpu, level, args):
    """the Octree is built on the GPU level by level; node boxes, numbering, list order and the
    lost-polygon count equal the host builder's arrays exactly and the oracle's tree structurally."""
    import os
    from tests.util import canon_octree
    mesh = meshes.shoebox() if level == "shoebox" else meshes.hall(level)
    T, To = _pair(gpu, mesh)
    g = gpu.Octree([T], *args)
    os.environ["HARE_OCT_HOST_BUILD"] = "1"
    try:
        h = gpu.Octree([T], *args)
    finally:
        del os.environ["HARE_OCT_HOST_BUILD"]
    for a, b in zip(g.arrays(), h.arrays()):
        assert np.array_equal(a, b)
    assert g.info() == h.info()
    o = ho.Octree(To, *args)
    assert canon_octree(*g.arrays()) == canon_octree(*o.arrays())
    assert (g.info()["nodes"], g.info()["list_entries"], g.info()["lost"]) == o.info()


@pytest.mark.parametrize("level,args", [("shoebox", (4, 1)), ("tiny", (8, 4)), ("2k", (14, 8)), ("10k", (18, 16)), ("50k", (24, 16)), ("50k", (3, 1))])
def test_gpu_kdtree_build_matches_host_build_and_oracle(gpu, level, args):
    """the KDTree is built on the GPU level by level (dense centroid ranks, one stable radix sort
    per level, scan + order-preserving scatter); node boxes, split values, numbering and list order equal the host
    builder's arrays exactly and the oracle's tree structurally."""
    import os
    from tests.util import canon_kdtree
    mesh = meshes.shoebox() if level == "shoebox" else meshes.hall(level)
    T, To = _pair(gpu, mesh)
    g = gpu.KDTree([T], *args)
    os.environ["HARE_KD_HOST_BUILD"] = "1"
    try:
        h = gpu.KDTree([T], *args)
    finally:
        del os.environ["HARE_KD_HOST_BUILD"]
    for a, b in zip(g.arrays(), h.arrays()):
        assert np.array_equal(a, b)
    assert g.info() == h.info()
    box, sp, ax, le, lo, lc, pol = g.arrays()
    o = ho.KDTree(To, *args)
    obox, osp, oax, ole, ori, olo, olc, opol = o.arrays()
    assert canon_kdtree(box, sp, ax, le, le + 1, lo, lc, pol) == canon_kdtree(obox, osp, oax, ole, ori, olo, olc, opol)


@pytest.mark.parametrize("level", ["10k", "50k"])
def test_gpu_topology_ingest_matches_host_ingest_and_oracle(gpu, level):
    """Round(15), the 1 mm first-seen weld, normals and bounds on the GPU (ingest.cu) give the host
    routine's arrays bit for bit -- also when many vertices share lattice cells with earlier ones."""
    import os
    mesh = meshes.hall(level)
    raw = np.array(mesh.verts, dtype=np.float64).reshape(-1, 4, 3)
    rng = np.random.default_rng(7)
    jit = raw + rng.uniform(-4e-4, 4e-4, raw.shape) * (rng.random(raw.shape[:2])[..., None] < 0.3)   # sub-millimetre jitter on 30 % of the corners
    for verts in (raw, jit):
        m2 = Mesh(verts, mesh.vcount, mesh.minpt - 0.01, mesh.maxpt + 0.01, mesh.name)
        Tg = gpu.Topology.from_mesh(m2)
        os.environ["HARE_INGEST_HOST"] = "1"
        try:
            Th = gpu.Topology.from_mesh(m2)
        finally:
            del os.environ["HARE_INGEST_HOST"]
        assert np.array_equal(Tg.verts, Th.verts) and np.array_equal(Tg.normals, Th.normals) and np.array_equal(Tg._mm, Th._mm)
        assert Tg.Vertex_Count == Th.Vertex_Count
        To = ho.Topology.from_mesh(m2)
        v, n, c, mm = To.arrays()
        assert np.array_equal(Tg.verts, v) and np.array_equal(Tg.normals, n) and np.array_equal(Tg._mm, mm) and Tg.Vertex_Count == To.Vertex_Count


def test_gpu_topology_ingest_declines_out_of_bounds_vertices(gpu):
    """A vertex outside the declared Topology bounds is not for the device path: the host routine runs, same result as with it forced."""
    import os
    mesh = meshes.hall("10k")
    m2 = Mesh(mesh.verts, mesh.vcount, mesh.minpt + 1.0, mesh.maxpt + 1.0, mesh.name)
    Tg = gpu.Topology.from_mesh(m2)
    os.environ["HARE_INGEST_HOST"] = "1"
    try:
        Th = gpu.Topology.from_mesh(m2)
    finally:
        del os.environ["HARE_INGEST_HOST"]
    assert np.array_equal(Tg.verts, Th.verts) and np.array_equal(Tg._mm, Th._mm) and Tg.Vertex_Count == Th.Vertex_Count


def test_partition_save_load_shoots_identically(gpu, tmp_path):
    """hare_part_save / hare_part_load: a reloaded Voxel_Grid, Octree and KDTree answer exactly like the one that was built."""
    mesh = meshes.hall("10k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(20_000, meshes.sources(4), stream=5)
    for cls, args in ((gpu.Voxel_Grid, (24,)), (gpu.Octree, (6, 16)), (gpu.KDTree, (18, 16))):
        part = cls([T], *args)
        f = str(tmp_path / (cls.__name__ + ".hare"))
        part.Save(f)
        back = cls.Load([T], f)
        a, b = part.Shoot_Batch(o, d), back.Shoot_Batch(o, d)
        for k in ("poly_id", "t", "xyz", "uv"):
            assert np.array_equal(a[k], b[k]), (cls.__name__, k)
    assert back.Char_Step == 0.0 and gpu.Voxel_Grid.Load([T], str(tmp_path / "Voxel_Grid.hare")).Char_Step > 0


def test_rays_from_far_outside_the_model(gpu):
    """The conservative FP32 culls work in a frame near the model (voxel exit point, root-cube / leaf entry point formed in
    FP64): rays shot from 1e5 m and 1e8 m away give the reference's events bit for bit on every partition."""
    mesh = meshes.hall("10k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(6_000, meshes.sources(4), stream=6)
    for back in (1e5, 1e8):
        of = o - d * back
        for kind, gargs, oargs in (("Voxel_Grid", (24,), (24, "fast")), ("Octree", (6, 16), (6, 16)), ("KDTree", (18, 16), (18, 16))):
            n = 6_000 if kind != "KDTree" else 1_500          # the reference's KDTree walk is exhaustive
            got = getattr(gpu, kind)([T], *gargs).Shoot_Batch(of[:n], d[:n])
            ref = getattr(ho, kind)(To, *oargs).Shoot(of[:n], d[:n], nthreads=8)
            assert_events_equal(got, ref, uv=kind != "Voxel_Grid", what=f"{kind} from {back:g} m")
            assert (ref["poly_id"] >= 0).mean() > 0.9


# ---------------------------------------------------------------- constructed quirk cases (Q4, Q11, Q14; CPU side: tests/test_quirks.py)
def test_q4_dda_tie_rule(gpu):
    """Voxel_Grid.cs:504-550: X only if strictly below both, Y only if below Z, else Z.  Cubic room => bit-identical tMax values."""
    from tests import quirk_cases as qc
    mesh = qc.cube_room()
    T, To = _pair(gpu, mesh)
    o, d = qc.dda_tie_rays()
    og = ho.Voxel_Grid(To, 8, mode="flat")
    ref = og.Shoot(o, d)
    got = gpu.Voxel_Grid([T], 8).Shoot_Batch(o, d, counters=True)
    assert_events_equal(got, ref, uv=False, what="Q4")
    assert int(got["counters"][0]) == int(ref["counters"][0])      # voxels entered: the same path through every tie


def test_q3_epsilon_skirt_of_the_voxels_in_the_gpu_build(gpu):
    """Voxel_Grid.cs:283-285: a polygon 0.4 mm beyond a voxel's nominal face is on that voxel's list, one 1.6 mm beyond it is not --
    in the lists the GPU builds, which equal the oracle's."""
    from tests.test_quirks import q3_mesh
    mesh, obox, vd, ct = q3_mesh()
    T, To = _pair(gpu, mesh)
    ooff, opol = ho.Voxel_Grid(To, 10, mode="flat").csr()
    iy = int((3.3 - obox[1]) / vd[1]); iz = int((1.5 - obox[2]) / vd[2])
    for g in (gpu.Voxel_Grid([T], 10),):
        off, pol = g.csr()
        assert np.array_equal(off, ooff) and np.array_equal(pol, opol)
        listed = lambda ix: set(int(p) for p in pol[off[(ix * int(ct[1]) + iy) * int(ct[2]) + iz]:off[(ix * int(ct[1]) + iy) * int(ct[2]) + iz + 1]])
        assert {6, 7} <= listed(3) and 6 in listed(2) and 7 not in listed(2) and 6 not in listed(1)
    # a ray through the squares from the x = 2 side meets the nearer one first, and the event is the oracle's
    o = np.array([[obox[0] + 2.5 * vd[0], 3.3, 1.5]]); d = np.array([[1.0, 0.0, 0.0]])
    got = gpu.Voxel_Grid([T], 10).Shoot_Batch(o, d); ref = ho.Voxel_Grid(To, 10, mode="flat").Shoot(o, d)
    assert_events_equal(got, ref, uv=False, what="Q3")
    assert int(got["poly_id"][0]) == 6


def test_q11_octree_early_return_is_not_the_closest_hit(gpu):
    """"Octree - alt.cs":233-237: far-first pops + early return.  The table is closer, the Octree reports the floor; and the hall rays
    contain such cases too (octree.poly_id != the KDTree's global closest hit), all reproduced."""
    from tests import quirk_cases as qc
    mesh = qc.table_room()
    T, To = _pair(gpu, mesh)
    o, d = qc.table_rays()
    ref = ho.Octree(To, 4, 1).Shoot(o, d)
    got = gpu.Octree([T], 4, 1).Shoot_Batch(o, d)
    assert_events_equal(got, ref, what="Q11 table")
    kd = gpu.KDTree([T], 6, 1).Shoot_Batch(o, d)
    assert (kd["poly_id"] == 6).all() and (got["poly_id"] == 0).sum() >= 50
    mesh = meshes.hall("10k")
    T, To = _pair(gpu, mesh)
    o, d = rays_from_sources(50_000, meshes.sources(8), stream=3)
    oc = gpu.Octree([T], 6, 16).Shoot_Batch(o, d)
    kd = gpu.KDTree([T], 18, 16).Shoot_Batch(o, d)
    farther = oc["t"] > kd["t"]
    assert farther.sum() >= 100 and (oc["poly_id"] != kd["poly_id"])[farther].all()
    ref = ho.Octree(To, 6, 16).Shoot(o[farther], d[farther], nthreads=8)
    assert_events_equal({k: oc[k][farther] for k in ("poly_id", "t", "xyz", "uv")}, ref, what="Q11 hall")


def test_q14_absolute_determinant_threshold(gpu):
    """Hare_Geometry_Polygons.cs:483, 494: |det| <= 1e-6 is absolute, so a short direction vector makes a small triangle invisible."""
    from tests import quirk_cases as qc
    mesh = qc.sliver_room()
    T, To = _pair(gpu, mesh)
    o, d = qc.sliver_rays()
    for kind, args, oargs in (("Voxel_Grid", (10,), (10, "flat")), ("Octree", (3, 2), (3, 2)), ("KDTree", (4, 1), (4, 1))):
        got = getattr(gpu, kind)([T], *args).Shoot_Batch(o, d)
        ref = getattr(ho, kind)(To, *oargs).Shoot(o, d)
        assert_events_equal(got, ref, uv=kind != "Voxel_Grid", what="Q14 " + kind)
        assert list(got["poly_id"]) == [6, 3, 6, 3]


def test_kdtree_exact_ties_oblique_rays(gpu):
    """Exact-t ties on oblique rays through shared edges / vertices: the tie rule reads the reference's node boxes, not the
    content-tightened device boxes (tests/test_tree_emu.py shows that the tightened ones fail this test)."""
    from tests.test_tree_emu import lattice_tie_rays
    mesh = meshes.lattice_room()
    T, To = _pair(gpu, mesh)
    o, d = lattice_tie_rays()
    for args in ((20, 2), (12, 8)):
        kd = ho.KDTree(To, *args)
        ref = kd.Shoot(o, d, nthreads=8)
        second = kd.Shoot(o, d, origin1=ref["poly_id"], nthreads=8)
        assert ((second["t"] == ref["t"]) & (ref["poly_id"] >= 0) & (second["poly_id"] >= 0)).sum() >= 100
        got = gpu.KDTree([T], *args).Shoot_Batch(o, d)
        assert_events_equal(got, ref, what=f"KDTree{args} oblique ties")


# ---------------------------------------------------------------- the BASELINE sizes (C3 / C4 / C5)
@pytest.fixture(scope="module")
def hall500k(gpu):
    mesh = meshes.hall("500k")
    return mesh, gpu.Topology.from_mesh(mesh), ho.Topology.from_mesh(mesh)


@pytest.fixture(scope="module")
def hall2m(gpu):
    mesh = meshes.hall("2m")
    return mesh, gpu.Topology.from_mesh(mesh), ho.Topology.from_mesh(mesh)


def test_c3_octree_500k(gpu, hall500k):
    """BASELINE config 3 at size: hall-500k, Octree(Model, 7, 32): tree structure and 100 k Shoots bit-equal to the oracle."""
    from tests.util import canon_octree
    mesh, T, To = hall500k
    o, d = rays_from_sources(100_000, meshes.sources(8), stream=3)
    g = gpu.Octree([T], 7, 32); og = ho.Octree(To, 7, 32)
    i = g.info(); n, l, lost = og.info()
    assert (i["nodes"], i["list_entries"], i["lost"]) == (n, l, lost)
    assert canon_octree(*g.arrays()) == canon_octree(*og.arrays())      # same boxes, same children, same list order (numbering-independent form)
    assert_events_equal(g.Shoot_Batch(o, d), og.Shoot(o, d, nthreads=16), what="C3 Octree(7,32)")


def test_c3_voxelgrid_and_kdtree_500k(gpu, hall500k):
    mesh, T, To = hall500k
    o, d = rays_from_sources(100_000, meshes.sources(8), stream=3)
    g = gpu.Voxel_Grid([T], 128); og = ho.Voxel_Grid(To, 128, mode="fast", nthreads=16)
    a, b = g.csr(), og.csr()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert_events_equal(g.Shoot_Batch(o, d), og.Shoot(o, d, nthreads=16), uv=False, what="hall-500k Voxel_Grid 128")
    k = gpu.KDTree([T], 24, 16); ok = ho.KDTree(To, 24, 16)
    assert_events_equal(k.Shoot_Batch(o[:1500], d[:1500]), ok.Shoot(o[:1500], d[:1500], nthreads=16), what="hall-500k KDTree(24,16)")


def test_c4_kdtree_2m(gpu, hall2m):
    """BASELINE config 4 at size: hall-2m, KDTree(Model, 24, 16), 2 k rays against the exhaustive oracle (O(P) per ray)."""
    mesh, T, To = hall2m
    o, d = rays_from_sources(2_000, meshes.sources(8), stream=4)
    k = gpu.KDTree([T], 24, 16); ok = ho.KDTree(To, 24, 16)
    assert k.info()["nodes"] == ok.info()[0] and k.info()["list_entries"] == ok.info()[1]
    assert_events_equal(k.Shoot_Batch(o, d), ok.Shoot(o, d, nthreads=16), what="C4 KDTree(24,16)")


def test_c4_c5_voxelgrid_256_2m(gpu, hall2m):
    """BASELINE configs 4 / 5 at size: 2 M polygons -> 256^3 cell lists, exact CSR equality with the oracle's hierarchical constructor
    Voxel_Grid(Model, MaxDomain = 8, Avg_polys = 0) (Voxel_Grid.cs:128-254), for the flat build and for the library's own
    hierarchical build (8, 0); then 200 k Shoots."""
    mesh, T, To = hall2m
    og = ho.Voxel_Grid(To, 8, mode="hier", avg_polys=0, nthreads=16)
    ooff, opol = og.csr()
    g = gpu.Voxel_Grid([T], 256)
    off, pol = g.csr()
    assert np.array_equal(off, ooff) and np.array_equal(pol, opol)
    ga = gpu.Voxel_Grid([T], 8, 0)
    assert tuple(ga.info()[2]) == (256, 256, 256)
    off, pol = ga.csr()
    assert np.array_equal(off, ooff) and np.array_equal(pol, opol)
    o, d = rays_from_sources(200_000, meshes.sources(8), stream=4)
    assert_events_equal(g.Shoot_Batch(o, d), og.Shoot(o, d, nthreads=16), uv=False, what="C4 Voxel_Grid 256")


def test_results_do_not_depend_on_ray_order_or_batch_cut(gpu):
    """Size-independent property behind the coherence pre-pass (ray_bin.cuh) and the chunked pipeline: a ray's event depends on
    nothing but the ray.  The same 300 k rays in generator order, shuffled, and cut into uneven batches (below and above the
    65 536-ray threshold of the pre-pass) give identical events ray by ray, on all three partitions."""
    mesh = meshes.hall("10k")
    T = gpu.Topology.from_mesh(mesh)
    n = 300_000
    o, d = rays_from_sources(n, meshes.sources(8), stream=21)
    perm = np.random.default_rng(7).permutation(n)
    for part, fields in ((gpu.Voxel_Grid([T], 32), ("poly_id", "t", "xyz")), (gpu.Octree([T], 6, 16), ("poly_id", "t", "xyz", "uv")),
                         (gpu.KDTree([T], 18, 16), ("poly_id", "t", "xyz", "uv"))):
        a = part.Shoot_Batch(o, d)
        b = part.Shoot_Batch(o[perm], d[perm])
        for k in fields:
            assert np.array_equal(a[k][perm], b[k]), (type(part).__name__, k, "shuffled")
        cuts = [0, 1, 1000, 70_000, 70_001, 200_000, n]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            c = part.Shoot_Batch(o[lo:hi], d[lo:hi])
            for k in fields:
                assert np.array_equal(a[k][lo:hi], c[k]), (type(part).__name__, k, lo, hi)
