"""Pins for the CPU oracle (oracle/hare_oracle.cpp).  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned by
  1. analytic known-answer tests on the C1 shoebox,
  2. committed golden vectors produced by the independent pure-Python restatement
     (oracle/hare_oracle_py.py via tests/golden/make_golden.py) -- bit-exact agreement of two
     restatements written separately from the C# sources,
  3. quirk tests for the reference behaviours listed in SURVEY.md 8(a) Q1-Q14.
"""
import os

import numpy as np
import pytest

from hare_b200.harness import meshes, rays_from_sources
from hare_b200.harness.meshes import Mesh
from oracle import hare_oracle as ho
from oracle import hare_oracle_py as hp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hare_golden.npz"))
SRC = np.array([[5.0, 3.5, 1.5]])


def golden_mesh(name):
    return Mesh(G[f"{name}_verts"], G[f"{name}_vcount"], G[f"{name}_minpt"], G[f"{name}_maxpt"], name)


def analytic_shoebox(o, d):
    tt = np.full(len(d), np.inf); pid = np.full(len(d), -1)
    for ax, val, idx in [(2, 0.0, 0), (2, 3.0, 1), (0, 0.0, 2), (0, 10.0, 3), (1, 0.0, 4), (1, 7.0, 5)]:
        with np.errstate(divide="ignore", invalid="ignore"):
            tc = (val - o[:, ax]) / d[:, ax]
        ok = (tc > 1e-10) & (tc < tt); tt[ok] = tc[ok]; pid[ok] = idx
    return pid, tt


# ---------------------------------------------------------------- 1. analytic KATs
@pytest.mark.parametrize("kind,args", [("Voxel_Grid", (10, "flat")), ("Voxel_Grid", (10, "fast")), ("Octree", (3, 2)), ("KDTree", (4, 1))])
def test_kat_shoebox(kind, args):
    T = ho.Topology.from_mesh(meshes.shoebox())
    o, d = rays_from_sources(20_000, SRC, stream=1)
    r = getattr(ho, kind)(T, *args).Shoot(o, d, nthreads=2)
    pid, tt = analytic_shoebox(o, d)
    assert np.array_equal(r["poly_id"], pid)
    assert np.allclose(r["t"], tt, rtol=1e-12, atol=0)
    assert np.allclose(r["xyz"], o + d * tt[:, None], rtol=0, atol=1e-12)


def test_kat_axis_rays_exact():
    T = ho.Topology.from_mesh(meshes.shoebox())
    g = ho.Voxel_Grid(T, 10, "flat")
    d = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], dtype=float)
    r = g.Shoot(np.repeat(SRC, 6, 0), d)
    assert r["poly_id"].tolist() == [3, 2, 5, 4, 1, 0]
    assert r["t"].tolist() == [5.0, 5.0, 3.5, 3.5, 1.5, 1.5]


def test_normals_and_bounds():
    T = ho.Topology.from_mesh(meshes.shoebox())
    v, n, c, mm = T.arrays()
    assert np.array_equal(np.abs(n), np.abs(np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0]], float)))
    assert np.allclose(mm, [-1e-12] * 3 + [10 + 1e-12, 7 + 1e-12, 3 + 1e-12], rtol=0, atol=1e-15)
    assert T.Vertex_Count == 8


# ---------------------------------------------------------------- 2. golden vectors (second restatement)
@pytest.mark.parametrize("name", ["shoebox", "tiny"])
def test_golden_topology(name):
    T = ho.Topology.from_mesh(golden_mesh(name))
    v, n, c, mm = T.arrays()
    assert np.array_equal(v, G[f"{name}_topo_verts"])
    assert np.array_equal(n, G[f"{name}_topo_normals"])
    assert np.array_equal(mm, G[f"{name}_topo_minmax"])


@pytest.mark.parametrize("name", ["shoebox", "tiny"])
def test_golden_voxelgrid_lists(name):
    T = ho.Topology.from_mesh(golden_mesh(name))
    dom = int(G[f"{name}_args"][0])
    for mode in ("flat", "fast"):
        off, pol = ho.Voxel_Grid(T, dom, mode).csr()
        assert np.array_equal(off, G[f"{name}_vg_offset"]) and np.array_equal(pol, G[f"{name}_vg_polys"]), mode


@pytest.mark.parametrize("name", ["shoebox", "tiny"])
@pytest.mark.parametrize("kind", ["vg", "vgo", "oct", "kd"])
def test_golden_shoot(name, kind):
    T = ho.Topology.from_mesh(golden_mesh(name))
    a = [int(x) for x in G[f"{name}_args"]]
    o, d = G[f"{name}_o"], G[f"{name}_d"]
    if kind in ("vg", "vgo"):
        part = ho.Voxel_Grid(T, a[0], "flat")
    elif kind == "oct":
        part = ho.Octree(T, a[1], a[2])
    else:
        part = ho.KDTree(T, a[3], a[4])
    o1 = o2 = None
    if kind == "vgo":
        o1 = G[f"{name}_vg_poly_id"].copy(); o2 = np.roll(o1, 1)
    r = part.Shoot(o, d, origin1=o1, origin2=o2)
    for k in ("poly_id", "t", "xyz", "uv"):
        assert np.array_equal(r[k], G[f"{name}_{kind}_{k}"]), (name, kind, k)
    assert np.array_equal(r["o"], G[f"{name}_{kind}_o_moved"])


def test_poly_box_overlap_matches_python_restatement():
    rng = np.random.default_rng(7)
    agree = 0
    for _ in range(3000):
        c = rng.uniform(-1, 1, 3); h = rng.uniform(0.05, 0.6, 3)
        n = int(rng.integers(3, 5))
        P = rng.uniform(-1.5, 1.5, (n, 3))
        a = ho.poly_box_overlap(c - h, c + h, P)
        b = hp.AABB(tuple(c - h), tuple(c + h)).PolyBoxOverlap([tuple(map(float, p)) for p in P])
        assert a == b
        agree += a
    assert 300 < agree < 2700      # both outcomes are exercised


def test_round15_matches_dotnet_rule():
    for x in [0.1, 1 / 3, 2.5e-16, 40 * 0.123456789012345678, -7.00000000000000051, 1e16, 123456.7890123456789]:
        assert ho.round15(x) == hp.round15(x)
    assert ho.round15(0.1234567890123456) == 0.123456789012346
    assert ho.round15(1e16 + 2.0) == 1e16 + 2.0


# ---------------------------------------------------------------- 3. quirks
def test_q1_rayid_zero_against_fresh_mailbox_never_hits():
    T = ho.Topology.from_mesh(meshes.shoebox())
    o, d = rays_from_sources(64, SRC, stream=1)
    zero = np.zeros(64, np.int32)
    assert (ho.Voxel_Grid(T, 10, "flat").Shoot(o, d, ray_id=zero)["poly_id"] == -1).all()
    assert (ho.KDTree(T, 4, 1).Shoot(o, d, ray_id=zero)["poly_id"] == -1).all()
    assert (ho.Octree(T, 3, 2).Shoot(o, d, ray_id=zero)["poly_id"] >= 0).all()     # mailbox commented out there


def test_q5_leaving_grid_is_a_miss_and_q6_outside_start_moves_ray():
    T = ho.Topology.from_mesh(meshes.shoebox())
    g = ho.Voxel_Grid(T, 10, "flat")
    o = np.array([[-20.0, 3.5, 1.5], [-20.0, 3.5, 1.5], [5.0, 3.5, 50.0]]); d = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, -1.0]])
    r = g.Shoot(o, d)
    assert r["poly_id"].tolist() == [2, -1, 1]
    assert r["t"][0] == 20.0 and r["t"][2] == 47.0                     # t includes t_start
    assert abs(r["o"][0, 0] + 0.101) < 1e-9 and abs(r["o"][2, 2] - 3.101) < 1e-9     # origin moved to the OBox face
    assert np.array_equal(r["o"][1], o[1])                               # missed OBox: untouched


def test_q7_negative_zero_direction():
    T = ho.Topology.from_mesh(meshes.shoebox())
    g = ho.Voxel_Grid(T, 10, "flat")
    r = g.Shoot(np.repeat(SRC, 2, 0), np.array([[0.0, 0.0, 1.0], [-0.0, 0.0, 1.0]]))
    assert r["poly_id"][0] == 1
    # dx == -0.0 makes tMaxX = -inf: the walk marches along +X and leaves the grid, or accepts on the way
    py = hp.Voxel_Grid([_py_topo(meshes.shoebox())], 10).Shoot(hp.Ray(5.0, 3.5, 1.5, -0.0, 0.0, 1.0))
    assert r["poly_id"][1] == (py[1].Poly_id if py[0] else -1)


def _py_topo(mesh):
    T = hp.Topology(tuple(mesh.minpt), tuple(mesh.maxpt))
    for i in range(mesh.P):
        T.Add_Polygon([tuple(map(float, mesh.verts[i, k])) for k in range(mesh.vcount[i])])
    T.Finish_Topology()
    return T


def test_q12_kdtree_is_global_closest_hit_and_voxelgrid_agrees_inside_closed_mesh():
    mesh = meshes.hall("tiny")
    T = ho.Topology.from_mesh(mesh)
    o, d = rays_from_sources(3000, meshes.sources(4), stream=12)
    kd = ho.KDTree(T, 6, 6).Shoot(o, d)
    brute = ho.KDTree(T, 0, 1 << 30).Shoot(o, d)          # a single leaf: plain loop over every polygon
    assert np.array_equal(kd["t"], brute["t"])
    vg = ho.Voxel_Grid(T, 6, "fast").Shoot(o, d)
    both = (vg["poly_id"] >= 0) & (kd["poly_id"] >= 0)
    assert both.mean() > 0.95
    assert np.array_equal(vg["t"][both], kd["t"][both])


def test_q10_octree_root_box_expression():
    T = ho.Topology.from_mesh(meshes.shoebox())
    box, fc, lo, lc, pol = ho.Octree(T, 3, 2).arrays()
    # center = max + min/2 (sic), half-extent = maxdim + 0.1
    assert np.allclose(box[0], [10 - 10 - 0.1, 7 - 10 - 0.1, 3 - 10 - 0.1, 10 + 10 + 0.1, 7 + 10 + 0.1, 3 + 10 + 0.1], rtol=0, atol=1e-12)
    assert fc[0] == 1 and (box[1:9, 3:] - box[1:9, :3] > 10.1).all()   # children padded by 0.1 on every side


def test_hier_ctor_matches_flat_on_small_case():
    T = ho.Topology.from_mesh(meshes.hall("tiny"))
    a = ho.Voxel_Grid(T, 8, "flat", nthreads=2).csr()
    b = ho.Voxel_Grid(T, 3, "hier", avg_polys=0, nthreads=2).csr()       # 2^3 = 8 per axis
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_more_than_four_sides_is_rejected():
    T = ho.Topology([0, 0, 0], [1, 1, 1])
    with pytest.raises(NotImplementedError):
        T.Add_Polygon(np.zeros((5, 3)))


# ---------------------------------------------------------------- harness chains: the C++ oracle against the Python restatement
@pytest.mark.parametrize("kind,args", [("Voxel_Grid", (6,)), ("Octree", (3, 4)), ("KDTree", (5, 4))])
def test_chain_event_rows_match_python_restatement(kind, args):
    """The specular chain is harness-defined: n = Normal[poly], k = 2*((dx*nx)+(dy*ny)+(dz*nz)), d' = d - k*n,
    o' = X_Point, poly_origin1 = poly.  Run it bounce by bounce on the independent pure-Python Shoots and compare every row the C++
    oracle's reflect_chain(points=True) reports: Poly_id, t, X_Point, u, v, the final ray and the number of Shoots."""
    mesh = meshes.hall("tiny")
    To = ho.Topology.from_mesh(mesh)
    Tp = _py_topo(mesh)
    order, n = 6, 60
    o, d = rays_from_sources(n, meshes.sources(4), stream=21)
    o[::7] += np.array([40.0, -5.0, 1.0])                         # some chains start outside the model
    cpp = (ho.Voxel_Grid(To, args[0], mode="flat") if kind == "Voxel_Grid" else getattr(ho, kind)(To, *args)).reflect_chain(o, d, order, points=True)
    part = getattr(hp, kind)([Tp], *args)
    hits = 0
    for i in range(n):
        R = hp.Ray(*map(float, o[i]), *map(float, d[i]), Ray_ID=1)
        o1, b = -1, 0
        while b < order:
            R.Ray_ID = (i * order + b) % 2147483646 + 1
            hit, ev = part.Shoot(R, 0, o1)
            row = (i, b)
            assert cpp["ev_poly_id"][row] == ev.Poly_id and cpp["ev_t"][row] == ev.t, (kind, row)
            X = ev.X_Point if hit else (0.0, 0.0, 0.0)
            assert tuple(cpp["ev_xyz"][row]) == tuple(X) and tuple(cpp["ev_uv"][row]) == (ev.u, ev.v), (kind, row)
            b += 1
            if not hit:
                break
            hits += 1
            N = Tp.Polys[ev.Poly_id][1]
            k = 2 * ((R.dx * N[0]) + (R.dy * N[1]) + (R.dz * N[2]))
            R.dx, R.dy, R.dz = R.dx - k * N[0], R.dy - k * N[1], R.dz - k * N[2]
            R.x, R.y, R.z = X
            o1 = ev.Poly_id
        assert cpp["nshots"][i] == b
        assert not cpp["ev_xyz"][i, b:].any() and not cpp["ev_uv"][i, b:].any() and np.all(cpp["ev_poly_id"][i, b:] == -3)
        assert tuple(cpp["o"][i]) == (R.x, R.y, R.z) and tuple(cpp["d"][i]) == (R.dx, R.dy, R.dz)
    assert hits > n * 2
