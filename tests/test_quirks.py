"""The contract quirks that had no constructed test: Q4 DDA tie rule (Voxel_Grid.cs:504-550), Q11 Octree far-first
order + early return ("Octree - alt.cs":233-237), Q14 absolute |det| <= 1e-6 threshold (Hare_Geometry_Polygons.cs:483, 494).
CPU side: the oracle against an explicit expectation and against the second (pure-Python) restatement, and the kernels' phase
functions replayed on the host (tests/emu) against the oracle.  The GPU side is in tests/test_gpu_parity.py."""
import ctypes as C
import itertools
import shutil

import numpy as np
import pytest

from oracle import hare_oracle as ho
from tests import quirk_cases as qc

needs_nvcc = pytest.mark.skipif(shutil.which("nvcc") is None, reason="the emulators are compiled with nvcc (host code only)")


def reference_dda_axis(tx, ty, tz):
    """Voxel_Grid.cs:504-550, literally."""
    if tx < ty:
        return 0 if tx < tz else 2
    return 1 if ty < tz else 2


@needs_nvcc
def test_q4_dda_axis_selection_all_orderings():
    from tests.emu import wave_emu
    L = wave_emu.lib()
    L.emu_dda_axis.argtypes = [C.c_double] * 3
    vals = [1.0, 2.0, 3.0]
    cases = set(itertools.product(vals, repeat=3)) | {(np.inf, 1.0, 1.0), (1.0, np.inf, 1.0), (np.inf, np.inf, 2.0), (np.inf, np.inf, np.inf),
                                                       (-0.0, 0.0, 0.0), (0.0, -0.0, 1.0), (np.nan, 1.0, 2.0), (1.0, np.nan, 0.5), (1.0, 2.0, np.nan)}
    for tx, ty, tz in cases:
        assert L.emu_dda_axis(tx, ty, tz) == reference_dda_axis(tx, ty, tz), (tx, ty, tz)
    # the rule in words: an exact tie goes to the LATER axis
    assert L.emu_dda_axis(1.0, 1.0, 1.0) == 2 and L.emu_dda_axis(1.0, 1.0, 2.0) == 1 and L.emu_dda_axis(1.0, 2.0, 1.0) == 2 and L.emu_dda_axis(2.0, 1.0, 1.0) == 2


def _tie_steps(obox, vd, o, d, n):
    """Count DDA steps of ray (o, d) whose tMax values tie exactly (numpy restatement of the set-up, Voxel_Grid.cs:357-422)."""
    ties = 0
    idx = np.floor((o - obox[:3]) / vd).astype(int)
    tmax = np.full(3, np.inf); tdelta = np.full(3, np.inf); step = np.ones(3, int)
    for a in range(3):
        if d[a] < 0:
            step[a] = -1; tmax[a] = ((idx[a] * vd[a] - 0.001) + obox[a] - o[a]) / d[a]; tdelta[a] = vd[a] / d[a] * -1.0
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                tmax[a] = (((idx[a] + 1) * vd[a] + 0.001) + obox[a] - o[a]) / d[a]; tdelta[a] = vd[a] / d[a] * 1.0
    for _ in range(64):
        if tmax[0] == tmax[1] or tmax[1] == tmax[2] or tmax[0] == tmax[2]:
            ties += 1
        a = reference_dda_axis(*tmax)
        idx[a] += step[a]
        if idx[a] < 0 or idx[a] >= n:
            break
        tmax[a] = tmax[a] + tdelta[a]
    return ties


def test_q4_constructed_dda_ties_oracle_pair():
    """Rays with bit-identical tMax values in a cubic room: the two restatements agree on every event and on the number of voxels
    visited, and the batch really contains tie steps."""
    from oracle import hare_oracle_py as hp
    mesh = qc.cube_room()
    o, d = qc.dda_tie_rays()
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, 8, mode="flat")
    ref = g.Shoot(o, d)
    obox, vd, ct, _ = g.info()
    assert sum(_tie_steps(obox, vd, o[i], d[i], 8) for i in range(len(o))) >= 100
    Tp = hp.Topology(mesh.minpt, mesh.maxpt)
    for i in range(mesh.P):
        Tp.Add_Polygon(mesh.verts[i, :mesh.vcount[i]])
    Tp.Finish_Topology()
    gp = hp.Voxel_Grid([Tp], 8)
    for i in range(len(o)):
        R = hp.Ray(*o[i], *d[i], Ray_ID=i + 1)
        hit, ev = gp.Shoot(R, 0)
        assert hit == (ref["poly_id"][i] >= 0)
        if hit:
            assert ev.Poly_id == ref["poly_id"][i] and ev.t == ref["t"][i]
    assert (ref["poly_id"] >= 0).all()


@needs_nvcc
def test_q4_constructed_dda_ties_kernel_replay():
    from tests.emu import wave_emu
    mesh = qc.cube_room()
    o, d = qc.dda_tie_rays()
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, 8, mode="flat")
    ref = g.Shoot(o, d)
    got = wave_emu.run(To.arrays(), g.info(), g.csr(), o, d, slots=64, wmax=8, n_warps=1)
    for k in ("poly_id", "t", "xyz"):
        assert np.array_equal(got[k], ref[k]), k
    assert int(got["counters"][0]) == int(ref["counters"][0])       # same number of voxels entered: same path through the ties


def test_q11_octree_returns_the_floor_behind_the_table():
    """The table is the closest polygon on every ray (KDTree, Voxel_Grid say so); the Octree's far-first walk returns the floor."""
    mesh = qc.table_room()
    o, d = qc.table_rays()
    To = ho.Topology.from_mesh(mesh)
    kd = ho.KDTree(To, 6, 1).Shoot(o, d)
    vg = ho.Voxel_Grid(To, 10, mode="flat").Shoot(o, d)
    oc = ho.Octree(To, 4, 1).Shoot(o, d)
    assert (kd["poly_id"] == 6).all() and (vg["poly_id"] == 6).all()
    assert (oc["poly_id"] == 0).sum() >= 50, "the constructed case no longer triggers the early return"
    floor = oc["poly_id"] == 0
    assert (oc["t"][floor] > kd["t"][floor]).all() and np.allclose(oc["xyz"][floor][:, 2], 0.0)


@needs_nvcc
def test_q11_kernel_replay():
    from tests.emu import oct_emu
    mesh = qc.table_room()
    o, d = qc.table_rays()
    To = ho.Topology.from_mesh(mesh)
    oc = ho.Octree(To, 4, 1)
    ref = oc.Shoot(o, d)
    got = oct_emu.run(To.arrays(), oc.arrays(), o, d, n_warps=1)
    for k in ("poly_id", "t", "xyz", "uv"):
        assert np.array_equal(got[k], ref[k]), k


def test_q14_absolute_determinant_threshold():
    mesh = qc.sliver_room()
    o, d = qc.sliver_rays()
    To = ho.Topology.from_mesh(mesh)
    for part in (ho.Voxel_Grid(To, 10, mode="flat"), ho.Octree(To, 3, 2), ho.KDTree(To, 4, 1)):
        r = part.Shoot(o, d)
        # |d| = 1 and 0.25: 2*area*|d| = 5e-4, 1.25e-4 > 1e-6 -> the triangle (index 6) is hit; |d| = 1e-3, 1.5e-3: 5e-7, 7.5e-7 <= 1e-6 -> invisible, the wall x = 10 (index 3)
        assert list(r["poly_id"]) == [6, 3, 6, 3], type(part).__name__
        assert np.allclose(r["xyz"][[0, 2], 0], 9.0) and np.allclose(r["xyz"][[1, 3], 0], 10.0)


@needs_nvcc
def test_q14_kernel_replays():
    from tests.emu import kd_emu, oct_emu, wave_emu
    mesh = qc.sliver_room()
    o, d = qc.sliver_rays()
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, 10, mode="flat"); oc = ho.Octree(To, 3, 2); kd = ho.KDTree(To, 4, 1)
    for got, ref in ((wave_emu.run(To.arrays(), g.info(), g.csr(), o, d, n_warps=1), g.Shoot(o, d)),
                     (oct_emu.run(To.arrays(), oc.arrays(), o, d, n_warps=1), oc.Shoot(o, d)),
                     (kd_emu.run(To.arrays(), kd.arrays(), o, d, n_warps=1), kd.Shoot(o, d))):
        for k in ("poly_id", "t", "xyz"):
            assert np.array_equal(got[k], ref[k]), k


# ---------------------------------------------------------------- Q2, Q3: the two quirks that only had implicit coverage
def _shoebox_with(extra_quads):
    from hare_b200.harness import meshes
    from hare_b200.harness.meshes import Mesh
    box = meshes.shoebox()
    verts = np.concatenate([box.verts, np.asarray(extra_quads, np.float64).reshape(-1, 4, 3)])
    vcount = np.concatenate([box.vcount, np.full(len(extra_quads), 4, box.vcount.dtype)])
    return Mesh(verts, vcount, box.minpt, box.maxpt, "shoebox+")


def q3_mesh():
    """The C1 shoebox plus two 20 cm squares parallel to the nominal face between voxels x = 2 and x = 3 of its 10^3 grid, 0.4 mm and
    1.6 mm beyond it (polygons 6 and 7).  Returns (mesh, obox, voxel dims, counts) of that grid."""
    from hare_b200.harness import meshes
    g0 = ho.Voxel_Grid(ho.Topology.from_mesh(meshes.shoebox()), 10, mode="flat")
    obox, vd, ct, _ = g0.info()
    face = obox[0] + 3 * vd[0]
    quads = []
    for dx in (0.0004, 0.0016):
        x = face + dx
        quads.append([[x, 3.2, 1.4], [x, 3.4, 1.4], [x, 3.4, 1.6], [x, 3.2, 1.6]])
    return _shoebox_with(quads), obox, vd, ct


def test_q3_voxels_are_inflated_by_epsilon_on_every_face():
    """Voxel_Grid.cs:283-285: a voxel's box is [(x*vd - eps) + omin, ((x+1)*vd + eps) + omin] with eps = 0.001 -- a polygon lying 0.4 mm
    beyond a voxel's nominal face is still on that voxel's list, one lying 1.6 mm beyond it is not.  Checked on the C++ oracle, on the
    pure-Python restatement and on the product's host-side expectation of the same lists (the GPU build is compared with these lists on
    the halls by tests/test_gpu_parity.py)."""
    from oracle import hare_oracle_py as hp
    mesh, obox, vd, ct = q3_mesh()
    To = ho.Topology.from_mesh(mesh)
    g = ho.Voxel_Grid(To, 10, mode="flat")
    obox2, vd2, ct2, _ = g.info()
    assert np.array_equal(obox, obox2) and np.array_equal(vd, vd2)    # the squares lie inside the room: same grid
    off, pol = g.csr()
    iy = int((3.3 - obox[1]) / vd[1]); iz = int((1.5 - obox[2]) / vd[2])

    def listed(ix):
        c = (ix * int(ct[1]) + iy) * int(ct[2]) + iz
        return set(int(p) for p in pol[off[c]:off[c + 1]])
    near, far = 6, 7                                              # the two added quads (after the shoebox's six)
    assert near in listed(3) and far in listed(3)                 # both lie in voxel 3
    assert near in listed(2) and far not in listed(2)             # only the one inside the epsilon skirt is ALSO in voxel 2
    assert near not in listed(1) and near not in listed(4)
    # the fast (polygon-major) oracle build and the Python restatement agree
    off2, pol2 = ho.Voxel_Grid(To, 10, mode="fast").csr()
    assert np.array_equal(off, off2) and np.array_equal(pol, pol2)
    Tp = hp.Topology(tuple(mesh.minpt), tuple(mesh.maxpt))
    for i in range(mesh.P):
        Tp.Add_Polygon([tuple(map(float, mesh.verts[i, k])) for k in range(mesh.vcount[i])])
    Tp.Finish_Topology()
    gp = hp.Voxel_Grid([Tp], 10)
    for ix in (1, 2, 3, 4):
        assert set(gp.Voxel_Inv[(ix, iy, iz)]) == listed(ix)


def test_q2_mailbox_never_changes_a_result():
    """Voxel_Grid.cs:478-480 / KDTree.cs:224-229: with unique non-zero Ray_IDs the mailbox only suppresses a second test of a polygon
    by the SAME ray, which the strict `t < tmin` would reject anyway.  The oracle keeps the reference's mailbox; its results do not
    depend on which rays went through a mailbox before (1 thread = one mailbox history, 8 threads = eight others, or ids in reverse),
    and the kernels, which keep no mailbox at all, are compared with it bit for bit everywhere else."""
    from hare_b200.harness import meshes, rays_from_sources
    To = ho.Topology.from_mesh(meshes.hall("2k"))
    o, d = rays_from_sources(6000, meshes.sources(4), stream=17)
    ids = np.arange(1, len(o) + 1, dtype=np.int32)
    for part in (ho.Voxel_Grid(To, 12, mode="fast"), ho.KDTree(To, 8, 8)):
        a = part.Shoot(o, d, ray_id=ids, nthreads=1)
        b = part.Shoot(o, d, ray_id=ids, nthreads=8)
        c = part.Shoot(o, d, ray_id=ids[::-1].copy(), nthreads=3)
        for k in ("poly_id", "t", "xyz", "uv"):
            assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k
        assert (a["poly_id"] >= 0).mean() > 0.9
