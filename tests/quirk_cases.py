"""Constructed inputs for the quirks SURVEY.md 8(a) lists as part of the contract (shared by the CPU and the GPU tests)."""
import numpy as np

from hare_b200.harness.meshes import Mesh, _finish, _box_polys, shoebox


def cube_room(L=8.0):
    """Q4: a CUBIC room, so OBox and VoxelDims are the same numbers on all three axes and a ray with equal direction components
    from a point with equal coordinates has bit-identical tMaxX == tMaxY == tMaxZ at every step."""
    return shoebox(L, L, L)


def dda_tie_rays(L=8.0):
    """Rays whose DDA parameters tie exactly (two-way and three-way), in all sign combinations, from several start points."""
    O, D = [], []
    for c in (0.37, 1.0, 2.5, 4.0, 6.25):
        for d in ((1, 1, 1), (-1, -1, -1), (1, 1, -1), (-1, 1, 1), (1, -1, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (-1, -1, 0), (0, -1, -1),
                  (2, 2, 1), (1, 2, 2), (2, 1, 2), (0.5, 0.5, 0.5), (3, 3, 3)):
            O.append((c, c, c)); D.append(d)
    return np.array(O, dtype=np.float64), np.array(D, dtype=np.float64)


def table_room():
    """Q11: the C1 shoebox plus a small horizontal 'table' quad at z = 0.75 above the floor.  A ray going down through the table
    meets the floor quad (listed in every leaf along the floor, also beyond the point where the ray leaves the room) first in the
    Octree's far -> near order, and returns it early -- although the table is closer."""
    box = shoebox()
    t = np.array([[[5.0, 3.0, 0.75], [6.5, 3.0, 0.75], [6.5, 4.0, 0.75], [5.0, 4.0, 0.75]]], dtype=np.float64)
    verts = np.concatenate([box.verts, t], axis=0)
    vcount = np.concatenate([box.vcount, np.array([4], np.int32)])
    pts = verts.reshape(-1, 3)
    return Mesh(np.ascontiguousarray(verts), vcount, pts.min(axis=0), pts.max(axis=0), "table-room")


def table_rays(n=400, seed=3):
    rng = np.random.default_rng(seed)
    o = np.tile(np.array([[5.0, 3.5, 1.5]]), (n, 1))
    tgt = np.stack([rng.uniform(5.05, 6.45, n), rng.uniform(3.05, 3.95, n), np.full(n, 0.75)], axis=1)   # points on the table
    return o, tgt - o


def sliver_room():
    """Q14: |det| <= 1e-6 is an ABSOLUTE threshold (Hare_Geometry_Polygons.cs:483, 494): det = d . (e1 x e2) = |d| * 2 * area * cos.
    A 2.5 cm x 2 cm triangle (2 * area = 5e-4) floating 1 m in front of the x = 10 wall is hit by a unit-length ray; the same ray
    with its direction scaled by 1e-3 has det = 5e-7: the triangle is invisible and the wall behind it (det = 1e-3 * 21) is hit."""
    box = shoebox()
    t = np.array([[[9.0, 3.5, 1.5], [9.0, 3.525, 1.5], [9.0, 3.5, 1.52], [9.0, 3.5, 1.52]]], dtype=np.float64)
    verts = np.concatenate([box.verts, t], axis=0)
    vcount = np.concatenate([box.vcount, np.array([3], np.int32)])
    pts = verts.reshape(-1, 3)
    return Mesh(np.ascontiguousarray(verts), vcount, pts.min(axis=0), pts.max(axis=0), "sliver-room")


def sliver_rays():
    o = np.array([[5.0, 3.5, 1.5]] * 4)
    tgt = np.array([9.0, 3.505, 1.505])                 # a point inside the small triangle
    u = (tgt - o[0]) / np.linalg.norm(tgt - o[0])        # unit direction: det = |d| * 2 * area * cos = |d| * 5e-4 (cos ~ 1)
    d = np.tile(u, (4, 1)) * np.array([[1.0], [1e-3], [0.25], [1.5e-3]])
    return o, d
