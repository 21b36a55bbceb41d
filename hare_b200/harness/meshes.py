"""Procedural meshes for BASELINE.json's configs.

shoebox()          C1: 6 quads, [0,10]x[0,7]x[0,3].
hall("50k")        C2: auditorium, ~50k tri+quad polygons.
hall("500k")       C3: concert hall, ~500k polygons.
hall("2m")         C4/C5: strongly non-uniform hall (seat blocks), ~2M polygons.
hall("tiny"/"2k")  small members of the same family for CPU-sized tests.

Every mesh has its min corner at the origin (the Octree root-box quirk,
"Octree - alt.cs":79, only contains such meshes), closed outer shell, vertices
more than 2 mm apart (Topology welds inside 1 mm cells,
Hare_Geometry_Topology.cs:342-377) and coordinates passed through the
Math.Round(x, 15) emulation (Hare_Geometry_Primitives.cs:230-235).  Only
+ - * / floor abs are used, all IEEE-exact.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Mesh:
    verts: np.ndarray    # (P, 4, 3) float64; triangles repeat vertex 2 in slot 3
    vcount: np.ndarray   # (P,) int32, 3 or 4
    minpt: np.ndarray    # (3,)
    maxpt: np.ndarray    # (3,)
    name: str = ""

    @property
    def P(self):
        return int(self.vcount.shape[0])


def round15(x):
    """.NET Math.Round(x, 15) for |x| < 1e16 (x*1e15 -> half-even -> /1e15)."""
    return np.rint(x * 1e15) / 1e15


def _finish(polys, name):
    verts = np.concatenate([p[0] for p in polys], axis=0)
    vcount = np.concatenate([p[1] for p in polys], axis=0).astype(np.int32)
    verts = round15(round15(verts))
    pts = verts.reshape(-1, 3)
    return Mesh(np.ascontiguousarray(verts), vcount, pts.min(axis=0), pts.max(axis=0), name)


def shoebox(lx=10.0, ly=7.0, lz=3.0):
    """C1 shoebox: floor, ceiling, x=0, x=lx, y=0, y=ly; CCW seen from inside."""
    q = np.array([
        [[0, 0, 0], [lx, 0, 0], [lx, ly, 0], [0, ly, 0]],          # floor (normal +z)
        [[0, 0, lz], [0, ly, lz], [lx, ly, lz], [lx, 0, lz]],      # ceiling (normal -z)
        [[0, 0, 0], [0, ly, 0], [0, ly, lz], [0, 0, lz]],          # x = 0 (normal +x)
        [[lx, 0, 0], [lx, 0, lz], [lx, ly, lz], [lx, ly, 0]],      # x = lx (normal -x)
        [[0, 0, 0], [0, 0, lz], [lx, 0, lz], [lx, 0, 0]],          # y = 0 (normal +y)
        [[0, ly, 0], [lx, ly, 0], [lx, ly, lz], [0, ly, lz]],      # y = ly (normal -y)
    ], dtype=np.float64)
    return _finish([(q, np.full(6, 4, np.int32))], "shoebox")


def _grid_polys(G, flip=False):
    """G: (n+1, m+1, 3) lattice of points -> mixed quads / triangle pairs.

    Cell (i,j) has corners a=G[i,j] b=G[i+1,j] c=G[i+1,j+1] d=G[i,j+1].
    h = hash(i,j) % 4: 0 -> triangles (a,b,c),(a,c,d); 1 -> (a,b,d),(b,c,d); else quad (a,b,c,d).
    Order inside a cell is kept (cell-major) so polygon indices are spatially coherent.
    """
    n, m = G.shape[0] - 1, G.shape[1] - 1
    a = G[:-1, :-1].reshape(-1, 3); b = G[1:, :-1].reshape(-1, 3)
    c = G[1:, 1:].reshape(-1, 3); d = G[:-1, 1:].reshape(-1, 3)
    if flip:
        b, d = d, b
    ii, jj = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(m, dtype=np.int64), indexing="ij")
    h = (((ii * 73856093) ^ (jj * 19349663)) % 4).reshape(-1)
    ncell = n * m
    # two slots per cell; second slot unused for quads
    V = np.zeros((ncell, 2, 4, 3), dtype=np.float64)
    C = np.zeros((ncell, 2), dtype=np.int32)
    q = h >= 2
    V[q, 0, 0], V[q, 0, 1], V[q, 0, 2], V[q, 0, 3] = a[q], b[q], c[q], d[q]
    C[q, 0] = 4
    t0 = h == 0
    V[t0, 0, 0], V[t0, 0, 1], V[t0, 0, 2], V[t0, 0, 3] = a[t0], b[t0], c[t0], c[t0]
    V[t0, 1, 0], V[t0, 1, 1], V[t0, 1, 2], V[t0, 1, 3] = a[t0], c[t0], d[t0], d[t0]
    t1 = h == 1
    V[t1, 0, 0], V[t1, 0, 1], V[t1, 0, 2], V[t1, 0, 3] = a[t1], b[t1], d[t1], d[t1]
    V[t1, 1, 0], V[t1, 1, 1], V[t1, 1, 2], V[t1, 1, 3] = b[t1], c[t1], d[t1], d[t1]
    C[t0 | t1, :] = 3
    keep = C.reshape(-1) > 0
    return V.reshape(-1, 4, 3)[keep], C.reshape(-1)[keep]


def _box_polys(lo, hi, n):
    """Closed axis-aligned box [lo,hi] with faces tessellated n=(nx,ny,nz)."""
    (x0, y0, z0), (x1, y1, z1) = lo, hi
    nx, ny, nz = n
    xs = x0 + (x1 - x0) * (np.arange(nx + 1) / nx)
    ys = y0 + (y1 - y0) * (np.arange(ny + 1) / ny)
    zs = z0 + (z1 - z0) * (np.arange(nz + 1) / nz)
    xs[-1], ys[-1], zs[-1] = x1, y1, z1
    out = []

    def lattice(A, B, fixed_axis, val):
        g = np.empty((A.size, B.size, 3))
        ax = [k for k in range(3) if k != fixed_axis]
        g[..., ax[0]] = A[:, None]
        g[..., ax[1]] = B[None, :]
        g[..., fixed_axis] = val
        return g
    out.append(_grid_polys(lattice(xs, ys, 2, z0), flip=True))
    out.append(_grid_polys(lattice(xs, ys, 2, z1)))
    out.append(_grid_polys(lattice(ys, zs, 0, x0), flip=True))
    out.append(_grid_polys(lattice(ys, zs, 0, x1)))
    out.append(_grid_polys(lattice(xs, zs, 1, y0)))
    out.append(_grid_polys(lattice(xs, zs, 1, y1), flip=True))
    return out


_LEVELS = {
    #         nu   nv  nz  balcony tess   seat rows, per row, seat tess (x, y, z)
    "tiny": (6, 6, 2, (2, 1, 1), 0, 0, (1, 1, 1)),
    "2k": (20, 20, 4, (4, 2, 1), 2, 3, (2, 1, 2)),
    "10k": (50, 50, 8, (8, 3, 1), 3, 6, (3, 1, 3)),
    "50k": (110, 110, 14, (24, 8, 2), 6, 10, (4, 1, 4)),
    "500k": (350, 350, 36, (60, 20, 3), 12, 24, (6, 2, 8)),
    "2m": (250, 250, 25, (40, 12, 2), 40, 50, (16, 3, 13)),
}

DEPTH, W0, W1, WMAX = 40.0, 20.0, 30.0, 30.0


def _plan_xy(u, v):
    x = WMAX / 2 + (u - 0.5) * (W0 + (W1 - W0) * v)
    y = DEPTH * v
    return x, y


def _floor_z(u, v, nsteps):
    rise = np.floor(np.maximum(v - 0.25, 0.0) * nsteps) / nsteps * 8.0   # stepped rake
    bowl = 0.5 * ((2 * u - 1) * (2 * u - 1)) * v
    return rise + bowl


def _ceil_z(u, v, cu, cv):
    fu = u * cu - np.floor(u * cu)
    fv = v * cv - np.floor(v * cv)
    return 14.0 + 3.0 * (1.0 - v) + 0.6 * np.abs(fu - 0.5) + 0.6 * np.abs(fv - 0.5)


def hall(level="50k"):
    nu, nv, nz, btess, srows, sper, stess = _LEVELS[level]
    us = np.arange(nu + 1) / nu
    vs = np.arange(nv + 1) / nv
    U, V = np.meshgrid(us, vs, indexing="ij")
    X, Y = _plan_xy(U, V)
    nsteps = max(4, nv // 3)
    ZF = _floor_z(U, V, nsteps)
    ZC = _ceil_z(U, V, 7.0, 9.0)
    polys = []
    polys.append(_grid_polys(np.stack([X, Y, ZF], axis=-1)))
    polys.append(_grid_polys(np.stack([X, Y, ZC], axis=-1), flip=True))
    # walls: boundary loop of the (u,v) lattice, nz vertical segments
    bi = np.concatenate([np.arange(nu), np.full(nv, nu), np.arange(nu, 0, -1), np.zeros(nv, np.int64)]).astype(np.int64)
    bj = np.concatenate([np.zeros(nu, np.int64), np.arange(nv), np.full(nu, nv), np.arange(nv, 0, -1)]).astype(np.int64)
    bi = np.append(bi, bi[0]); bj = np.append(bj, bj[0])
    k = (np.arange(nz + 1) / nz)[None, :]
    WX = np.repeat(X[bi, bj][:, None], nz + 1, axis=1)
    WY = np.repeat(Y[bi, bj][:, None], nz + 1, axis=1)
    zf, zc = ZF[bi, bj][:, None], ZC[bi, bj][:, None]
    WZ = zf + (zc - zf) * k
    WZ[:, -1] = zc[:, 0]
    polys.append(_grid_polys(np.stack([WX, WY, WZ], axis=-1), flip=True))
    # balcony slab (free-standing closed box)
    polys += _box_polys((6.0, 33.0, 9.0), (24.0, 38.5, 9.6), btess)
    # seat-back blocks, floating 5 cm above the local floor
    for r in range(srows):
        v = 0.30 + 0.45 * (r + 0.5) / max(srows, 1)
        yc = DEPTH * v
        half = 0.5 * (W0 + (W1 - W0) * v) * 0.7
        zbase = float(_floor_z(np.array(0.0 + 0.15), np.array(v + 0.02), nsteps)) + 0.05
        for s in range(sper):
            xc = WMAX / 2 + (-half + 2 * half * (s + 0.5) / sper)
            w = min(0.5, 0.8 * half / sper)
            polys += _box_polys((xc - w / 2, yc - 0.04, zbase), (xc + w / 2, yc + 0.04, zbase + 0.9), stess)
    return _finish(polys, "hall-" + level)


def sources(n):
    """n interior source points (well inside every hall level)."""
    base = np.array([[15.0, 6.0, 5.0], [11.0, 14.0, 8.0], [19.0, 22.0, 10.5], [15.0, 30.0, 11.5],
                     [13.0, 9.0, 6.5], [17.0, 17.0, 9.5], [9.5, 26.0, 11.0], [20.5, 11.0, 7.5]])
    return base[:n].copy()


def lattice_room(n=32, size=16.0, height=8.0):
    """Closed box room whose floor, ceiling and walls are n x n lattices of mixed triangles / quads with DYADIC coordinates
    (size / n a power of two): rays with small-integer origins and directions hit shared edges and vertices at bit-identical t
    for every adjacent polygon -- the exact-t ties of the KDTree / Octree tie rules."""
    return _finish(_box_polys((0.0, 0.0, 0.0), (size, size, height), (n, n, max(1, n // 2))), "lattice-room")
