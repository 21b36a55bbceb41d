"""Second, independent restatement of Hare's Shoot path in plain Python (TEST INFRASTRUCTURE ONLY).

Why it exists: the reference ships no tests or golden vectors and cannot be run here (C#, no
.NET in the image), so parity is UNPINNED by the reference itself.
One of the substitute pins is that two restatements written separately from the C# sources --
oracle/hare_oracle.cpp (C++) and this file (Python floats are IEEE binary64; CPython never
contracts a*b+c) -- agree bit for bit.  tests/test_oracle_pins.py checks that, and
tests/golden/make_golden.py uses this file to produce the committed golden vectors.

Written to read like the C# it follows; every function cites the reference file:line.
Slow (pure Python loops): small cases only.
"""
import math

DBL_MAX = 1.7976931348623157e308
DENORM_MIN = 5e-324          # double.Epsilon
INT_MIN = -2147483648


# ---------------------------------------------------------------- BCL semantics
def net_max(a, b):           # Math.Max: NaN-propagating
    if a != a:
        return a
    if b != b:
        return b
    return a if a > b else (b if b > a else (b if math.copysign(1.0, a) < 0 else a))


def net_min(a, b):           # Math.Min
    if a != a:
        return a
    if b != b:
        return b
    return a if a < b else (b if b < a else (a if math.copysign(1.0, a) < 0 else b))


def floor_to_int(x):         # (int)Math.Floor(x), x64 cvttsd2si
    if x != x or math.isinf(x):
        return INT_MIN
    f = math.floor(x)
    if f < -2147483648 or f >= 2147483648:
        return INT_MIN
    return int(f)


def round15(x):              # Math.Round(x, 15)
    if abs(x) < 1e16:
        x = x * 1e15
        x = float(round(x))  # half to even
        x = x / 1e15
    return x


def dot(ax, ay, az, bx, by, bz):   # Hare_Geometry_Math.cs:43-46
    return (ax * bx) + (ay * by) + (az * bz)


def cross(a, b):             # Hare_Geometry_Math.cs:62-73
    return (a[1] * b[2] - a[2] * b[1], -(a[0] * b[2] - a[2] * b[0]), a[0] * b[1] - a[1] * b[0])


# ---------------------------------------------------------------- AABB
class AABB:
    def __init__(self, mn, mx):          # AABB_Main.cs:39-67
        self.Min, self.Max = tuple(mn), tuple(mx)
        self.Center = tuple((self.Max[a] + self.Min[a]) / 2 for a in range(3))
        width = tuple(self.Max[a] - self.Min[a] for a in range(3))
        self.halfwidth = tuple(w / 2 for w in width)

    def IsPointInBox(self, x, y, z):     # AABB_Main.cs:75-84
        if x < self.Min[0]: return False
        if y < self.Min[1]: return False
        if z < self.Min[2]: return False
        if x > self.Max[0]: return False
        if y > self.Max[1]: return False
        if z > self.Max[2]: return False
        return True

    def Intersect(self, R):              # AABB_Main.cs:173-260; returns (ok, tmin) and moves R
        tmin, tmax = 0.0, DBL_MAX
        o = [R.x, R.y, R.z]; d = [R.dx, R.dy, R.dz]
        for a in range(3):
            if abs(d[a]) < DENORM_MIN:
                if o[a] < self.Min[a] or o[a] > self.Max[a]:
                    return False, tmin
            else:
                ood = 1 / d[a]
                t1 = (self.Min[a] - o[a]) * ood
                t2 = (self.Max[a] - o[a]) * ood
                if t1 > t2:
                    t1, t2 = t2, t1
                tmin = net_max(tmin, t1)
                tmax = net_min(tmax, t2)
                if tmin > tmax:
                    return False, tmin
        R.x = R.x + R.dx * tmin
        R.y = R.y + R.dy * tmin
        R.z = R.z + R.dz * tmin
        return True, tmin

    # ---- AABB_Tri_Int.cs ----
    def _plane_box(self, n, vert, maxbox):          # :51-95
        vmin = [0.0] * 3; vmax = [0.0] * 3
        for a in range(3):
            v = vert[a]
            if n[a] > 0.0:
                vmin[a] = -maxbox[a] - v; vmax[a] = maxbox[a] - v
            else:
                vmin[a] = maxbox[a] - v; vmax[a] = -maxbox[a] - v
        if dot(n[0], n[1], n[2], vmin[0], vmin[1], vmin[2]) > 0.0:
            return False
        if dot(n[0], n[1], n[2], vmax[0], vmax[1], vmax[2]) >= 0.0:
            return True
        return False

    def PolyBoxOverlap(self, P):                    # :165-260
        hw = self.halfwidth
        C = self.Center
        for j in range(1, len(P) - 1):
            tri = (P[0], P[j], P[j + 1])
            v0, v1, v2 = (tuple(tri[k][a] - C[a] for a in range(3)) for k in range(3))
            e0 = tuple(v1[a] - v0[a] for a in range(3))
            e1 = tuple(v2[a] - v1[a] for a in range(3))
            e2 = tuple(v0[a] - v2[a] for a in range(3))

            def ax_x(a, b, fa, fb, p, q):           # AXISTEST_X01 / X2 (:101-118)
                pa = a * p[1] - b * p[2]; pb = a * q[1] - b * q[2]
                mn, mx = (pa, pb) if pa < pb else (pb, pa)
                rad = fa * hw[1] + fb * hw[2]
                return not (mn > rad or mx < -rad)

            def ax_y(a, b, fa, fb, p, q):           # AXISTEST_Y02 / Y1 (:121-139)
                pa = -a * p[0] + b * p[2]; pb = -a * q[0] + b * q[2]
                mn, mx = (pa, pb) if pa < pb else (pb, pa)
                rad = fa * hw[0] + fb * hw[2]
                return not (mn > rad or mx < -rad)

            def ax_z12(a, b, fa, fb):               # AXISTEST_Z12 (:143-150): note (p2 < p1)
                p1 = a * v1[0] - b * v1[1]; p2 = a * v2[0] - b * v2[1]
                mn, mx = (p2, p1) if p2 < p1 else (p1, p2)
                rad = fa * hw[0] + fb * hw[1]
                return not (mn > rad or mx < -rad)

            def ax_z0(a, b, fa, fb):                # AXISTEST_Z0 (:153-160)
                p0 = a * v0[0] - b * v0[1]; p1 = a * v1[0] - b * v1[1]
                mn, mx = (p0, p1) if p0 < p1 else (p1, p0)
                rad = fa * hw[0] + fb * hw[1]
                return not (mn > rad or mx < -rad)

            fex, fey, fez = abs(e0[0]), abs(e0[1]), abs(e0[2])
            if not ax_x(e0[2], e0[1], fez, fey, v0, v2): continue
            if not ax_y(e0[2], e0[0], fez, fex, v0, v2): continue
            if not ax_z12(e0[1], e0[0], fey, fex): continue
            fex, fey, fez = abs(e1[0]), abs(e1[1]), abs(e1[2])
            if not ax_x(e1[2], e1[1], fez, fey, v0, v2): continue
            if not ax_y(e1[2], e1[0], fez, fex, v0, v2): continue
            if not ax_z0(e1[1], e1[0], fey, fex): continue
            fex, fey, fez = abs(e2[0]), abs(e2[1]), abs(e2[2])
            if not ax_x(e2[2], e2[1], fez, fey, v0, v1): continue
            if not ax_y(e2[2], e2[0], fez, fex, v0, v1): continue
            if not ax_z12(e2[1], e2[0], fey, fex): continue
            sep = False
            for a in range(3):                      # FINDMINMAX + box axes (:41-49, :240-249)
                mn = mx = v0[a]
                if v1[a] < mn: mn = v1[a]
                if v1[a] > mx: mx = v1[a]
                if v2[a] < mn: mn = v2[a]
                if v2[a] > mx: mx = v2[a]
                if mn > hw[a] or mx < -hw[a]:
                    sep = True
                    break
            if sep: continue
            if not self._plane_box(cross(e0, e1), v0, hw): continue
            return True
        return False


# ---------------------------------------------------------------- primitives
class Ray:
    def __init__(self, x, y, z, dx, dy, dz, Ray_ID=1):
        self.x, self.y, self.z, self.dx, self.dy, self.dz, self.Ray_ID = x, y, z, dx, dy, dz, Ray_ID


class X_Event:
    def __init__(self, P=None, u=0.0, v=0.0, t=0.0, Poly_index=-1):
        self.Hit = P is not None
        self.X_Point, self.u, self.v, self.t, self.Poly_id = P, u, v, t, Poly_index


# ---------------------------------------------------------------- Topology
class Topology:
    """Topology(Point min, Point max) + Add_Polygon + Finish_Topology (Hare_Geometry_Topology.cs)."""

    def __init__(self, minpt, maxpt):               # :85-91
        self.Max = tuple(m + 0.000000000001 for m in maxpt)
        self.Min = tuple(m - 0.000000000001 for m in minpt)
        self._modspace(self.Min, self.Max)
        self.weld = {}
        self.Vertices_List = []
        self.Polys = []       # (points tuple, normal, vertex count)

    def _modspace(self, mn, mx):                    # MS_AABB :677-697
        self.ms_min = mn
        self.ms_dim = max(int(math.ceil(mx[a] - mn[a])) for a in range(3))
        self.ms_xytot = self.ms_dim * self.ms_dim

    def _add_get_index(self, p):                    # :342-377 + Primitives.cs:230-250
        q = (round15(p[0]), round15(p[1]), round15(p[2]))
        off = [q[a] - self.ms_min[a] for a in range(3)]
        loc = [int(math.floor(off[a])) for a in range(3)]
        bucket = self.ms_xytot * loc[2] + self.ms_dim * loc[0] + loc[1]
        pos3 = [int((off[a] - loc[a]) * 1000) for a in range(3)]
        pos = 1000000 * pos3[2] + 1000 * pos3[0] + pos3[1]
        key = (bucket, pos)
        if key in self.weld:
            return self.Vertices_List[self.weld[key]]
        self.weld[key] = len(self.Vertices_List)
        self.Vertices_List.append(q)
        return q

    def Add_Polygon(self, P):                       # :225-254 + Polygons.cs:148-171
        if len(P) not in (3, 4):
            raise NotImplementedError("Hare Does not yet support polygons of more than 4 sides.")
        V = [self._add_get_index(p) for p in P]
        N = (0.0, 0.0, 0.0)
        for j in range(2, len(V)):
            N = cross(tuple(V[1][a] - V[0][a] for a in range(3)), tuple(V[j][a] - V[0][a] for a in range(3)))
            if not ((N[0] * N[0] + N[1] * N[1] + N[2] * N[2]) < DENORM_MIN):
                break
        f = N[0] * N[0] + N[1] * N[1] + N[2] * N[2]
        if f != 0:
            f = math.sqrt(f)
            N = (N[0] / f, N[1] / f, N[2] / f)
        self.Polys.append((tuple(V), N, len(V)))

    def Finish_Topology(self):                      # :148-167
        mn = [DBL_MAX] * 3; mx = [-DBL_MAX] * 3
        for p in self.Vertices_List:
            for a in range(3):
                if mn[a] > p[a]: mn[a] = p[a]
                if mx[a] < p[a]: mx[a] = p[a]
        self._modspace(tuple(mn), tuple(mx))
        self.Min = tuple(m - 0.000000000001 for m in mn)
        self.Max = tuple(m + 0.000000000001 for m in mx)

    @property
    def Polygon_Count(self):
        return len(self.Polys)

    def Polygon_Vertices(self, i):
        return self.Polys[i][0]

    def Polygon_Centroid(self, i):                  # :566-575
        P = self.Polys[i][0]
        s = (0.0, 0.0, 0.0)
        for p in P:
            s = (s[0] + p[0], s[1] + p[1], s[2] + p[2])
        n = len(P)
        return (s[0] / n, s[1] / n, s[2] / n)

    # ---- Polygon.Intersect ----
    def _ray_side(self, i, R):                      # Polygons.cs:601-606
        N = self.Polys[i][1]
        return not (dot(R.dx, R.dy, R.dz, N[0], N[1], N[2]) < 0)

    @staticmethod
    def _rayxtri_fast(R, a, b, c):                  # Polygons.cs:449-510 -> (hit, t)
        e1 = (b[0] - a[0], b[1] - a[1], b[2] - a[2]); e2 = (c[0] - a[0], c[1] - a[1], c[2] - a[2])
        px = R.dy * e2[2] - R.dz * e2[1]; py = R.dz * e2[0] - R.dx * e2[2]; pz = R.dx * e2[1] - R.dy * e2[0]
        det = dot(e1[0], e1[1], e1[2], px, py, pz)
        tx, ty, tz = R.x - a[0], R.y - a[1], R.z - a[2]
        qx = ty * e1[2] - tz * e1[1]; qy = tz * e1[0] - tx * e1[2]; qz = tx * e1[1] - ty * e1[0]
        if det > 0.000001:
            u = dot(tx, ty, tz, px, py, pz)
            if u < 0.0 or u > det: return False, 0.0
            v = dot(R.dx, R.dy, R.dz, qx, qy, qz)
            if v < 0.0 or u + v > det: return False, 0.0
        elif det < -0.000001:
            u = dot(tx, ty, tz, px, py, pz)
            if u > 0.0 or u < det: return False, 0.0
            v = dot(R.dx, R.dy, R.dz, qx, qy, qz)
            if v > 0.0 or u + v < det: return False, 0.0
        else:
            return False, 0.0
        invdet = 1.0 / det
        return True, dot(e2[0], e2[1], e2[2], qx, qy, qz) * invdet

    @staticmethod
    def _rayxtri_slow(R, a, b, c):                  # Polygons.cs:385-435 -> (hit, t, u, v)
        e1 = (b[0] - a[0], b[1] - a[1], b[2] - a[2]); e2 = (c[0] - a[0], c[1] - a[1], c[2] - a[2])
        p = cross((R.dx, R.dy, R.dz), e2)
        det = (e1[0] * p[0]) + (e1[1] * p[1]) + (e1[2] * p[2])
        tv = (R.x - a[0], R.y - a[1], R.z - a[2])
        q = cross(tv, e1)
        if det > 0.000001:
            u = dot(tv[0], tv[1], tv[2], p[0], p[1], p[2])
            if u < 0.0 or u > det: return False, 0.0, 0.0, 0.0
            v = dot(R.dx, R.dy, R.dz, q[0], q[1], q[2])
            if v < 0.0 or u + v > det: return False, 0.0, 0.0, 0.0
        elif det < -0.000001:
            u = dot(tv[0], tv[1], tv[2], p[0], p[1], p[2])
            if u > 0.0 or u < det: return False, 0.0, 0.0, 0.0
            v = dot(R.dx, R.dy, R.dz, q[0], q[1], q[2])
            if v > 0.0 or u + v < det: return False, 0.0, 0.0, 0.0
        else:
            return False, 0.0, 0.0, 0.0
        invdet = 1.0 / det
        t = ((e2[0] * q[0]) + (e2[1] * q[1]) + (e2[2] * q[2])) * invdet
        return True, t, u * invdet, v * invdet

    def intersect_fast(self, i, R):                 # Polygons.cs:637-660, 784-823 -> (hit, x, y, z, t)
        P = self.Polys[i][0]
        if self._ray_side(i, R):
            order = [(P[0], P[1], P[2])] + ([(P[2], P[3], P[0])] if len(P) == 4 else [])
        else:
            order = [(P[2], P[1], P[0])] + ([(P[0], P[3], P[2])] if len(P) == 4 else [])
        for a, b, c in order:
            hit, t = self._rayxtri_fast(R, a, b, c)
            if hit:
                return True, R.x + R.dx * t, R.y + R.dy * t, R.z + R.dz * t, t
        return False, 0.0, 0.0, 0.0, 0.0

    def intersect_slow(self, i, R):                 # Polygons.cs:662-688, 731-782 -> (hit, X, u, v, t)
        P = self.Polys[i][0]
        if self._ray_side(i, R):
            order = [(P[0], P[1], P[2])] + ([(P[2], P[3], P[0])] if len(P) == 4 else [])
        else:
            order = [(P[2], P[1], P[0])] + ([(P[0], P[3], P[2])] if len(P) == 4 else [])
        for a, b, c in order:
            hit, t, u, v = self._rayxtri_slow(R, a, b, c)
            if hit:
                return True, (R.x + R.dx * t, R.y + R.dy * t, R.z + R.dz * t), u, v, t
        return False, None, 0.0, 0.0, 0.0


# ---------------------------------------------------------------- Voxel_Grid
class Voxel_Grid:
    Epsilon = 0.001

    def __init__(self, Model, Domain):              # Voxel_Grid.cs:48-121, single topology
        self.Model = Model
        T = Model[0]
        self.mail = [0] * T.Polygon_Count           # one Poly_Ray_ID slot, zero-initialised (:54-62)
        MaxPT = [-math.inf] * 3; MinPT = [math.inf] * 3
        for a in range(3):
            if (T.Max[a] + 0.01) > MaxPT[a]: MaxPT[a] = T.Max[a] + self.Epsilon
            if (T.Min[a] - 0.01) < MinPT[a]: MinPT[a] = T.Min[a] - self.Epsilon
        self.OBox = AABB([m - .1 for m in MinPT], [m + .1 for m in MaxPT])
        self.Ct = (Domain, Domain, Domain)
        self.BoxDims = tuple(self.OBox.Max[a] - self.OBox.Min[a] for a in range(3))
        self.VoxelDims = tuple(self.BoxDims[a] / self.Ct[a] for a in range(3))
        self.Voxel_Inv = {}
        for x in range(Domain):                     # Fill_Voxels :273-304
            for y in range(Domain):
                for z in range(Domain):
                    box = self.voxel(x, y, z)
                    self.Voxel_Inv[(x, y, z)] = [i for i in range(T.Polygon_Count) if box.PolyBoxOverlap(T.Polygon_Vertices(i))]

    def voxel(self, x, y, z):                       # :283-285
        vd, e, om = self.VoxelDims, self.Epsilon, self.OBox.Min
        mn = (x * vd[0] - e, y * vd[1] - e, z * vd[2] - e)
        mx = ((x + 1) * vd[0] + e, (y + 1) * vd[1] + e, (z + 1) * vd[2] + e)
        return AABB([mn[a] + om[a] for a in range(3)], [mx[a] + om[a] for a in range(3)])

    def Shoot(self, R, top_index=0, poly_origin1=-1, poly_origin2=-1):   # :351-552 ; returns (hit, X_Event)
        T = self.Model[top_index]
        om, vd, Ct = self.OBox.Min, self.VoxelDims, self.Ct
        X = floor_to_int((R.x - om[0]) / vd[0])
        Y = floor_to_int((R.y - om[1]) / vd[1])
        Z = floor_to_int((R.z - om[2]) / vd[2])
        t_start = 0.0
        if X < 0 or X >= Ct[0] or Y < 0 or Y >= Ct[1] or Z < 0 or Z >= Ct[2]:
            ok, t_start = self.OBox.Intersect(R)
            if not ok:
                return False, X_Event()
            X = floor_to_int((R.x - om[0] + R.dx * 1E-6) / vd[0])
            Y = floor_to_int((R.y - om[1] + R.dy * 1E-6) / vd[1])
            Z = floor_to_int((R.z - om[2] + R.dz * 1E-6) / vd[2])
            if X < 0 or X >= Ct[0] or Y < 0 or Y >= Ct[1] or Z < 0 or Z >= Ct[2]:
                raise IndexError("Voxels[X, Y, Z]")
        V0 = self.voxel(X, Y, Z)

        def div(a, b):                               # IEEE division incl. /0
            try:
                return a / b
            except ZeroDivisionError:
                if a != a or a == 0: return math.nan
                return math.copysign(math.inf, a) * math.copysign(1.0, b)
        step = [0, 0, 0]; tMax = [0.0] * 3; tDelta = [0.0] * 3
        o = (R.x, R.y, R.z); d = (R.dx, R.dy, R.dz)
        for a in range(3):
            if d[a] < 0:
                step[a] = -1; tMax[a] = div(V0.Min[a] - o[a], d[a]); tDelta[a] = div(vd[a], d[a]) * step[a]
            else:
                step[a] = 1; tMax[a] = div(V0.Max[a] - o[a], d[a]); tDelta[a] = div(vd[a], d[a]) * step[a]
        Xpt = None; tmin = DBL_MAX; pid = -1
        P = [X, Y, Z]
        while True:
            for i in self.Voxel_Inv[(P[0], P[1], P[2])]:
                if i == poly_origin1 or i == poly_origin2: continue
                if self.mail[i] != R.Ray_ID:
                    self.mail[i] = R.Ray_ID
                    hit, x, y, z, t = T.intersect_fast(i, R)
                    if hit and t > 0.0000000001:
                        if t < tmin:
                            Xpt = (x, y, z); tmin = t; pid = i
            if Xpt is not None and self.voxel(P[0], P[1], P[2]).IsPointInBox(*Xpt):
                return True, X_Event(Xpt, 0.0, 0.0, tmin + t_start, pid)
            if tMax[0] < tMax[1]:
                a = 0 if tMax[0] < tMax[2] else 2
            else:
                a = 1 if tMax[1] < tMax[2] else 2
            P[a] += step[a]
            if P[a] < 0 or P[a] >= Ct[a]:
                return False, X_Event()
            tMax[a] = tMax[a] + tDelta[a]


# ---------------------------------------------------------------- Octree ("Octree - alt.cs")
class _ONode(AABB):
    def __init__(self, mn, mx):
        super().__init__(mn, mx)
        self.Polygons = []
        self.Children = None


class Octree:
    def __init__(self, Model, maxDepth, maxPolygonsPerNode):       # :45-89
        self.Model, self.maxDepth, self.maxPolys = Model, maxDepth, maxPolygonsPerNode
        topo = Model[0]
        mn = [math.inf] * 3; mx = [-math.inf] * 3
        for v in topo.Vertices_List:
            for a in range(3):
                if v[a] < mn[a]: mn[a] = v[a]
                if v[a] > mx[a]: mx[a] = v[a]
        maxdim = net_max(mx[0] - mn[0], net_max(mx[1] - mn[1], mx[2] - mn[2]))
        center = [mx[a] + mn[a] / 2 for a in range(3)]               # "max + min / 2" :79
        self.root = _ONode([center[a] - maxdim - 1e-1 for a in range(3)], [center[a] + maxdim + 1e-1 for a in range(3)])
        self.root.Polygons = list(range(topo.Polygon_Count))
        self._build(self.root, 0)

    def _build(self, node, depth):                                  # :91-138
        if depth >= self.maxDepth or len(node.Polygons) <= self.maxPolys:
            return
        c = node.Center
        node.Children = []
        for i in range(8):
            mn = [((node.Min[a] if (i & (4 >> a)) == 0 else c[a]) - 0.1) for a in range(3)]
            mx = [((c[a] if (i & (4 >> a)) == 0 else node.Max[a]) + 0.1) for a in range(3)]
            node.Children.append(_ONode(mn, mx))
        for p in node.Polygons:
            verts = self.Model[0].Polygon_Vertices(p)
            for ch in node.Children:
                if ch.PolyBoxOverlap(verts):
                    ch.Polygons.append(p)
        node.Polygons = []
        for ch in node.Children:
            self._build(ch, depth + 1)

    def Shoot(self, ray, top_index=0, poly_origin1=-1, poly_origin2=-1):    # :159-306
        T = self.Model[top_index]
        inv = [(1.0 / dv if abs(dv) > 1e-16 else 1e16) for dv in (ray.dx, ray.dy, ray.dz)]
        o = (ray.x, ray.y, ray.z)

        def interval(n):
            t0 = [(n.Min[a] - o[a]) * inv[a] for a in range(3)]
            t1 = [(n.Max[a] - o[a]) * inv[a] for a in range(3)]
            for a in range(3):
                if inv[a] < 0:
                    t0[a], t1[a] = t1[a], t0[a]
            return net_max(net_max(t0[0], t0[1]), t0[2]), net_min(net_min(t1[0], t1[1]), t1[2])
        tmin, tmax = interval(self.root)
        if tmax < tmin or tmax < 0:
            return False, X_Event()
        order = []
        dirs = [(0 if dv >= 0 else 1) for dv in (ray.dx, ray.dy, ray.dz)]
        inc = [(1 if dv >= 0 else -1) for dv in (ray.dx, ray.dy, ray.dz)]
        ix = dirs[0]
        while 0 <= ix <= 1:
            iy = dirs[1]
            while 0 <= iy <= 1:
                iz = dirs[2]
                while 0 <= iz <= 1:
                    order.append((ix << 2) | (iy << 1) | iz)
                    iz += inc[2]
                iy += inc[1]
            ix += inc[0]
        stack = [(self.root, tmin, tmax)]
        hit = False; closestT = DBL_MAX; ev = None
        while stack:
            node, a, b = stack.pop()
            if b < a or b < 0: continue
            if hit and closestT <= a: continue
            if node.Children is None:
                for p in node.Polygons:
                    if p == poly_origin1 or p == poly_origin2: continue
                    h, Xp, u, v, t = T.intersect_slow(p, ray)
                    if h and t > 0.0000000001:
                        if t < closestT:
                            closestT = t
                            ev = X_Event(Xp, u, v, t, p)
                            hit = True
                            if closestT <= a:
                                return True, ev
            else:
                for ci in order:
                    ch = node.Children[ci]
                    ca, cb = interval(ch)
                    if cb < ca or cb < 0 or ca > b or cb < a: continue
                    stack.append((ch, net_max(ca, a), net_min(cb, b)))
        if hit:
            return True, ev
        return False, X_Event()


# ---------------------------------------------------------------- KDTree (KDTree.cs)
class _KNode(AABB):
    def __init__(self, mn, mx):
        super().__init__(mn, mx)
        self.Polygons = []
        self.Left = self.Right = None
        self.SplitAxis = 0; self.SplitValue = 0.0


class KDTree:
    def __init__(self, Model, maxDepth, maxPolygonsPerNode):       # :51-88
        self.Model, self.maxDepth, self.maxPolys = Model, maxDepth, maxPolygonsPerNode
        topo = Model[0]
        self.mail = [0] * topo.Polygon_Count
        mn = [math.inf] * 3; mx = [-math.inf] * 3
        for v in topo.Vertices_List:
            for a in range(3):
                if v[a] < mn[a]: mn[a] = v[a]
                if v[a] > mx[a]: mx[a] = v[a]
        self.root = _KNode(mn, mx)
        self.root.Polygons = list(range(topo.Polygon_Count))
        self._build(self.root, 0, list(mn), list(mx))

    def _build(self, node, depth, mn, mx):                          # :90-139
        if depth >= self.maxDepth or len(node.Polygons) <= self.maxPolys:
            return
        topo = self.Model[0]
        axis = depth % 3
        cen = {p: topo.Polygon_Centroid(p)[axis] for p in node.Polygons}
        srt = sorted(node.Polygons, key=lambda p: cen[p])            # OrderBy: stable
        split = cen[srt[len(srt) // 2]]
        node.SplitAxis, node.SplitValue = axis, split
        lmx = list(mx); lmx[axis] = split
        rmn = list(mn); rmn[axis] = split
        node.Left = _KNode(mn, lmx); node.Right = _KNode(rmn, mx)
        for p in srt:
            verts = topo.Polygon_Vertices(p)
            if any(v[axis] <= split for v in verts): node.Left.Polygons.append(p)
            if any(v[axis] > split for v in verts): node.Right.Polygons.append(p)
        node.Polygons = []
        self._build(node.Left, depth + 1, mn, lmx)
        self._build(node.Right, depth + 1, rmn, mx)

    def Shoot(self, ray, top_index=0, poly_origin1=-1, poly_origin2=-1):    # :198-361
        T = self.Model[top_index]
        ev = X_Event(); hit = False; closestT = DBL_MAX
        stack = [self.root]
        o = (ray.x, ray.y, ray.z); d = (ray.dx, ray.dy, ray.dz)
        while stack:
            cur = stack.pop()
            if cur.Left is None and cur.Right is None:
                for p in cur.Polygons:
                    if p == poly_origin1 or p == poly_origin2: continue
                    if self.mail[p] == ray.Ray_ID: continue
                    self.mail[p] = ray.Ray_ID
                    h, Xp, u, v, t = T.intersect_slow(p, ray)
                    if h and t > 0.0000000001:
                        if t < closestT:
                            closestT = t
                            ev = X_Event(Xp, u, v, t, p)
                            hit = True
            else:
                a = cur.SplitAxis
                b1, b2 = [k for k in range(3) if k != a]
                side = o[a] - cur.SplitValue
                try:
                    tSplit = -side / d[a]
                except ZeroDivisionError:
                    tSplit = math.nan if side == 0 else math.copysign(math.inf, -side) * math.copysign(1.0, d[a])
                s1 = o[b1] + tSplit * d[b1]
                s2 = o[b2] + tSplit * d[b2]
                inside = s1 <= cur.Max[b1] and s1 >= cur.Min[b1] and s2 <= cur.Max[b2] and s2 >= cur.Min[b2]
                if inside:
                    first, second = (cur.Right, cur.Left) if side >= 0 else (cur.Left, cur.Right)
                else:
                    first, second = (cur.Left, cur.Right) if side >= 0 else (cur.Right, cur.Left)
                stack.append(second)
                stack.append(first)
        return hit, ev
