"""Host-side mirror of Hare's API for the Shoot path, over the C ABI.

Hare is C#; no .NET toolchain exists in this image, so this module plays the part of the
Hare_NC host for tests and benchmarks.  Names, argument meaning and error behaviour follow
the reference (paths relative to the reference checkout):

    Topology(minpt, maxpt) / Add_Polygon / Finish_Topology   Hare_Geometry_Topology.cs:85-91, 225-254, 148-179
    Ray, X_Event                                             Hare_Geometry_Primitives.cs:393-481
    Spatial_Partition.Shoot (both overloads)                 Spatial_Partition.cs:32-33
    Voxel_Grid(Model, Domain)                                Voxel_Grid.cs:48
    Octree(Model, maxDepth, maxPolygonsPerNode)              "Octree - alt.cs":45
    KDTree(Model, maxDepth, maxPolygonsPerNode)              KDTree.cs:51
plus the new batched overload Shoot_Batch and Reflect_Chain.

All computation happens in libhare_b200.so (CUDA, sm_100a); nothing here computes results.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, ptr


@dataclass
class Point:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class Ray:
    """Hare.Geometry.Ray.  Shoot may move x, y, z (Voxel_Grid, ray starting outside the grid)."""
    x: float
    y: float
    z: float
    dx: float
    dy: float
    dz: float
    ThreadID: int = 0
    Ray_ID: int = 1
    poly_origin1: int = -1
    poly_origin2: int = -1


class X_Event:
    """Hare.Geometry.X_Event; X_Event() is the miss / empty event."""

    def __init__(self, P=None, u=0.0, v=0.0, t=0.0, Poly_index=-1):
        self.Hit = P is not None
        self.X_Point = P
        self.u, self.v, self.t, self.Poly_id = u, v, t, Poly_index

    def __repr__(self):
        return f"X_Event(Hit={self.Hit}, Poly_id={self.Poly_id}, t={self.t!r}, X_Point={self.X_Point})"


class Topology:
    """Polygon container.  Construction mirrors `new Topology(min, max)`, Add_Polygon, Finish_Topology()."""

    def __init__(self, Minpt, Maxpt):
        self._minpt = as_f64(Minpt, (3,))
        self._maxpt = as_f64(Maxpt, (3,))
        self._raw, self._cnt = [], []
        self._h = None
        self.verts = self.normals = self.vcount = None
        self.Min = self.Max = None
        self.Vertex_Count = 0

    @classmethod
    def from_mesh(cls, mesh):
        t = cls(mesh.minpt, mesh.maxpt)
        t._bulk = (as_f64(mesh.verts, (-1, 4, 3)), as_i32(mesh.vcount))
        t.Finish_Topology()
        return t

    def Add_Polygon(self, P):
        P = as_f64(P).reshape(-1, 3)
        if P.shape[0] not in (3, 4):
            raise NotImplementedError("Hare Does not yet support polygons of more than 4 sides.")
        q = np.zeros((4, 3))
        q[:P.shape[0]] = P
        self._raw.append(q)
        self._cnt.append(P.shape[0])

    def Finish_Topology(self):
        if getattr(self, "_bulk", None) is not None:
            raw, cnt = self._bulk
        else:
            raw, cnt = as_f64(np.array(self._raw)).reshape(-1, 4, 3), as_i32(np.array(self._cnt))
        P = int(cnt.shape[0])
        L = _lib.lib()
        verts = np.empty((P, 4, 3)); normals = np.empty((P, 3)); mm = np.empty(6); nv = C.c_int64()
        check(L.hare_topology_ingest(ptr(raw), ptr(cnt), P, ptr(self._minpt), ptr(self._maxpt), ptr(verts), ptr(normals), ptr(mm), C.byref(nv)),
              "hare_topology_ingest")
        self.verts, self.normals, self.vcount = verts, normals, cnt
        self.Min, self.Max, self.Vertex_Count = mm[:3].copy(), mm[3:].copy(), nv.value
        self._mm = mm
        if self._h:
            # partitions built on the previous flattening still read its device records: the library refuses to free it
            check(L.hare_topology_destroy(self._h), "Finish_Topology (re-finishing a Topology that still has live partitions)")
            self._h = None
        h = C.c_void_p()
        check(L.hare_topology_create(ptr(verts), ptr(normals), ptr(cnt), P, ptr(mm), C.byref(h)), "hare_topology_create")
        self._h = h

    @property
    def Polygon_Count(self):
        return 0 if self.vcount is None else int(self.vcount.shape[0])

    def Polygon_Vertices(self, i):
        return self.verts[i, :self.vcount[i]].copy()

    def Normal(self, i):
        return self.normals[i].copy()

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.hare_topology_destroy(self._h)
            self._h = None


class Spatial_Partition:
    """Abstract parent (Spatial_Partition.cs:27-35): Model, Char_Step, Shoot x2."""
    has_uv = True

    def __init__(self, Model):
        if isinstance(Model, Topology):
            Model = [Model]
        if len(Model) != 1:
            raise NotImplementedError("hare_b200 flattens a single Topology (Model[0]); see DESIGN.md, quirks Q8/Q9")
        if Model[0]._h is None:
            raise ValueError("call Topology.Finish_Topology() before building a partition")
        self.Model = list(Model)
        self.Char_Step = 0.0
        self._h = C.c_void_p()

    # ---- on-disk form (none in the reference) ------------------------
    def Save(self, path):
        """Write the flattened partition (cells or nodes and lists) to `path`."""
        check(_lib.lib().hare_part_save(self._h, os.fsencode(path)), "hare_part_save")

    @classmethod
    def Load(cls, Model, path):
        """Rebuild a partition saved by Save() for the same Topology, without re-running its constructor."""
        self = cls.__new__(cls)
        Spatial_Partition.__init__(self, Model)
        check(_lib.lib().hare_part_load(self.Model[0]._h, os.fsencode(path), C.byref(self._h)), "hare_part_load")
        kind = {1: "Voxel_Grid", 2: "Octree", 3: "KDTree"}[_lib.lib().hare_part_kind(self._h)]
        if kind != cls.__name__:
            h, self._h = self._h, C.c_void_p()
            _lib.lib().hare_part_destroy(h)
            raise ValueError(f"{path} holds a {kind}, not a {cls.__name__}")
        if hasattr(self, "_after_load"):
            self._after_load()
        return self

    # ---- reference-shaped single-ray overloads ------------------------------------------
    def Shoot(self, R, top_index=0, poly_origin1=None, poly_origin2=-1):
        """bool Shoot(Ray R, int top_index, out X_Event[, int poly_origin1, int poly_origin2 = -1]).
        Returns (hit, X_Event).  A ray the reference would throw on raises IndexError."""
        if top_index != 0:
            raise IndexError("top_index")
        o1 = None if poly_origin1 is None else np.array([poly_origin1], np.int32)
        o2 = None if poly_origin1 is None else np.array([poly_origin2], np.int32)
        r = self.Shoot_Batch(np.array([[R.x, R.y, R.z]]), np.array([[R.dx, R.dy, R.dz]]), o1, o2, np.array([R.Ray_ID], np.int32), moved=True)
        R.x, R.y, R.z = (float(v) for v in r["o"][0])
        pid = int(r["poly_id"][0])
        if pid == -2:
            raise IndexError("Index was outside the bounds of the array.")   # Voxels[X, Y, Z]
        if pid < 0:
            return False, X_Event()
        return True, X_Event(Point(*map(float, r["xyz"][0])), float(r["uv"][0, 0]), float(r["uv"][0, 1]), float(r["t"][0]), pid)

    # ---- new batched overload --------------------------------------------------------------
    def Shoot_Batch(self, o, d, poly_origin1=None, poly_origin2=None, Ray_ID=None, counters=False, moved=False):
        """bool[] Shoot(Ray[] R, 0, out X_Event[] events, int[] poly_origin1 = null, int[] poly_origin2 = null).
        o, d: (N,3).  Returns dict(hit, poly_id, t, xyz, uv[, o][, counters])."""
        o = as_f64(o).reshape(-1, 3); d = as_f64(d).reshape(-1, 3)
        N = o.shape[0]
        o1, o2, rid = as_i32(poly_origin1), as_i32(poly_origin2), as_i32(Ray_ID)
        t = np.empty(N); xyz = np.empty((N, 3)); pid = np.empty(N, np.int32); uv = np.empty((N, 2))
        om = np.empty((N, 3)) if moved else None
        cnt = np.zeros(4, np.uint64) if counters else None
        check(_lib.lib().hare_shoot_batch(self._h, ptr(o), ptr(d), ptr(o1), ptr(o2), ptr(rid), N, ptr(t), ptr(xyz), ptr(pid), ptr(uv), ptr(om), ptr(cnt)),
              "hare_shoot_batch")
        out = dict(hit=pid >= 0, poly_id=pid, t=t, xyz=xyz, uv=uv)
        if moved:
            out["o"] = om
        if counters:
            out["counters"] = cnt
        return out

    def Reflect_Chain(self, o, d, order, events=True, counters=False, points=False):
        """Specular chains of `order` Shoots kept on the device (harness op).  events: per-bounce Poly_id and t;
        points: also per-bounce X_Point (N, order, 3) and u, v (N, order, 2) -- the rest of every bounce's X_Event."""
        o = as_f64(o).reshape(-1, 3); d = as_f64(d).reshape(-1, 3)
        N = o.shape[0]
        evp = np.empty((N, order), np.int32) if events else None
        evt = np.empty((N, order)) if events else None
        evx = np.empty((N, order, 3)) if points else None
        evu = np.empty((N, order, 2)) if points else None
        fo = np.empty((N, 3)); fd = np.empty((N, 3)); ns = np.empty(N, np.int32)
        tot = C.c_uint64(); cnt = np.zeros(4, np.uint64) if counters else None
        check(_lib.lib().hare_reflect_chain_events(self._h, ptr(o), ptr(d), N, order, ptr(evp), ptr(evt), ptr(evx), ptr(evu), ptr(fo), ptr(fd), ptr(ns),
                                                   C.byref(tot), ptr(cnt)), "hare_reflect_chain_events")
        out = dict(ev_poly_id=evp, ev_t=evt, o=fo, d=fd, nshots=ns, total_shots=tot.value)
        if points:
            out["ev_xyz"] = evx; out["ev_uv"] = evu
        if counters:
            out["counters"] = cnt
        return out

    @property
    def device_bytes(self):
        return _lib.lib().hare_part_device_bytes(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.hare_part_destroy(self._h)
            self._h = None


class Voxel_Grid(Spatial_Partition):
    """new Voxel_Grid(Topology[] Model_in, int Domain): cell lists built on the GPU.
    Voxel_Grid.from_lists(...) uploads host-built lists (e.g. Hare's hierarchical ctor)."""
    has_uv = False

    def __init__(self, Model_in, Domain, Avg_polys=None):
        """Voxel_Grid(Model, Domain)  or, with a third argument, Voxel_Grid(Model, MaxDomain, Avg_polys) (Voxel_Grid.cs:128)."""
        super().__init__(Model_in)
        if Avg_polys is None:
            check(_lib.lib().hare_voxelgrid_build(self.Model[0]._h, int(Domain), C.byref(self._h)), "hare_voxelgrid_build")
        else:
            check(_lib.lib().hare_voxelgrid_build_adaptive(self.Model[0]._h, int(Domain), int(Avg_polys), C.byref(self._h)), "hare_voxelgrid_build_adaptive")
        self._post()

    @classmethod
    def from_lists(cls, Model_in, obox, ct, cell_offset, cell_poly):
        self = cls.__new__(cls)
        Spatial_Partition.__init__(self, Model_in)
        obox = as_f64(obox, (6,)); ct = as_i32(ct)
        off = np.ascontiguousarray(cell_offset, np.uint32); pol = np.ascontiguousarray(cell_poly, np.uint32)
        check(_lib.lib().hare_voxelgrid_upload(self.Model[0]._h, ptr(obox), ptr(ct), ptr(off), ptr(pol), C.byref(self._h)), "hare_voxelgrid_upload")
        self._post()
        return self

    def _after_load(self):
        self._post()

    def _post(self):
        obox, vd, ct, n = self.info()
        self.Char_Step = float(min(vd[0], vd[2]) if vd[0] < vd[1] else min(vd[1], vd[2]))   # Voxel_Grid.cs:90

    def info(self):
        obox = np.empty(6); vd = np.empty(3); ct = np.empty(3, np.int32); n = C.c_int64()
        check(_lib.lib().hare_voxelgrid_info(self._h, ptr(obox), ptr(vd), ptr(ct), C.byref(n)), "hare_voxelgrid_info")
        return obox, vd, ct, n.value

    def build_ms(self):
        """(device time of the build kernels, host wall time of the constructor) in ms."""
        k, w = C.c_double(), C.c_double()
        check(_lib.lib().hare_part_build_ms(self._h, C.byref(k), C.byref(w)), "hare_part_build_ms")
        return k.value, w.value

    def csr(self):
        _, _, ct, n = self.info()
        off = np.empty(int(ct[0]) * int(ct[1]) * int(ct[2]) + 1, np.uint32); pol = np.empty(max(n, 1), np.uint32)
        check(_lib.lib().hare_voxelgrid_download(self._h, ptr(off), ptr(pol)), "hare_voxelgrid_download")
        return off, pol[:n]

    # ---- the rest of Voxel_Grid's public surface (Voxel_Grid.cs:256-267, 322-332, 763-791); host-side integer helpers
    def VoxelCode(self, X, Y, Z):
        _, _, ct, _ = self.info()
        return int(ct[0]) * int(ct[1]) * Z + int(ct[1]) * X + Y          # XYTot * Z + VoxelCtY * X + Y

    def VoxelDecode(self, Code):
        _, _, ct, _ = self.info()
        xytot = int(ct[0]) * int(ct[1])
        Z = Code // xytot; Code -= Z * xytot
        Y = Code // int(ct[1]); X = Code - Y * int(ct[1])
        return X, Y, Z

    def PointInVoxel(self, Pt):
        import math
        obox, vd, _, _ = self.info()
        p = (Pt.x, Pt.y, Pt.z) if hasattr(Pt, "x") else tuple(Pt)
        return tuple(int(math.floor((p[a] - obox[a]) / vd[a])) for a in range(3))

    Xdim = property(lambda s: float(s.info()[0][3] - s.info()[0][0]))
    Ydim = property(lambda s: float(s.info()[0][4] - s.info()[0][1]))
    Zdim = property(lambda s: float(s.info()[0][5] - s.info()[0][2]))
    MinPt = property(lambda s: Point(*map(float, s.info()[0][:3])))


class Octree(Spatial_Partition):
    """new Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)."""

    def __init__(self, Model_In, maxDepth, maxPolygonsPerNode):
        super().__init__(Model_In)
        check(_lib.lib().hare_octree_build(self.Model[0]._h, int(maxDepth), int(maxPolygonsPerNode), C.byref(self._h)), "hare_octree_build")

    @classmethod
    def from_nodes(cls, Model_In, box, first_child, list_off, list_cnt, polys):
        self = cls.__new__(cls)
        Spatial_Partition.__init__(self, Model_In)
        box = as_f64(box).reshape(-1, 6); fc = as_i32(first_child)
        lo = np.ascontiguousarray(list_off, np.uint32); lc = np.ascontiguousarray(list_cnt, np.uint32); pol = np.ascontiguousarray(polys, np.uint32)
        check(_lib.lib().hare_octree_upload(self.Model[0]._h, ptr(box), ptr(fc), ptr(lo), ptr(lc), ptr(pol), box.shape[0], pol.shape[0], C.byref(self._h)),
              "hare_octree_upload")
        return self

    def info(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        check(_lib.lib().hare_octree_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "hare_octree_info")
        return dict(nodes=a.value, list_entries=b.value, lost=c.value, depth=d.value)

    def arrays(self):
        i = self.info(); n, l = i["nodes"], i["list_entries"]
        box = np.empty((n, 6)); fc = np.empty(n, np.int32); lo = np.empty(n, np.uint32); lc = np.empty(n, np.uint32); pol = np.empty(max(l, 1), np.uint32)
        check(_lib.lib().hare_octree_download(self._h, ptr(box), ptr(fc), ptr(lo), ptr(lc), ptr(pol)), "hare_octree_download")
        return box, fc, lo, lc, pol[:l]


class KDTree(Spatial_Partition):
    """new KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)."""

    def __init__(self, Model_In, maxDepth, maxPolygonsPerNode):
        super().__init__(Model_In)
        check(_lib.lib().hare_kdtree_build(self.Model[0]._h, int(maxDepth), int(maxPolygonsPerNode), C.byref(self._h)), "hare_kdtree_build")

    @classmethod
    def from_nodes(cls, Model_In, box, split, axis, left, list_off, list_cnt, polys):
        self = cls.__new__(cls)
        Spatial_Partition.__init__(self, Model_In)
        box = as_f64(box).reshape(-1, 6); split = as_f64(split); axis = as_i32(axis); left = as_i32(left)
        lo = np.ascontiguousarray(list_off, np.uint32); lc = np.ascontiguousarray(list_cnt, np.uint32); pol = np.ascontiguousarray(polys, np.uint32)
        check(_lib.lib().hare_kdtree_upload(self.Model[0]._h, ptr(box), ptr(split), ptr(axis), ptr(left), ptr(lo), ptr(lc), ptr(pol),
                                            box.shape[0], pol.shape[0], C.byref(self._h)), "hare_kdtree_upload")
        return self

    def info(self):
        a, b, d = C.c_int64(), C.c_int64(), C.c_int32()
        check(_lib.lib().hare_kdtree_info(self._h, C.byref(a), C.byref(b), C.byref(d)), "hare_kdtree_info")
        return dict(nodes=a.value, list_entries=b.value, depth=d.value)

    def arrays(self):
        i = self.info(); n, l = i["nodes"], i["list_entries"]
        box = np.empty((n, 6)); sp = np.empty(n); ax = np.empty(n, np.int32); le = np.empty(n, np.int32)
        lo = np.empty(n, np.uint32); lc = np.empty(n, np.uint32); pol = np.empty(max(l, 1), np.uint32)
        check(_lib.lib().hare_kdtree_download(self._h, ptr(box), ptr(sp), ptr(ax), ptr(le), ptr(lo), ptr(lc), ptr(pol)), "hare_kdtree_download")
        return box, sp, ax, le, lo, lc, pol[:l]
