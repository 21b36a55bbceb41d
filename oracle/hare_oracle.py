"""ctypes binding of oracle/libhare_oracle.so (the C++ restatement of Hare's CPU path).

TEST INFRASTRUCTURE ONLY -- see the header of hare_oracle.cpp.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by hare_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhare_oracle.so")
_lib = None

_d = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build(force=False):
    src = os.path.join(_HERE, "hare_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libhare_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.ho_topology_new.restype = vp
        L.ho_topology_new.argtypes = [_d, _d]
        L.ho_topology_add_polygon.argtypes = [vp, _d, i32]
        L.ho_topology_add_polygons.argtypes = [vp, _d, _i32, i64]
        L.ho_topology_finish.argtypes = [vp]
        L.ho_topology_polygon_count.restype = i64
        L.ho_topology_polygon_count.argtypes = [vp]
        L.ho_topology_vertex_count.restype = i64
        L.ho_topology_vertex_count.argtypes = [vp]
        L.ho_topology_get.argtypes = [vp, vp, vp, vp, vp]
        L.ho_topology_free.argtypes = [vp]
        L.ho_voxelgrid_new.restype = vp
        L.ho_voxelgrid_new.argtypes = [vp, i32, i32, i32, i32]
        L.ho_voxelgrid_info.argtypes = [vp, _d, _d, _i32, C.POINTER(i64)]
        L.ho_voxelgrid_csr.argtypes = [vp, _u32, _u32]
        L.ho_octree_new.restype = vp
        L.ho_octree_new.argtypes = [vp, i32, i32]
        L.ho_octree_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        L.ho_octree_get.argtypes = [vp, _d, _i32, _u32, _u32, _u32]
        L.ho_kdtree_new.restype = vp
        L.ho_kdtree_new.argtypes = [vp, i32, i32]
        L.ho_kdtree_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
        L.ho_kdtree_get.argtypes = [vp, _d, _d, _i32, _i32, _i32, _u32, _u32, _u32]
        L.ho_partition_free.argtypes = [vp]
        L.ho_shoot.argtypes = [vp, i64, _d, _d, vp, vp, vp, _d, _d, _i32, vp, vp, i32]
        L.ho_reflect_chain.argtypes = [vp, i64, _d, _d, i32, vp, vp, vp, vp, vp, vp, i32]
        L.ho_reflect_chain_events.argtypes = [vp, i64, _d, _d, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32]
        L.ho_poly_box_overlap.argtypes = [_d, _d, _d, i32]
        L.ho_round15.restype = dbl
        L.ho_round15.argtypes = [dbl]
        L.ho_intersect.argtypes = [vp, i64, _d, _d, i32, _d]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Topology:
    """Hare.Geometry.Topology(minpt, maxpt) + Add_Polygon + Finish_Topology()."""

    def __init__(self, minpt, maxpt):
        self._h = lib().ho_topology_new(np.ascontiguousarray(minpt, np.float64), np.ascontiguousarray(maxpt, np.float64))

    @classmethod
    def from_mesh(cls, mesh):
        t = cls(mesh.minpt, mesh.maxpt)
        if lib().ho_topology_add_polygons(t._h, np.ascontiguousarray(mesh.verts, np.float64).reshape(-1), np.ascontiguousarray(mesh.vcount, np.int32), mesh.P) != 0:
            raise NotImplementedError("Hare Does not yet support polygons of more than 4 sides.")
        t.Finish_Topology()
        return t

    def Add_Polygon(self, pts):
        pts = np.ascontiguousarray(pts, np.float64)
        if lib().ho_topology_add_polygon(self._h, pts.reshape(-1), pts.shape[0]) != 0:
            raise NotImplementedError("Hare Does not yet support polygons of more than 4 sides.")

    def Finish_Topology(self):
        lib().ho_topology_finish(self._h)

    @property
    def Polygon_Count(self):
        return lib().ho_topology_polygon_count(self._h)

    @property
    def Vertex_Count(self):
        return lib().ho_topology_vertex_count(self._h)

    def arrays(self):
        P = self.Polygon_Count
        verts = np.empty((P, 4, 3)); normals = np.empty((P, 3)); vcount = np.empty(P, np.int32); mm = np.empty(6)
        lib().ho_topology_get(self._h, _p(verts), _p(normals), _p(vcount), _p(mm))
        return verts, normals, vcount, mm

    def intersect(self, i, o, d, slow=False):
        out = np.zeros(6)
        h = lib().ho_intersect(self._h, i, np.ascontiguousarray(o, np.float64), np.ascontiguousarray(d, np.float64), int(slow), out)
        return bool(h), out

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ho_topology_free(self._h)
            self._h = None


class _Partition:
    uv = True

    def __init__(self, topo):
        self.topo = topo      # keep alive
        self._h = None

    def Shoot(self, o, d, origin1=None, origin2=None, ray_id=None, nthreads=1):
        """Batched Shoot.  Returns dict(t, xyz, poly_id, uv, o (moved origins), counters)."""
        o = np.array(o, dtype=np.float64, order="C").reshape(-1, 3)   # copy: origins are IN/OUT
        d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
        N = o.shape[0]
        t = np.zeros(N); xyz = np.zeros((N, 3)); pid = np.zeros(N, np.int32); uv = np.zeros((N, 2)); cnt = np.zeros(4, np.uint64)
        o1 = None if origin1 is None else np.ascontiguousarray(origin1, np.int32)
        o2 = None if origin2 is None else np.ascontiguousarray(origin2, np.int32)
        rid = None if ray_id is None else np.ascontiguousarray(ray_id, np.int32)
        lib().ho_shoot(self._h, N, o.reshape(-1), d.reshape(-1), _p(o1), _p(o2), _p(rid), t, xyz.reshape(-1), pid, _p(uv), _p(cnt), nthreads)
        return dict(t=t, xyz=xyz, poly_id=pid, uv=uv, o=o, counters=cnt)

    def reflect_chain(self, o, d, order, events=True, nthreads=1, points=False):
        """points=True adds the per-bounce X_Point (N, order, 3) and u, v (N, order, 2) streams."""
        o = np.ascontiguousarray(o, np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
        N = o.shape[0]
        ev_pid = np.zeros((N, order), np.int32) if events else None
        ev_t = np.zeros((N, order)) if events else None
        fo = np.zeros((N, 3)); fd = np.zeros((N, 3)); nb = np.zeros(N, np.int32); cnt = np.zeros(4, np.uint64)
        if points:
            ev_xyz = np.full((N, order, 3), np.nan); ev_uv = np.full((N, order, 2), np.nan)
            lib().ho_reflect_chain_events(self._h, N, o.reshape(-1), d.reshape(-1), order, _p(ev_pid), _p(ev_t), _p(ev_xyz), _p(ev_uv), _p(fo), _p(fd),
                                          _p(nb), _p(cnt), nthreads)
            return dict(ev_poly_id=ev_pid, ev_t=ev_t, ev_xyz=ev_xyz, ev_uv=ev_uv, o=fo, d=fd, nshots=nb, counters=cnt)
        lib().ho_reflect_chain(self._h, N, o.reshape(-1), d.reshape(-1), order, _p(ev_pid), _p(ev_t), _p(fo), _p(fd), _p(nb), _p(cnt), nthreads)
        return dict(ev_poly_id=ev_pid, ev_t=ev_t, o=fo, d=fd, nshots=nb, counters=cnt)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ho_partition_free(self._h)
            self._h = None


class Voxel_Grid(_Partition):
    """mode: 'flat' = Voxel_Grid(Model, Domain) literal O(D^3 P); 'fast' = same result, polygon-major;
    'hier' = Voxel_Grid(Model, MaxDomain, Avg_polys)."""

    def __init__(self, topo, domain, mode="fast", avg_polys=0, nthreads=1):
        super().__init__(topo)
        m = {"flat": 0, "fast": 1, "hier": 2}[mode]
        self._h = lib().ho_voxelgrid_new(topo._h, m, domain, avg_polys, nthreads)

    def info(self):
        obox = np.zeros(6); vd = np.zeros(3); ct = np.zeros(3, np.int32); n = C.c_int64()
        lib().ho_voxelgrid_info(self._h, obox, vd, ct, C.byref(n))
        return obox, vd, ct, n.value

    def csr(self):
        _, _, ct, n = self.info()
        off = np.zeros(int(ct[0]) * int(ct[1]) * int(ct[2]) + 1, np.uint32); pol = np.zeros(max(n, 1), np.uint32)
        lib().ho_voxelgrid_csr(self._h, off, pol)
        return off, pol[:n]


class Octree(_Partition):
    def __init__(self, topo, maxDepth, maxPolygonsPerNode):
        super().__init__(topo)
        self._h = lib().ho_octree_new(topo._h, maxDepth, maxPolygonsPerNode)

    def info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        lib().ho_octree_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def arrays(self):
        n, l, _ = self.info()
        box = np.zeros((n, 6)); fc = np.zeros(n, np.int32); lo = np.zeros(n, np.uint32); lc = np.zeros(n, np.uint32); pol = np.zeros(max(l, 1), np.uint32)
        lib().ho_octree_get(self._h, box.reshape(-1), fc, lo, lc, pol)
        return box, fc, lo, lc, pol[:l]


class KDTree(_Partition):
    def __init__(self, topo, maxDepth, maxPolygonsPerNode):
        super().__init__(topo)
        self._h = lib().ho_kdtree_new(topo._h, maxDepth, maxPolygonsPerNode)

    def info(self):
        a, b = C.c_int64(), C.c_int64()
        lib().ho_kdtree_info(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def arrays(self):
        n, l = self.info()
        box = np.zeros((n, 6)); sp = np.zeros(n); ax = np.zeros(n, np.int32); le = np.zeros(n, np.int32); ri = np.zeros(n, np.int32)
        lo = np.zeros(n, np.uint32); lc = np.zeros(n, np.uint32); pol = np.zeros(max(l, 1), np.uint32)
        lib().ho_kdtree_get(self._h, box.reshape(-1), sp, ax, le, ri, lo, lc, pol)
        return box, sp, ax, le, ri, lo, lc, pol[:l]


def poly_box_overlap(bmin, bmax, pts):
    pts = np.ascontiguousarray(pts, np.float64)
    return bool(lib().ho_poly_box_overlap(np.ascontiguousarray(bmin, np.float64), np.ascontiguousarray(bmax, np.float64), pts.reshape(-1), pts.shape[0]))


def round15(x):
    return lib().ho_round15(float(x))
