"""ctypes driver of tests/emu/libkd_emu.so: the CPU replay of kd_wave.cuh's wavefront scheduler (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkd_emu.so")
_lib = None
PHASES = "SF N C T".split()


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libkd_emu.so"])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(topo_arrays, tree_arrays, o, d, origin1=None, origin2=None, ray_id=None, chain=False, order=1, slots=64, nmax=4, n_warps=4, tie_rule_on_tight_boxes=False):
    """topo_arrays = oracle Topology.arrays(); tree_arrays = oracle KDTree.arrays() = (box, split, axis, left, right, list_off, list_cnt, polys)."""
    verts, normals, vcount, _ = topo_arrays
    box, sp, ax, le, ri, lo, lc, pol = tree_arrays
    internal = le >= 0
    assert np.array_equal(ri[internal], le[internal] + 1), "the flattened kd-tree keeps Right = Left + 1"
    verts = np.ascontiguousarray(verts, np.float64); normals = np.ascontiguousarray(normals, np.float64); vcount = np.ascontiguousarray(vcount, np.int32)
    box = np.ascontiguousarray(box, np.float64); sp = np.ascontiguousarray(sp, np.float64)
    ax = np.ascontiguousarray(ax, np.int32); le = np.ascontiguousarray(le, np.int32)
    lo = np.ascontiguousarray(lo, np.uint32); lc = np.ascontiguousarray(lc, np.uint32)
    npol = len(pol)
    pol = np.ascontiguousarray(pol if npol else np.zeros(1), np.uint32)
    o = np.ascontiguousarray(o, np.float64).reshape(-1, 3); d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
    N = o.shape[0]
    o1 = None if origin1 is None else np.ascontiguousarray(origin1, np.int32)
    o2 = None if origin2 is None else np.ascontiguousarray(origin2, np.int32)
    rid = None if ray_id is None else np.ascontiguousarray(ray_id, np.int32)
    stats = np.zeros(16); counters = np.zeros(4, np.uint64)
    if chain:
        ev_pid = np.zeros((N, order), np.int32); ev_t = np.zeros((N, order)); fo = np.zeros((N, 3)); fd = np.zeros((N, 3))
        ns = np.zeros(N, np.int32); tot = np.zeros(1, np.uint64)
        ev_xyz = np.full((N, order, 3), np.nan); ev_uv = np.full((N, order, 2), np.nan)     # every row must be written
        args = [None] * 5 + [_p(ev_pid), _p(ev_t), _p(fo), _p(fd), _p(ns), _p(tot)]
        res = dict(ev_poly_id=ev_pid, ev_t=ev_t, ev_xyz=ev_xyz, ev_uv=ev_uv, o=fo, d=fd, nshots=ns, total=tot)
    else:
        t = np.zeros(N); xyz = np.zeros((N, 3)); pid = np.zeros(N, np.int32); uv = np.ones((N, 2)); om = np.zeros((N, 3))
        args = [_p(t), _p(xyz), _p(pid), _p(uv), _p(om)] + [None] * 6
        ev_xyz = ev_uv = None
        res = dict(t=t, xyz=xyz, poly_id=pid, uv=uv, o=om)
    rc = lib().kd_emu(_p(verts), _p(normals), _p(vcount), C.c_int64(len(vcount)), _p(box), _p(sp), _p(ax), _p(le), _p(lo), _p(lc), _p(pol),
                      C.c_int64(len(le)), C.c_int64(npol), _p(o), _p(d), _p(o1), _p(o2), _p(rid), C.c_int64(N), int(chain), int(order),
                      *args, int(slots), int(nmax), int(n_warps), int(tie_rule_on_tight_boxes), _p(stats), _p(counters), _p(ev_xyz), _p(ev_uv))
    if rc != 0:
        raise ValueError("kd_emu: unsupported (slots, nmax)")
    res["stats"] = dict(exec=dict(zip(PHASES, stats[0:4])), lanes=dict(zip(PHASES, stats[4:8])), trips=stats[8])
    res["counters"] = counters
    return res
