"""ctypes driver of tests/emu/liboct_emu.so: the CPU replay of oct_wave.cuh's wavefront scheduler (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboct_emu.so")
_lib = None
PHASES = "SF N G C T".split()


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboct_emu.so"])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def run(topo_arrays, tree_arrays, o, d, origin1=None, origin2=None, chain=False, order=1, slots=64, nmax=4, n_warps=4, regular=True, ray_steps=False):
    """topo_arrays = oracle Topology.arrays(); tree_arrays = oracle Octree.arrays() = (box, first_child, list_off, list_cnt, polys)."""
    verts, normals, vcount, _ = topo_arrays
    box, fc, lo, lc, pol = tree_arrays
    verts = np.ascontiguousarray(verts, np.float64); normals = np.ascontiguousarray(normals, np.float64); vcount = np.ascontiguousarray(vcount, np.int32)
    box = np.ascontiguousarray(box, np.float64); fc = np.ascontiguousarray(fc, np.int32)
    lo = np.ascontiguousarray(lo, np.uint32); lc = np.ascontiguousarray(lc, np.uint32)
    npol = len(pol)
    pol = np.ascontiguousarray(pol if npol else np.zeros(1), np.uint32)
    o = np.ascontiguousarray(o, np.float64).reshape(-1, 3); d = np.ascontiguousarray(d, np.float64).reshape(-1, 3)
    N = o.shape[0]
    o1 = None if origin1 is None else np.ascontiguousarray(origin1, np.int32)
    o2 = None if origin2 is None else np.ascontiguousarray(origin2, np.int32)
    stats = np.zeros(16); counters = np.zeros(4, np.uint64)
    steps = np.zeros(N, np.uint32) if ray_steps else None   # per ray: phase executions it took part in (its dependent chain)
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
    rc = lib().oct_emu(_p(verts), _p(normals), _p(vcount), C.c_int64(len(vcount)), _p(box), _p(fc), _p(lo), _p(lc), _p(pol),
                       C.c_int64(len(fc)), C.c_int64(npol), _p(o), _p(d), _p(o1), _p(o2), C.c_int64(N), int(chain), int(order),
                       *args, int(slots), int(nmax), int(n_warps), int(regular), _p(stats), _p(counters), _p(steps), _p(ev_xyz), _p(ev_uv))
    if rc != 0:
        raise ValueError("oct_emu: unsupported (slots, nmax)")
    res["stats"] = dict(exec=dict(zip(PHASES, stats[0:5])), lanes=dict(zip(PHASES, stats[5:10])), trips=stats[10])
    res["counters"] = counters
    if ray_steps:
        res["ray_steps"] = steps
    return res
