"""ctypes loader for hare_b200/libhare_b200.so (the C ABI in include/hare_b200.h).

The library is the product: there is no Python or CPU fallback.  A missing .so,
a missing symbol or a failing call raises.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
SO_PATH = os.environ.get("HARE_B200_LIB") or os.path.join(_PKG, "libhare_b200.so")   # override: tuning experiments only
HEADER = os.path.join(ROOT, "include", "hare_b200.h")

_lib = None


class HareError(RuntimeError):
    pass


def build(force=False):
    """Compile the extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_PKG, "csrc")
    if not force and os.path.exists(SO_PATH):
        newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir)
                     if f.endswith((".cu", ".cuh", ".cpp", ".hpp")))
        newest = max(newest, os.path.getmtime(HEADER))
        if os.path.getmtime(SO_PATH) >= newest:
            return SO_PATH
    subprocess.check_call(["make", "-C", src_dir, "-s", "-B"])
    return SO_PATH


def declared_symbols():
    """Every function include/hare_b200.h declares."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hare_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise HareError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(hare_b200 has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64
    pp = C.POINTER(vp)
    L.hare_version.restype = C.c_char_p
    L.hare_last_error.restype = C.c_char_p
    L.hare_device_count.restype = i32
    L.hare_init.argtypes = [vp, i32]
    L.hare_topology_ingest.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, C.POINTER(i64)]
    L.hare_topology_create.argtypes = [vp, vp, vp, i64, vp, pp]
    L.hare_topology_polygon_count.restype = i64
    L.hare_topology_polygon_count.argtypes = [vp]
    L.hare_topology_destroy.argtypes = [vp]
    L.hare_voxelgrid_build.argtypes = [vp, i32, pp]
    L.hare_voxelgrid_build_adaptive.argtypes = [vp, i32, i32, pp]
    L.hare_voxelgrid_upload.argtypes = [vp, vp, vp, vp, vp, pp]
    L.hare_voxelgrid_info.argtypes = [vp, vp, vp, vp, C.POINTER(i64)]
    L.hare_voxelgrid_download.argtypes = [vp, vp, vp]
    L.hare_octree_build.argtypes = [vp, i32, i32, pp]
    L.hare_octree_upload.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, pp]
    L.hare_octree_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.hare_octree_download.argtypes = [vp, vp, vp, vp, vp, vp]
    L.hare_kdtree_build.argtypes = [vp, i32, i32, pp]
    L.hare_kdtree_upload.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, pp]
    L.hare_kdtree_info.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    L.hare_kdtree_download.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.hare_part_save.argtypes = [vp, C.c_char_p]
    L.hare_part_load.argtypes = [vp, C.c_char_p, pp]
    L.hare_part_kind.argtypes = [vp]
    L.hare_part_build_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.hare_part_device_bytes.restype = i64
    L.hare_part_device_bytes.argtypes = [vp]
    L.hare_part_destroy.argtypes = [vp]
    L.hare_shoot_batch.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.hare_shoot_batch_device.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp]
    L.hare_reflect_chain.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp]
    L.hare_reflect_chain_device.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.hare_reflect_chain_events.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.hare_reflect_chain_events_device.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.hare_launch_count.restype = u64
    sz = C.c_size_t
    L.hare_host_alloc.argtypes = [sz, pp]
    L.hare_host_free.argtypes = [vp]
    L.hare_host_register.argtypes = [vp, sz]
    L.hare_host_unregister.argtypes = [vp]
    L.hare_host_is_pinned.argtypes = [vp]
    L.hare_device_alloc.argtypes = [i32, sz, pp]
    L.hare_device_free.argtypes = [i32, vp]
    L.hare_device_memcpy.argtypes = [vp, vp, sz, i32, i32]
    L.hare_ipc_export.argtypes = [i32, vp, vp]
    L.hare_ipc_open.argtypes = [i32, vp, pp]
    L.hare_ipc_close.argtypes = [i32, vp]
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().hare_last_error().decode(errors="replace")
        if rc == -3:
            raise NotImplementedError(msg)
        raise HareError(f"{what} failed ({rc}): {msg}")


def ptr(a):
    """Host pointer of a numpy array (None -> NULL)."""
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def as_i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)
