#!/usr/bin/env python
"""Companion of oracle/csharp/HareOracle.cs: dumps a case for the real C# reference and compares what it wrote.

    python tools/csharp_golden.py dump case.bin [--mesh tiny --kind 1 --arg0 6 --arg1 0 --rays 2000]
    python tools/csharp_golden.py compare case.bin case.out
"""
import argparse
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hare_b200.harness import meshes, rays_from_sources  # noqa: E402


def dump(a):
    mesh = meshes.shoebox() if a.mesh == "shoebox" else meshes.hall(a.mesh)
    src = np.array([[5.0, 3.5, 1.5]]) if a.mesh == "shoebox" else meshes.sources(4)
    o, d = rays_from_sources(a.rays, src, stream=21)
    with open(a.file, "wb") as f:
        f.write(struct.pack("<5i", mesh.P, a.rays, a.kind, a.arg0, a.arg1))
        f.write(np.asarray(mesh.minpt, "<f8").tobytes()); f.write(np.asarray(mesh.maxpt, "<f8").tobytes())
        for i in range(mesh.P):
            n = int(mesh.vcount[i])
            f.write(struct.pack("<i", n)); f.write(np.asarray(mesh.verts[i, :n], "<f8").tobytes())
        f.write(np.concatenate([o, d], axis=1).astype("<f8").tobytes())
        f.write(np.full((a.rays, 2), -1, "<i4").tobytes())
    print("wrote", a.file)


def compare(a):
    from oracle import hare_oracle as ho
    raw = open(a.file, "rb").read()
    P, N, kind, arg0, arg1 = struct.unpack_from("<5i", raw, 0)
    off = 20
    mn = np.frombuffer(raw, "<f8", 3, off); mx = np.frombuffer(raw, "<f8", 3, off + 24); off += 48
    T = ho.Topology(mn, mx)
    for _ in range(P):
        n, = struct.unpack_from("<i", raw, off); off += 4
        T.Add_Polygon(np.frombuffer(raw, "<f8", 3 * n, off).reshape(n, 3)); off += 24 * n
    T.Finish_Topology()
    od = np.frombuffer(raw, "<f8", 6 * N, off).reshape(N, 6)
    part = ho.Voxel_Grid(T, arg0, "flat") if kind == 1 else (ho.Octree(T, arg0, arg1) if kind == 2 else ho.KDTree(T, arg0, arg1))
    ref = part.Shoot(od[:, :3], od[:, 3:])
    rec = np.dtype([("pid", "<i4"), ("t", "<f8"), ("xyz", "<f8", 3), ("uv", "<f8", 2), ("o", "<f8", 3)])
    got = np.frombuffer(open(a.out, "rb").read(), rec)
    ok = (np.array_equal(got["pid"], ref["poly_id"]) and np.array_equal(got["t"], ref["t"]) and np.array_equal(got["xyz"], ref["xyz"])
          and np.array_equal(got["uv"], ref["uv"]) and np.array_equal(got["o"], ref["o"]))
    print("C# reference == C++ oracle, bit for bit:", ok)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["dump", "compare"]); ap.add_argument("file"); ap.add_argument("out", nargs="?")
    ap.add_argument("--mesh", default="tiny"); ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--arg0", type=int, default=6); ap.add_argument("--arg1", type=int, default=0); ap.add_argument("--rays", type=int, default=2000)
    a = ap.parse_args()
    dump(a) if a.cmd == "dump" else compare(a)
