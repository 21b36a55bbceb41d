"""Regenerates tests/golden/*.npz with the pure-Python restatement (oracle/hare_oracle_py.py).

    python tests/golden/make_golden.py

The reference (C#) cannot run in this image, so these vectors are produced by the second,
independent restatement; the C++ oracle (CPU tests) and the CUDA path (-m gpu tests) are both
checked against them.  Inputs are stored next to the outputs, so the fixtures do not depend on
the generators staying unchanged.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from hare_b200.harness import meshes, rays_from_sources  # noqa: E402
from oracle import hare_oracle_py as hp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def topo_of(mesh):
    T = hp.Topology(tuple(mesh.minpt), tuple(mesh.maxpt))
    for i in range(mesh.P):
        T.Add_Polygon([tuple(map(float, mesh.verts[i, k])) for k in range(mesh.vcount[i])])
    T.Finish_Topology()
    return T


def shoot_all(part, o, d, o1=None, o2=None):
    n = len(o)
    pid = np.full(n, -1, np.int32); t = np.zeros(n); xyz = np.zeros((n, 3)); uv = np.zeros((n, 2)); om = np.zeros((n, 3))
    for i in range(n):
        R = hp.Ray(*map(float, o[i]), *map(float, d[i]), Ray_ID=i + 1)
        try:
            hit, ev = part.Shoot(R, 0, -1 if o1 is None else int(o1[i]), -1 if o2 is None else int(o2[i]))
        except IndexError:
            pid[i] = -2; om[i] = (R.x, R.y, R.z)
            continue
        om[i] = (R.x, R.y, R.z)
        if hit:
            pid[i], t[i], xyz[i], uv[i] = ev.Poly_id, ev.t, ev.X_Point, (ev.u, ev.v)
    return dict(poly_id=pid, t=t, xyz=xyz, uv=uv, o_moved=om)


def main():
    out = {}
    cases = [("shoebox", meshes.shoebox(), np.array([[5.0, 3.5, 1.5]]), 10, (3, 2), (4, 1), 600),
             ("tiny", meshes.hall("tiny"), meshes.sources(4), 6, (3, 8), (6, 6), 400)]
    for name, mesh, src, dom, octa, kda, n in cases:
        T = topo_of(mesh)
        o, d = rays_from_sources(n, src, stream=11)
        # a few rays from outside the grid and a few special directions
        o[-40:] = np.array([5.0, 3.5, 1.5]) - 25.0 * d[-40:] + 4.0 * np.roll(d[-40:], 1, axis=0)
        d[:6] = np.array([[1, 0, 0], [0, -1, 0], [0, 0, 1], [0.6, -0.0, 0.8], [-0.0, 0.6, 0.8], [0.0, 0.8, -0.6]])
        out[f"{name}_verts"] = mesh.verts; out[f"{name}_vcount"] = mesh.vcount
        out[f"{name}_minpt"] = mesh.minpt; out[f"{name}_maxpt"] = mesh.maxpt
        out[f"{name}_o"] = o; out[f"{name}_d"] = d
        out[f"{name}_args"] = np.array([dom, *octa, *kda], np.int32)
        out[f"{name}_topo_verts"] = np.array([[list(p) for p in (poly[0] + (poly[0][-1],) * (4 - poly[2]))] for poly in T.Polys])
        out[f"{name}_topo_normals"] = np.array([poly[1] for poly in T.Polys])
        out[f"{name}_topo_minmax"] = np.array(list(T.Min) + list(T.Max))
        vg = hp.Voxel_Grid([T], dom)
        off = [0]; pol = []
        for x in range(dom):
            for y in range(dom):
                for z in range(dom):
                    pol += vg.Voxel_Inv[(x, y, z)]; off.append(len(pol))
        out[f"{name}_vg_offset"] = np.array(off, np.uint32); out[f"{name}_vg_polys"] = np.array(pol, np.uint32)
        for k, v in shoot_all(vg, o, d).items():
            out[f"{name}_vg_{k}"] = v
        first = out[f"{name}_vg_poly_id"].copy()
        vg.mail = [0] * T.Polygon_Count
        for k, v in shoot_all(vg, o, d, first, np.roll(first, 1)).items():
            out[f"{name}_vgo_{k}"] = v
        for k, v in shoot_all(hp.Octree([T], *octa), o, d).items():
            out[f"{name}_oct_{k}"] = v
        for k, v in shoot_all(hp.KDTree([T], *kda), o, d).items():
            out[f"{name}_kd_{k}"] = v
        print(name, "done:", mesh.P, "polygons,", n, "rays; hits vg/oct/kd =",
              int((out[f"{name}_vg_poly_id"] >= 0).sum()), int((out[f"{name}_oct_poly_id"] >= 0).sum()), int((out[f"{name}_kd_poly_id"] >= 0).sum()))
    np.savez_compressed(os.path.join(HERE, "hare_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "hare_golden.npz"))


if __name__ == "__main__":
    main()
