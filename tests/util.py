"""Shared helpers for the parity tests."""
import numpy as np


def canon_octree(box, fc, lo, lc, pol):
    """Numbering-independent form of an octree: DFS in child order -> list of (depth, box, leaf list)."""
    out = []
    st = [(0, 0)]
    while st:
        n, dep = st.pop()
        if fc[n] < 0:
            out.append((dep, tuple(box[n]), tuple(pol[lo[n]:lo[n] + lc[n]].tolist())))
        else:
            out.append((dep, tuple(box[n]), None))
            for i in range(7, -1, -1):
                st.append((fc[n] + i, dep + 1))
    return out


def canon_kdtree(box, split, axis, left, right, lo, lc, pol):
    out = []
    st = [(0, 0)]
    while st:
        n, dep = st.pop()
        if left[n] < 0:
            out.append((dep, tuple(box[n]), tuple(pol[lo[n]:lo[n] + lc[n]].tolist())))
        else:
            out.append((dep, tuple(box[n]), (int(axis[n]), float(split[n]))))
            st.append((right[n], dep + 1))
            st.append((left[n], dep + 1))
    return out


def assert_events_equal(got, ref, uv=True, what=""):
    """Bit-exact comparison of batched X_Events (hit flag, Poly_id, t, X_Point[, u, v])."""
    assert np.array_equal(got["poly_id"], ref["poly_id"]), f"{what}: poly_id mismatch at {np.nonzero(got['poly_id'] != ref['poly_id'])[0][:10]}"
    assert np.array_equal(got["t"], ref["t"]), f"{what}: t mismatch"
    assert np.array_equal(got["xyz"], ref["xyz"]), f"{what}: X_Point mismatch"
    if uv:
        assert np.array_equal(got["uv"], ref["uv"]), f"{what}: u,v mismatch"
