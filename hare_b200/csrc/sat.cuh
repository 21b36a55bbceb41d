// sat.cuh -- polygon / axis-aligned-box overlap predicate used by the grid build kernel
// (device) and by the host octree builder.
//
// Follows AABB.PolyBoxOverlap (AABB_Tri_Int.cs:165-260): fan-triangulate the polygon from
// vertex 0; per triangle translate by the box centre, run the nine edge-cross axis tests
// in the reference's order and with the reference's vertex pairs, then the three box axes,
// then the plane test (planeBoxOverlap :51-95).  Same arithmetic, same order, no FMA, so the
// predicate flips on exactly the same inputs as the C# code.
#pragma once
#include "hare_math.cuh"

namespace hare {

struct Box3 {
    double cx, cy, cz;   // AABB.Center    = (Max + Min) / 2          AABB_Main.cs:64
    double hx, hy, hz;   // AABB.halfwidth = (Max - Min) / 2          AABB_Main.cs:65-67
};

HD Box3 make_box(double mnx, double mny, double mnz, double mxx, double mxy, double mxz) {
    Box3 b;
    b.cx = (mxx + mnx) / 2; b.cy = (mxy + mny) / 2; b.cz = (mxz + mnz) / 2;
    b.hx = (mxx - mnx) / 2; b.hy = (mxy - mny) / 2; b.hz = (mxz - mnz) / 2;
    return b;
}

// one edge-cross axis: projections pa, pb of the two vertices that matter, radius r
HD bool sat_axis_separates(double pa, double pb, double r) {
    double lo, hi;
    if (pa < pb) { lo = pa; hi = pb; } else { lo = pb; hi = pa; }
    return (lo > r) || (hi < -r);
}

HD bool tri_box_overlap(const Box3& B, const double* A0, const double* A1, const double* A2) {
    const double v0x = A0[0] - B.cx, v0y = A0[1] - B.cy, v0z = A0[2] - B.cz;
    const double v1x = A1[0] - B.cx, v1y = A1[1] - B.cy, v1z = A1[2] - B.cz;
    const double v2x = A2[0] - B.cx, v2y = A2[1] - B.cy, v2z = A2[2] - B.cz;
    const double e0x = v1x - v0x, e0y = v1y - v0y, e0z = v1z - v0z;
    const double e1x = v2x - v1x, e1y = v2y - v1y, e1z = v2z - v1z;
    const double e2x = v0x - v2x, e2y = v0y - v2y, e2z = v0z - v2z;
    double fx, fy, fz;

    // edge 0: X01, Y02, Z12
    fx = fabs(e0x); fy = fabs(e0y); fz = fabs(e0z);
    if (sat_axis_separates(e0z * v0y - e0y * v0z, e0z * v2y - e0y * v2z, fz * B.hy + fy * B.hz)) return false;
    if (sat_axis_separates(-e0z * v0x + e0x * v0z, -e0z * v2x + e0x * v2z, fz * B.hx + fx * B.hz)) return false;
    if (sat_axis_separates(e0y * v2x - e0x * v2y, e0y * v1x - e0x * v1y, fy * B.hx + fx * B.hy)) return false;
    // edge 1: X01, Y02, Z0
    fx = fabs(e1x); fy = fabs(e1y); fz = fabs(e1z);
    if (sat_axis_separates(e1z * v0y - e1y * v0z, e1z * v2y - e1y * v2z, fz * B.hy + fy * B.hz)) return false;
    if (sat_axis_separates(-e1z * v0x + e1x * v0z, -e1z * v2x + e1x * v2z, fz * B.hx + fx * B.hz)) return false;
    if (sat_axis_separates(e1y * v0x - e1x * v0y, e1y * v1x - e1x * v1y, fy * B.hx + fx * B.hy)) return false;
    // edge 2: X2, Y1, Z12
    fx = fabs(e2x); fy = fabs(e2y); fz = fabs(e2z);
    if (sat_axis_separates(e2z * v0y - e2y * v0z, e2z * v1y - e2y * v1z, fz * B.hy + fy * B.hz)) return false;
    if (sat_axis_separates(-e2z * v0x + e2x * v0z, -e2z * v1x + e2x * v1z, fz * B.hx + fx * B.hz)) return false;
    if (sat_axis_separates(e2y * v2x - e2x * v2y, e2y * v1x - e2x * v1y, fy * B.hx + fx * B.hy)) return false;

    // box axes
    double lo, hi;
    lo = fmin(v0x, fmin(v1x, v2x)); hi = fmax(v0x, fmax(v1x, v2x));
    if (lo > B.hx || hi < -B.hx) return false;
    lo = fmin(v0y, fmin(v1y, v2y)); hi = fmax(v0y, fmax(v1y, v2y));
    if (lo > B.hy || hi < -B.hy) return false;
    lo = fmin(v0z, fmin(v1z, v2z)); hi = fmax(v0z, fmax(v1z, v2z));
    if (lo > B.hz || hi < -B.hz) return false;

    // triangle plane: normal = Cross(e0, e1)  (Hare_Geometry_Math.cs:70-73)
    const double nx = e0y * e1z - e0z * e1y, ny = -(e0x * e1z - e0z * e1x), nz = e0x * e1y - e0y * e1x;
    double ax, ay, az, bx, by, bz;   // vmin, vmax
    if (nx > 0.0) { ax = -B.hx - v0x; bx = B.hx - v0x; } else { ax = B.hx - v0x; bx = -B.hx - v0x; }
    if (ny > 0.0) { ay = -B.hy - v0y; by = B.hy - v0y; } else { ay = B.hy - v0y; by = -B.hy - v0y; }
    if (nz > 0.0) { az = -B.hz - v0z; bz = B.hz - v0z; } else { az = B.hz - v0z; bz = -B.hz - v0z; }
    if (dot3(nx, ny, nz, ax, ay, az) > 0.0) return false;
    return dot3(nx, ny, nz, bx, by, bz) >= 0.0;
}

// P = PolyRec doubles (vertices at 0,3,6,9); n = 3 or 4.
HD bool poly_box_overlap(const Box3& B, const double* P, int n) {
    if (tri_box_overlap(B, P, P + 3, P + 6)) return true;
    if (n == 4 && tri_box_overlap(B, P, P + 6, P + 9)) return true;
    return false;
}

}  // namespace hare
