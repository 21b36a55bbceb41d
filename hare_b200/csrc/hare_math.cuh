// hare_math.cuh -- FP64 primitives shared by host builders and sm_100a kernels.
//
// Everything here must produce the same bits as the reference's C# doubles:
// IEEE binary64, round-to-nearest, NO fused multiply-add (nvcc -fmad=false,
// host -ffp-contract=off), operands and operation order as in the cited source.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

namespace hare {

// 128-byte polygon record, 128-byte aligned: one cache line per polygon test.
//   v[0..11]  Polys[i].Points[0..3] xyz (a triangle repeats vertex 2 in slot 3)
//   v[12..14] Polys[i].Normal
//   v[15]     VertextCT as a double (3.0 or 4.0)
struct alignas(128) PolyRec { double v[16]; };

struct Ray3 { double x, y, z, dx, dy, dz; };

// .NET Math.Max / Math.Min (NaN-propagating; used at "Octree - alt.cs":182-183, AABB_Main.cs:199-200)
#if defined(__CUDA_ARCH__)
// device: DMNMX already orders -0 < +0 like .NET; only the NaN case needs the fix-up
HD double net_max(double a, double b) { const double r = fmax(a, b); return (a != a) ? a : ((b != b) ? b : r); }
HD double net_min(double a, double b) { const double r = fmin(a, b); return (a != a) ? a : ((b != b) ? b : r); }
#else
HD double net_max(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return (copysign(1.0, a) < 0) ? b : a;
    return a > b ? a : b;
}
HD double net_min(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return (copysign(1.0, a) < 0) ? a : b;
    return a < b ? a : b;
}
#endif

// (int)Math.Floor(x) as RyuJIT x64 evaluates it: cvttsd2si gives 0x80000000 for NaN/overflow.
HD int32_t floor_to_int(double x) {
    double f = floor(x);
    if (!(f >= -2147483648.0 && f < 2147483648.0)) return INT32_MIN;
    return (int32_t)f;
}

// Hare_math.Dot(6 doubles)  Hare_Geometry_Math.cs:43-46
HD double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return (ax * bx) + (ay * by) + (az * bz);
}

// Moller-Trumbore in Hare's unnormalised-determinant form.
//   SLOW = false: Polygon.RayXtri(ref Ray, ref Point x3, ref t)   Hare_Geometry_Polygons.cs:449-510
//   SLOW = true : Polygon.RayXtri(Ray, Point x3, ref t, ref u, ref v)   :385-435, whose cross products go
//                 through Hare_math.Cross (y = -(ax*bz - az*bx), Hare_Geometry_Math.cs:66-69) and which
//                 also scales u, v by 1/det.
// The divide is issued only after the u/v rejections: 1/det feeds nothing but t, u, v.
template <bool SLOW>
HD bool ray_x_tri(const Ray3& R, const double* a, const double* b, const double* c, double& t, double& u, double& v) {
    const double e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
    const double e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
    const double px = R.dy * e2z - R.dz * e2y;
    const double py = SLOW ? -(R.dx * e2z - R.dz * e2x) : (R.dz * e2x - R.dx * e2z);
    const double pz = R.dx * e2y - R.dy * e2x;
    const double det = dot3(e1x, e1y, e1z, px, py, pz);
    const double tx = R.x - a[0], ty = R.y - a[1], tz = R.z - a[2];
    const double qx = ty * e1z - tz * e1y;
    const double qy = SLOW ? -(tx * e1z - tz * e1x) : (tz * e1x - tx * e1z);
    const double qz = tx * e1y - ty * e1x;
    double uu, vv;
    if (det > 0.000001) {
        uu = dot3(tx, ty, tz, px, py, pz);
        if (SLOW) u = uu;
        if (uu < 0.0 || uu > det) return false;
        vv = dot3(R.dx, R.dy, R.dz, qx, qy, qz);
        if (SLOW) v = vv;
        if (vv < 0.0 || uu + vv > det) return false;
    } else if (det < -0.000001) {
        uu = dot3(tx, ty, tz, px, py, pz);
        if (SLOW) u = uu;
        if (uu > 0.0 || uu < det) return false;
        vv = dot3(R.dx, R.dy, R.dz, qx, qy, qz);
        if (SLOW) v = vv;
        if (vv > 0.0 || uu + vv < det) return false;
    } else {
        return false;
    }
    const double invdet = 1.0 / det;
    t = dot3(e2x, e2y, e2z, qx, qy, qz) * invdet;
    if (SLOW) { u = uu * invdet; v = vv * invdet; }
    return true;
}

// Same arithmetic as ray_x_tri<false>, arranged as one code path for a converged warp: the two
// determinant-sign branches of the reference (Hare_Geometry_Polygons.cs:483-505) differ only in the
// direction of four comparisons, so they are folded into selects.  Values are bit-identical.
HD bool ray_x_tri_fast1(const Ray3& R, double ax, double ay, double az, double bx, double by, double bz,
                        double cx, double cy, double cz, double& t) {
    const double e1x = bx - ax, e1y = by - ay, e1z = bz - az;
    const double e2x = cx - ax, e2y = cy - ay, e2z = cz - az;
    const double px = R.dy * e2z - R.dz * e2y;
    const double py = R.dz * e2x - R.dx * e2z;
    const double pz = R.dx * e2y - R.dy * e2x;
    const double det = dot3(e1x, e1y, e1z, px, py, pz);
    const bool pos = det > 0.000001, neg = det < -0.000001;
    if (!(pos || neg)) return false;
    const double tx = R.x - ax, ty = R.y - ay, tz = R.z - az;
    const double uu = dot3(tx, ty, tz, px, py, pz);
    if (pos ? (uu < 0.0 || uu > det) : (uu > 0.0 || uu < det)) return false;
    const double qx = ty * e1z - tz * e1y;
    const double qy = tz * e1x - tx * e1z;
    const double qz = tx * e1y - ty * e1x;
    const double vv = dot3(R.dx, R.dy, R.dz, qx, qy, qz);
    const double uv = uu + vv;
    if (pos ? (vv < 0.0 || uv > det) : (vv > 0.0 || uv < det)) return false;
    const double invdet = 1.0 / det;
    t = dot3(e2x, e2y, e2z, qx, qy, qz) * invdet;
    return true;
}

// The slow path (ray_x_tri<true>: Polygon.RayXtri(Ray, Point x3, ref t, ref u, ref v), Hare_Geometry_Polygons.cs:385-435) as one
// code path for a converged warp: cross products through Hare_math.Cross (y = -(ax*bz - az*bx), Hare_Geometry_Math.cs:66-69),
// u and v scaled by 1/det (:430-432).  Values are bit-identical to ray_x_tri<true>; u, v are written only on success.
HD bool ray_x_tri_slow1(const Ray3& R, double ax, double ay, double az, double bx, double by, double bz,
                        double cx, double cy, double cz, double& t, double& u, double& v) {
    const double e1x = bx - ax, e1y = by - ay, e1z = bz - az;
    const double e2x = cx - ax, e2y = cy - ay, e2z = cz - az;
    const double px = R.dy * e2z - R.dz * e2y;
    const double py = -(R.dx * e2z - R.dz * e2x);
    const double pz = R.dx * e2y - R.dy * e2x;
    const double det = dot3(e1x, e1y, e1z, px, py, pz);
    const bool pos = det > 0.000001, neg = det < -0.000001;
    if (!(pos || neg)) return false;
    const double tx = R.x - ax, ty = R.y - ay, tz = R.z - az;
    const double uu = dot3(tx, ty, tz, px, py, pz);
    if (pos ? (uu < 0.0 || uu > det) : (uu > 0.0 || uu < det)) return false;
    const double qx = ty * e1z - tz * e1y;
    const double qy = -(tx * e1z - tz * e1x);
    const double qz = tx * e1y - ty * e1x;
    const double vv = dot3(R.dx, R.dy, R.dz, qx, qy, qz);
    const double uv = uu + vv;
    if (pos ? (vv < 0.0 || uv > det) : (vv > 0.0 || uv < det)) return false;
    const double invdet = 1.0 / det;
    t = dot3(e2x, e2y, e2z, qx, qy, qz) * invdet;
    u = uu * invdet; v = vv * invdet;
    return true;
}

// Triangle.Intersect / Quadrilateral.Intersect  (Hare_Geometry_Polygons.cs:637-688, 731-823) with
// Polygon.Ray_Side (:601-606) choosing the winding.  P = the 16 doubles of a PolyRec.
// On success writes t (and u, v for SLOW); the caller forms X_Point = o + d*t.
template <bool SLOW>
HD bool poly_intersect(const double* P, const Ray3& R, double& t, double& u, double& v) {
    const double* P0 = P; const double* P1 = P + 3; const double* P2 = P + 6; const double* P3 = P + 9;
    const bool quad = (P[15] == 4.0);
    t = 0; if (SLOW) { u = 0; v = 0; }
    const double n = dot3(R.dx, R.dy, R.dz, P[12], P[13], P[14]);
    if (!(n < 0)) {   // Ray_Side true
        if (ray_x_tri<SLOW>(R, P0, P1, P2, t, u, v)) return true;
        if (quad && ray_x_tri<SLOW>(R, P2, P3, P0, t, u, v)) return true;
    } else {
        if (ray_x_tri<SLOW>(R, P2, P1, P0, t, u, v)) return true;
        if (quad && ray_x_tri<SLOW>(R, P0, P3, P2, t, u, v)) return true;
    }
    return false;
}

}  // namespace hare
