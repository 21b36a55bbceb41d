// hare_oracle.cpp -- CPU restatement of PachydermAcoustic/Hare's closest-hit
// Spatial_Partition.Shoot path.  TEST INFRASTRUCTURE ONLY.
//
//   * This file is the parity oracle and the CPU baseline.  Only tests/,
//     __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
//     legs may load it.  Nothing under hare_b200/ links, imports or calls it.
//   * PARITY UNPINNED: the reference ships no tests, golden vectors or
//     fixtures and no .NET toolchain exists in this image,
//     so the reference itself cannot be executed here.  The pins are
//     (i) analytic known-answer tests, (ii) bit-for-bit agreement with the
//     independent Python transliteration oracle/hare_oracle_py.py, and
//     (iii) oracle/csharp/HareOracle.cs, a harness to run against the real
//     reference wherever a .NET SDK exists.
//
// Every function cites the reference file:line it restates (paths relative to
// the reference checkout).  Arithmetic is IEEE binary64, no contraction
// (compile with -ffp-contract=off), operation order exactly as the C# source.
//
// Build: see oracle/Makefile  (g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math)

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <numeric>
#include <thread>
#include <utility>
#include <vector>

namespace ho {

// ---------------------------------------------------------------------------
// BCL semantics that matter
// ---------------------------------------------------------------------------

// .NET Math.Max/Math.Min propagate NaN (unlike fmax/fmin).
static inline double NetMax(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return std::signbit(a) ? b : a;  // Max(-0,+0) = +0
    return a > b ? a : b;
}
static inline double NetMin(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return std::signbit(a) ? a : b;  // Min(-0,+0) = -0
    return a < b ? a : b;
}

// (int)Math.Floor(x): x64 cvttsd2si -> 0x80000000 for NaN / out of range.
static inline int32_t FloorToInt(double x) {
    double f = std::floor(x);
    if (!(f >= -2147483648.0 && f < 2147483648.0)) return INT32_MIN;
    return (int32_t)f;
}

// Math.Round(x, 15) (MidpointRounding.ToEven):
//   if (Abs(x) < 1e16) { x *= 1e15; x = Round(x); x /= 1e15; }
static inline double Round15(double x) {
    if (std::fabs(x) < 1e16) {
        double p = 1e15;
        x = x * p;
        x = std::nearbyint(x);  // default FE_TONEAREST = half-to-even
        x = x / p;
    }
    return x;
}

// Hare_Geometry_Math.cs:43-46
static inline double Dot(double ax, double ay, double az, double bx, double by, double bz) {
    return (ax * bx) + (ay * by) + (az * bz);
}

struct P3 { double x, y, z; };

// ---------------------------------------------------------------------------
// AABB  (AABB_Main.cs:39-67 ctor, :75-84 IsPointInBox, :173-260 Intersect)
// ---------------------------------------------------------------------------
struct AABB {
    P3 Min, Max, Center, halfwidth;
    AABB() {}
    AABB(P3 mn, P3 mx) { set(mn, mx); }
    void set(P3 mn, P3 mx) {
        Min = mn; Max = mx;
        // Center = (Max + Min) / 2;  Width = Max - Min;  halfwidth = Width / 2
        Center = { (Max.x + Min.x) / 2, (Max.y + Min.y) / 2, (Max.z + Min.z) / 2 };
        P3 W = { Max.x - Min.x, Max.y - Min.y, Max.z - Min.z };
        halfwidth = { W.x / 2, W.y / 2, W.z / 2 };
    }
    bool IsPointInBox(double x, double y, double z) const {
        if (x < Min.x) return false;
        if (y < Min.y) return false;
        if (z < Min.z) return false;
        if (x > Max.x) return false;
        if (y > Max.y) return false;
        if (z > Max.z) return false;
        return true;
    }
    // AABB_Main.cs:173-260.  Moves the ray origin to the entry point.
    // Math.Abs(d) < double.Epsilon  <=>  d == 0 (Epsilon is the smallest denormal).
    bool Intersect(double& rx, double& ry, double& rz, double dx, double dy, double dz, double& tmin) const {
        tmin = 0;
        double tmax = DBL_MAX;
        const double denorm_min = 4.9406564584124654e-324;
        const double o[3] = { rx, ry, rz }, d[3] = { dx, dy, dz };
        const double mn[3] = { Min.x, Min.y, Min.z }, mx[3] = { Max.x, Max.y, Max.z };
        for (int a = 0; a < 3; ++a) {
            if (std::fabs(d[a]) < denorm_min) {
                if (o[a] < mn[a] || o[a] > mx[a]) return false;
            } else {
                double ood = (1 / d[a]);
                double t1 = (mn[a] - o[a]) * ood;
                double t2 = (mx[a] - o[a]) * ood;
                if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
                tmin = NetMax(tmin, t1);
                tmax = NetMin(tmax, t2);
                if (tmin > tmax) return false;
            }
        }
        rx = rx + dx * tmin;
        ry = ry + dy * tmin;
        rz = rz + dz * tmin;
        return true;
    }

    // ---- AABB_Tri_Int.cs:41-260 (Akenine-Moller SAT as transcribed by Hare) ----
    bool planeBoxOverlap(P3 n, P3 vert, P3 maxbox) const {  // :51-95
        P3 vmin, vmax; double v;
        v = vert.x;
        if (n.x > 0.0) { vmin.x = -maxbox.x - v; vmax.x = maxbox.x - v; }
        else           { vmin.x = maxbox.x - v;  vmax.x = -maxbox.x - v; }
        v = vert.y;
        if (n.y > 0.0) { vmin.y = -maxbox.y - v; vmax.y = maxbox.y - v; }
        else           { vmin.y = maxbox.y - v;  vmax.y = -maxbox.y - v; }
        v = vert.z;
        if (n.z > 0.0) { vmin.z = -maxbox.z - v; vmax.z = maxbox.z - v; }
        else           { vmin.z = maxbox.z - v;  vmax.z = -maxbox.z - v; }
        if (Dot(n.x, n.y, n.z, vmin.x, vmin.y, vmin.z) > 0.0) return false;
        if (Dot(n.x, n.y, n.z, vmax.x, vmax.y, vmax.z) >= 0.0) return true;
        return false;
    }

    // PolyBoxOverlap(Point[] P)  AABB_Tri_Int.cs:165-260.  P = n vertices (3 or 4).
    bool PolyBoxOverlap(const double* P, int n) const {
        const P3 hw = halfwidth;
        for (int j = 1, k = 2; k < n; ++j, ++k) {  // fan (P0,Pj,Pk)  :167-172
            const double* T0 = P; const double* T1 = P + 3 * j; const double* T2 = P + 3 * k;
            P3 v0 = { T0[0] - Center.x, T0[1] - Center.y, T0[2] - Center.z };
            P3 v1 = { T1[0] - Center.x, T1[1] - Center.y, T1[2] - Center.z };
            P3 v2 = { T2[0] - Center.x, T2[1] - Center.y, T2[2] - Center.z };
            P3 e0 = { v1.x - v0.x, v1.y - v0.y, v1.z - v0.z };
            P3 e1 = { v2.x - v1.x, v2.y - v1.y, v2.z - v1.z };
            P3 e2 = { v0.x - v2.x, v0.y - v2.y, v0.z - v2.z };
            double p0, p1, p2, mn, mx, rad, a, b, fa, fb, fex, fey, fez;
#define HO_MINMAX(A, B) if ((A) < (B)) { mn = (A); mx = (B); } else { mn = (B); mx = (A); }
#define HO_REJECT if (mn > rad || mx < -rad) continue;
            // ---- edge 0 ----
            fex = std::fabs(e0.x); fey = std::fabs(e0.y); fez = std::fabs(e0.z);
            // AXISTEST_X01(e0.z, e0.y, fez, fey)
            a = e0.z; b = e0.y; fa = fez; fb = fey;
            p0 = a * v0.y - b * v0.z; p2 = a * v2.y - b * v2.z; HO_MINMAX(p0, p2)
            rad = fa * hw.y + fb * hw.z; HO_REJECT
            // AXISTEST_Y02(e0.z, e0.x, fez, fex)
            a = e0.z; b = e0.x; fa = fez; fb = fex;
            p0 = -a * v0.x + b * v0.z; p2 = -a * v2.x + b * v2.z; HO_MINMAX(p0, p2)
            rad = fa * hw.x + fb * hw.z; HO_REJECT
            // AXISTEST_Z12(e0.y, e0.x, fey, fex)   note: (p2 < p1) ordering
            a = e0.y; b = e0.x; fa = fey; fb = fex;
            p1 = a * v1.x - b * v1.y; p2 = a * v2.x - b * v2.y;
            if (p2 < p1) { mn = p2; mx = p1; } else { mn = p1; mx = p2; }
            rad = fa * hw.x + fb * hw.y; HO_REJECT
            // ---- edge 1 ----
            fex = std::fabs(e1.x); fey = std::fabs(e1.y); fez = std::fabs(e1.z);
            a = e1.z; b = e1.y; fa = fez; fb = fey;                         // X01
            p0 = a * v0.y - b * v0.z; p2 = a * v2.y - b * v2.z; HO_MINMAX(p0, p2)
            rad = fa * hw.y + fb * hw.z; HO_REJECT
            a = e1.z; b = e1.x; fa = fez; fb = fex;                         // Y02
            p0 = -a * v0.x + b * v0.z; p2 = -a * v2.x + b * v2.z; HO_MINMAX(p0, p2)
            rad = fa * hw.x + fb * hw.z; HO_REJECT
            a = e1.y; b = e1.x; fa = fey; fb = fex;                         // Z0
            p0 = a * v0.x - b * v0.y; p1 = a * v1.x - b * v1.y; HO_MINMAX(p0, p1)
            rad = fa * hw.x + fb * hw.y; HO_REJECT
            // ---- edge 2 ----
            fex = std::fabs(e2.x); fey = std::fabs(e2.y); fez = std::fabs(e2.z);
            a = e2.z; b = e2.y; fa = fez; fb = fey;                         // X2
            p0 = a * v0.y - b * v0.z; p1 = a * v1.y - b * v1.z; HO_MINMAX(p0, p1)
            rad = fa * hw.y + fb * hw.z; HO_REJECT
            a = e2.z; b = e2.x; fa = fez; fb = fex;                         // Y1
            p0 = -a * v0.x + b * v0.z; p1 = -a * v1.x + b * v1.z; HO_MINMAX(p0, p1)
            rad = fa * hw.x + fb * hw.z; HO_REJECT
            a = e2.y; b = e2.x; fa = fey; fb = fex;                         // Z12
            p1 = a * v1.x - b * v1.y; p2 = a * v2.x - b * v2.y;
            if (p2 < p1) { mn = p2; mx = p1; } else { mn = p1; mx = p2; }
            rad = fa * hw.x + fb * hw.y; HO_REJECT
#undef HO_MINMAX
#undef HO_REJECT
            // ---- box axes (FINDMINMAX :41-49, :240-249) ----
            mn = v0.x; mx = v0.x; if (v1.x < mn) mn = v1.x; if (v1.x > mx) mx = v1.x; if (v2.x < mn) mn = v2.x; if (v2.x > mx) mx = v2.x;
            if (mn > hw.x || mx < -hw.x) continue;
            mn = v0.y; mx = v0.y; if (v1.y < mn) mn = v1.y; if (v1.y > mx) mx = v1.y; if (v2.y < mn) mn = v2.y; if (v2.y > mx) mx = v2.y;
            if (mn > hw.y || mx < -hw.y) continue;
            mn = v0.z; mx = v0.z; if (v1.z < mn) mn = v1.z; if (v1.z > mx) mx = v1.z; if (v2.z < mn) mn = v2.z; if (v2.z > mx) mx = v2.z;
            if (mn > hw.z || mx < -hw.z) continue;
            // ---- plane (Cross(Vector,Vector) Hare_Geometry_Math.cs:70-73) ----
            P3 nrm = { e0.y * e1.z - e0.z * e1.y, -(e0.x * e1.z - e0.z * e1.x), e0.x * e1.y - e0.y * e1.x };
            if (!planeBoxOverlap(nrm, v0, hw)) continue;
            return true;
        }
        return false;
    }
};

// ---------------------------------------------------------------------------
// Topology  (Hare_Geometry_Topology.cs:85-91 ctor(min,max), :225-254 Add_Polygon,
//            :342-377 AddGetIndex, :148-179 Finish_Topology, :677-697 MS_AABB;
//            Hare_Geometry_Primitives.cs:230-250 Round/Hash2;
//            Hare_Geometry_Polygons.cs:148-194 Polygon ctor -> Normal)
// ---------------------------------------------------------------------------
struct Topology {
    P3 Min, Max;               // padded bounds
    P3 MsMin;                  // Modspace.Min
    int64_t msdim = 0, msXYTot = 0;
    std::map<std::pair<uint64_t, uint64_t>, int> weld;   // (bucket,pos) -> vertex index
    std::vector<P3> Vertices_List;
    std::vector<double> verts;     // P x 4 x 3 (triangles repeat v2 into slot 3, unused)
    std::vector<double> normals;   // P x 3
    std::vector<int32_t> vcount;   // P
    int64_t P = 0;

    void set_modspace(P3 mn, P3 mx) {   // MS_AABB ctor
        MsMin = mn;
        double xl = mx.x - mn.x, yl = mx.y - mn.y, zl = mx.z - mn.z;
        int m = std::max((int)std::ceil(xl), std::max((int)std::ceil(yl), (int)std::ceil(zl)));
        msdim = m; msXYTot = (int64_t)m * m;
    }
    Topology(P3 minpt, P3 maxpt) {
        Max = { maxpt.x + 0.000000000001, maxpt.y + 0.000000000001, maxpt.z + 0.000000000001 };
        Min = { minpt.x - 0.000000000001, minpt.y - 0.000000000001, minpt.z - 0.000000000001 };
        set_modspace(Min, Max);
    }
    // AddGetIndex.  Returns the (possibly welded) vertex.
    P3 AddGetIndex(P3 p) {
        p.x = Round15(p.x); p.y = Round15(p.y); p.z = Round15(p.z);
        double Xoff = p.x - MsMin.x, Yoff = p.y - MsMin.y, Zoff = p.z - MsMin.z;
        uint64_t xl = (uint64_t)std::floor(Xoff), yl = (uint64_t)std::floor(Yoff), zl = (uint64_t)std::floor(Zoff);
        uint64_t bucket = (uint64_t)msXYTot * zl + (uint64_t)msdim * xl + yl;
        uint64_t xp = (uint64_t)((Xoff - (double)xl) * 1000), yp = (uint64_t)((Yoff - (double)yl) * 1000), zp = (uint64_t)((Zoff - (double)zl) * 1000);
        uint64_t pos = 1000000 * zp + 1000 * xp + yp;
        auto key = std::make_pair(bucket, pos);
        auto it = weld.find(key);
        if (it != weld.end()) return Vertices_List[it->second];
        weld.emplace(key, (int)Vertices_List.size());
        Vertices_List.push_back(p);
        return p;
    }
    // Add_Polygon + Polygon ctor (Normal only; the rest is not read by Shoot).
    int Add_Polygon(const double* pts, int n) {
        if (n != 3 && n != 4) return -1;   // NotImplementedException
        P3 V[4];
        for (int i = 0; i < n; ++i) V[i] = AddGetIndex({ pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] });
        if (n == 3) V[3] = V[2];
        P3 N = { 0, 0, 0 };
        for (int j = 2; j < n; ++j) {
            P3 a = { V[1].x - V[0].x, V[1].y - V[0].y, V[1].z - V[0].z };
            P3 b = { V[j].x - V[0].x, V[j].y - V[0].y, V[j].z - V[0].z };
            N = { a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x };
            if (!((N.x * N.x + N.y * N.y + N.z * N.z) < 4.9406564584124654e-324)) break;  // !IsZeroVector
        }
        double f = N.x * N.x + N.y * N.y + N.z * N.z;   // Vector.Normalize
        if (f != 0) { f = std::sqrt(f); N.x /= f; N.y /= f; N.z /= f; }
        for (int i = 0; i < 4; ++i) { verts.push_back(V[i].x); verts.push_back(V[i].y); verts.push_back(V[i].z); }
        normals.push_back(N.x); normals.push_back(N.y); normals.push_back(N.z);
        vcount.push_back(n);
        ++P;
        return 0;
    }
    void Finish_Topology() {
        double Minx = DBL_MAX, Miny = DBL_MAX, Minz = DBL_MAX, Maxx = -DBL_MAX, Maxy = -DBL_MAX, Maxz = -DBL_MAX;
        for (const P3& p : Vertices_List) {
            if (Minx > p.x) Minx = p.x; if (Miny > p.y) Miny = p.y; if (Minz > p.z) Minz = p.z;
            if (Maxx < p.x) Maxx = p.x; if (Maxy < p.y) Maxy = p.y; if (Maxz < p.z) Maxz = p.z;
        }
        set_modspace({ Minx, Miny, Minz }, { Maxx, Maxy, Maxz });
        Min = { Minx - 0.000000000001, Miny - 0.000000000001, Minz - 0.000000000001 };
        Max = { Maxx + 0.000000000001, Maxy + 0.000000000001, Maxz + 0.000000000001 };
    }
    const double* PV(int64_t i) const { return &verts[12 * i]; }
    const double* NV(int64_t i) const { return &normals[3 * i]; }
};

// ---------------------------------------------------------------------------
// Polygon intersection
// ---------------------------------------------------------------------------
struct Ray { double x, y, z, dx, dy, dz; int32_t Ray_ID; };

// Fast path RayXtri(ref ...)  Hare_Geometry_Polygons.cs:449-510
static inline bool RayXtri_fast(const Ray& R, const double* a, const double* b, const double* c, double& t) {
    double edge1x = b[0] - a[0], edge1y = b[1] - a[1], edge1z = b[2] - a[2];
    double edge2x = c[0] - a[0], edge2y = c[1] - a[1], edge2z = c[2] - a[2];
    double u, v;
    double pvecx = R.dy * edge2z - R.dz * edge2y;
    double pvecy = R.dz * edge2x - R.dx * edge2z;
    double pvecz = R.dx * edge2y - R.dy * edge2x;
    double det = Dot(edge1x, edge1y, edge1z, pvecx, pvecy, pvecz);
    double tvecx = R.x - a[0], tvecy = R.y - a[1], tvecz = R.z - a[2];
    double invdet = 1.0 / det;
    double qvecx = tvecy * edge1z - tvecz * edge1y;
    double qvecy = tvecz * edge1x - tvecx * edge1z;
    double qvecz = tvecx * edge1y - tvecy * edge1x;
    if (det > 0.000001) {
        u = Dot(tvecx, tvecy, tvecz, pvecx, pvecy, pvecz);
        if (u < 0.0 || u > det) return false;
        v = Dot(R.dx, R.dy, R.dz, qvecx, qvecy, qvecz);
        if (v < 0.0 || u + v > det) return false;
    } else if (det < -0.000001) {
        u = Dot(tvecx, tvecy, tvecz, pvecx, pvecy, pvecz);
        if (u > 0.0 || u < det) return false;
        v = Dot(R.dx, R.dy, R.dz, qvecx, qvecy, qvecz);
        if (v > 0.0 || u + v < det) return false;
    } else return false;
    t = Dot(edge2x, edge2y, edge2z, qvecx, qvecy, qvecz) * invdet;
    return true;
}

// Slow path RayXtri(Ray, ...)  Hare_Geometry_Polygons.cs:385-435 ; Cross :74-77
static inline bool RayXtri_slow(const Ray& R, const double* a, const double* b, const double* c, double& t, double& u, double& v) {
    double e1x = b[0] - a[0], e1y = b[1] - a[1], e1z = b[2] - a[2];
    double e2x = c[0] - a[0], e2y = c[1] - a[1], e2z = c[2] - a[2];
    double px = R.dy * e2z - R.dz * e2y, py = -(R.dx * e2z - R.dz * e2x), pz = R.dx * e2y - R.dy * e2x;
    double det = (e1x * px) + (e1y * py) + (e1z * pz);
    double tvx = R.x - a[0], tvy = R.y - a[1], tvz = R.z - a[2];
    double invdet = 1.0 / det;
    double qx = tvy * e1z - tvz * e1y, qy = -(tvx * e1z - tvz * e1x), qz = tvx * e1y - tvy * e1x;
    if (det > 0.000001) {
        u = Dot(tvx, tvy, tvz, px, py, pz);
        if (u < 0.0 || u > det) return false;
        v = Dot(R.dx, R.dy, R.dz, qx, qy, qz);
        if (v < 0.0 || u + v > det) return false;
    } else if (det < -0.000001) {
        u = Dot(tvx, tvy, tvz, px, py, pz);
        if (u > 0.0 || u < det) return false;
        v = Dot(R.dx, R.dy, R.dz, qx, qy, qz);
        if (v > 0.0 || u + v < det) return false;
    } else return false;
    t = ((e2x * qx) + (e2y * qy) + (e2z * qz)) * invdet;
    u *= invdet;
    v *= invdet;
    return true;
}

// Polygon.Ray_Side  Hare_Geometry_Polygons.cs:601-606
static inline bool Ray_Side(const Ray& R, const double* N) {
    double n = Dot(R.dx, R.dy, R.dz, N[0], N[1], N[2]);
    if (n < 0) return false;
    return true;
}

// Triangle.Intersect(ref...) :637-660  /  Quadrilateral.Intersect(ref...) :784-823
static inline bool intersect_fast(const Topology& T, int64_t i, const Ray& R, double& x, double& y, double& z, double& t) {
    const double* P = T.PV(i); const double* P0 = P, *P1 = P + 3, *P2 = P + 6, *P3_ = P + 9;
    t = 0;
    bool hit;
    if (T.vcount[i] == 3) {
        if (Ray_Side(R, T.NV(i))) hit = RayXtri_fast(R, P0, P1, P2, t);
        else hit = RayXtri_fast(R, P2, P1, P0, t);
    } else {
        if (Ray_Side(R, T.NV(i))) {
            hit = RayXtri_fast(R, P0, P1, P2, t);
            if (!hit) hit = RayXtri_fast(R, P2, P3_, P0, t);
        } else {
            hit = RayXtri_fast(R, P2, P1, P0, t);
            if (!hit) hit = RayXtri_fast(R, P0, P3_, P2, t);
        }
    }
    if (hit) { x = R.x + R.dx * t; y = R.y + R.dy * t; z = R.z + R.dz * t; return true; }
    x = 0; y = 0; z = 0;
    return false;
}

// Triangle.Intersect(Ray...) :662-688  /  Quadrilateral.Intersect(Ray...) :731-782
static inline bool intersect_slow(const Topology& T, int64_t i, const Ray& R, P3& X, double& u, double& v, double& t) {
    const double* P = T.PV(i); const double* P0 = P, *P1 = P + 3, *P2 = P + 6, *P3_ = P + 9;
    u = 0; v = 0; t = 0;
    bool hit;
    if (T.vcount[i] == 3) {
        if (Ray_Side(R, T.NV(i))) hit = RayXtri_slow(R, P0, P1, P2, t, u, v);
        else hit = RayXtri_slow(R, P2, P1, P0, t, u, v);
    } else {
        if (Ray_Side(R, T.NV(i))) {
            hit = RayXtri_slow(R, P0, P1, P2, t, u, v);
            if (!hit) hit = RayXtri_slow(R, P2, P3_, P0, t, u, v);
        } else {
            hit = RayXtri_slow(R, P2, P1, P0, t, u, v);
            if (!hit) hit = RayXtri_slow(R, P0, P3_, P2, t, u, v);
        }
    }
    if (hit) { X = { R.x + R.dx * t, R.y + R.dy * t, R.z + R.dz * t }; return true; }
    return false;
}

struct XEvent {   // Hare_Geometry_Primitives.cs:435-481 ; miss = X_Event()
    bool Hit = false; int32_t Poly_id = -1; double t = 0, u = 0, v = 0; P3 X = { 0, 0, 0 };
};
struct Counters { uint64_t cells = 0, entries = 0, tests = 0, hits = 0; };

// ---------------------------------------------------------------------------
// Partitions
// ---------------------------------------------------------------------------
struct Partition {
    const Topology* T = nullptr;
    virtual ~Partition() {}
    // returns 1 hit, 0 miss, -2 reference would throw (index out of range)
    virtual int Shoot(Ray& R, XEvent& ev, int o1, int o2, int32_t* mailbox, Counters& c) const = 0;
};

// ---- Voxel_Grid  (Voxel_Grid.cs) -------------------------------------------
struct VoxelGrid : Partition {
    int Ct[3] = { 0, 0, 0 };
    AABB OBox;
    P3 BoxDims, VoxelDims;
    double Epsilon = 0.001;
    std::vector<uint32_t> cell_offset;   // row-major ((x*Ny+y)*Nz+z), +1
    std::vector<uint32_t> cell_poly;
    std::vector<std::vector<int>> lists;   // build-time

    void bounds() {   // Voxel_Grid.cs:50-75 (single topology)
        const Topology& M = *T;
        P3 MaxPT = { -INFINITY, -INFINITY, -INFINITY }, MinPT = { INFINITY, INFINITY, INFINITY };
        if ((M.Max.x + 0.01) > MaxPT.x) MaxPT.x = (M.Max.x + Epsilon);
        if ((M.Max.y + 0.01) > MaxPT.y) MaxPT.y = (M.Max.y + Epsilon);
        if ((M.Max.z + 0.01) > MaxPT.z) MaxPT.z = (M.Max.z + Epsilon);
        if ((M.Min.x - 0.01) < MinPT.x) MinPT.x = (M.Min.x - Epsilon);
        if ((M.Min.y - 0.01) < MinPT.y) MinPT.y = (M.Min.y - Epsilon);
        if ((M.Min.z - 0.01) < MinPT.z) MinPT.z = (M.Min.z - Epsilon);
        OBox.set({ MinPT.x - .1, MinPT.y - .1, MinPT.z - .1 }, { MaxPT.x + .1, MaxPT.y + .1, MaxPT.z + .1 });
        BoxDims = { OBox.Max.x - OBox.Min.x, OBox.Max.y - OBox.Min.y, OBox.Max.z - OBox.Min.z };
    }
    void set_domain(int nx, int ny, int nz) {
        Ct[0] = nx; Ct[1] = ny; Ct[2] = nz;
        VoxelDims = { BoxDims.x / nx, BoxDims.y / ny, BoxDims.z / nz };
    }
    // voxel box, Voxel_Grid.cs:283-285 : new AABB(VoxelMin + OBox.Min, VoxelMax + OBox.Min)
    AABB voxel(int x, int y, int z) const {
        P3 mn = { x * VoxelDims.x - Epsilon, y * VoxelDims.y - Epsilon, z * VoxelDims.z - Epsilon };
        P3 mx = { (x + 1) * VoxelDims.x + Epsilon, (y + 1) * VoxelDims.y + Epsilon, (z + 1) * VoxelDims.z + Epsilon };
        return AABB({ mn.x + OBox.Min.x, mn.y + OBox.Min.y, mn.z + OBox.Min.z }, { mx.x + OBox.Min.x, mx.y + OBox.Min.y, mx.z + OBox.Min.z });
    }
    size_t cid(int x, int y, int z) const { return ((size_t)x * Ct[1] + y) * Ct[2] + z; }
    void to_csr() {
        size_t nc = lists.size();
        cell_offset.assign(nc + 1, 0);
        for (size_t c = 0; c < nc; ++c) cell_offset[c + 1] = cell_offset[c] + (uint32_t)lists[c].size();
        cell_poly.resize(cell_offset[nc]);
        for (size_t c = 0; c < nc; ++c) std::copy(lists[c].begin(), lists[c].end(), cell_poly.begin() + cell_offset[c]);
        lists.clear(); lists.shrink_to_fit();
    }
    // Voxel_Grid(Model, Domain) + Fill_Voxels, Voxel_Grid.cs:48-121, :273-304: O(D^3 P) literal.
    void build_flat(int Domain, int nthreads) {
        bounds(); set_domain(Domain, Domain, Domain);
        lists.assign((size_t)Domain * Domain * Domain, {});
        auto work = [&](int x0, int x1) {
            for (int x = x0; x < x1; ++x) for (int y = 0; y < Ct[1]; ++y) for (int z = 0; z < Ct[2]; ++z) {
                AABB Box = voxel(x, y, z);
                auto& L = lists[cid(x, y, z)];
                for (int64_t i = 0; i < T->P; ++i)
                    if (Box.PolyBoxOverlap(T->PV(i), T->vcount[i])) L.push_back((int)i);
            }
        };
        std::vector<std::thread> th;
        for (int p = 0; p < nthreads; ++p) th.emplace_back(work, p * Domain / nthreads, (p + 1) * Domain / nthreads);
        for (auto& t : th) t.join();
        to_csr();
    }
    // Same result as build_flat, evaluated polygon-major over a conservative voxel
    // range (the SAT's three box-axis tests reject every voxel outside it).  NOT a
    // reference code path: an accelerated evaluation of the same predicate, checked
    // against build_flat and the committed golden lists in tests/test_oracle_pins.py (test_golden_voxelgrid_lists, test_hier_ctor_matches_flat_on_small_case).
    void build_flat_fast(int Domain) {
        bounds(); set_domain(Domain, Domain, Domain);
        lists.assign((size_t)Domain * Domain * Domain, {});
        const double vd[3] = { VoxelDims.x, VoxelDims.y, VoxelDims.z }, om[3] = { OBox.Min.x, OBox.Min.y, OBox.Min.z };
        for (int64_t i = 0; i < T->P; ++i) {
            const double* P = T->PV(i); int n = T->vcount[i];
            int lo[3], hi[3];
            for (int a = 0; a < 3; ++a) {
                double mn = P[a], mx = P[a];
                for (int k = 1; k < n; ++k) { mn = std::min(mn, P[3 * k + a]); mx = std::max(mx, P[3 * k + a]); }
                lo[a] = std::max(0, (int)std::floor((mn - om[a]) / vd[a]) - 2);
                hi[a] = std::min(Ct[a] - 1, (int)std::floor((mx - om[a]) / vd[a]) + 2);
            }
            for (int x = lo[0]; x <= hi[0]; ++x) for (int y = lo[1]; y <= hi[1]; ++y) for (int z = lo[2]; z <= hi[2]; ++z)
                if (voxel(x, y, z).PolyBoxOverlap(P, n)) lists[cid(x, y, z)].push_back((int)i);
        }
        to_csr();
    }
    // Voxel_Grid(Model, MaxDomain, Avg_polys)  Voxel_Grid.cs:128-254 (hierarchical 2x refinement)
    void build_hier(int MaxDomain, int Avg_polys, int nthreads) {
        bounds();
        int n = 1;
        std::vector<std::vector<int>> cur(1);
        cur[0].resize(T->P); std::iota(cur[0].begin(), cur[0].end(), 0);
        Ct[0] = Ct[1] = Ct[2] = 1;
        for (int k = 0; k < MaxDomain; ++k) {
            int pn = n; n = 2 * n;
            set_domain(n, n, n);
            std::vector<std::vector<int>> nxt((size_t)n * n * n);
            auto work = [&](int x0, int x1) {
                for (int x = x0; x < x1; ++x) for (int y = 0; y < n; ++y) for (int z = 0; z < n; ++z) {
                    AABB Box = voxel(x, y, z);
                    int xp = (int)std::floor((double)x / 2), yp = (int)std::floor((double)y / 2), zp = (int)std::floor((double)z / 2);
                    auto& L = nxt[cid(x, y, z)];
                    for (int i : cur[((size_t)xp * pn + yp) * pn + zp])
                        if (Box.PolyBoxOverlap(T->PV(i), T->vcount[i])) L.push_back(i);
                }
            };
            std::vector<std::thread> th;
            for (int p = 0; p < nthreads; ++p) th.emplace_back(work, p * n / nthreads, (p + 1) * n / nthreads);
            for (auto& t : th) t.join();
            cur.swap(nxt);
            double sum = 0; int ct = 0;
            for (auto& L : cur) if (!L.empty()) { sum += (double)L.size(); ct++; }
            if (k > 1 && sum / ct < Avg_polys) break;
        }
        lists.swap(cur);
        to_csr();
    }

    // Voxel_Grid.Shoot  Voxel_Grid.cs:351-552 (o1=o2=-1 reproduces :561-761 provided no
    // polygon has a negative index).
    int Shoot(Ray& R, XEvent& ev, int o1, int o2, int32_t* mailbox, Counters& c) const override {
        ev = XEvent();
        int X = FloorToInt((R.x - OBox.Min.x) / VoxelDims.x);
        int Y = FloorToInt((R.y - OBox.Min.y) / VoxelDims.y);
        int Z = FloorToInt((R.z - OBox.Min.z) / VoxelDims.z);
        double tDeltaX, tDeltaY, tDeltaZ, tMaxX = 0, tMaxY = 0, tMaxZ = 0;
        int stepX, stepY, stepZ;
        double t_start = 0;
        if (X < 0 || X >= Ct[0] || Y < 0 || Y >= Ct[1] || Z < 0 || Z >= Ct[2]) {
            if (!OBox.Intersect(R.x, R.y, R.z, R.dx, R.dy, R.dz, t_start)) return 0;
            X = FloorToInt((R.x - OBox.Min.x + R.dx * 1E-6) / VoxelDims.x);
            Y = FloorToInt((R.y - OBox.Min.y + R.dy * 1E-6) / VoxelDims.y);
            Z = FloorToInt((R.z - OBox.Min.z + R.dz * 1E-6) / VoxelDims.z);
            if (X < 0 || X >= Ct[0] || Y < 0 || Y >= Ct[1] || Z < 0 || Z >= Ct[2]) return -2;  // C#: IndexOutOfRangeException at Voxels[X,Y,Z]
        }
        AABB V0 = voxel(X, Y, Z);
        if (R.dx < 0) { stepX = -1; tMaxX = (V0.Min.x - R.x) / R.dx; tDeltaX = VoxelDims.x / R.dx * stepX; }
        else          { stepX = 1;  tMaxX = (V0.Max.x - R.x) / R.dx; tDeltaX = VoxelDims.x / R.dx * stepX; }
        if (R.dy < 0) { stepY = -1; tMaxY = (V0.Min.y - R.y) / R.dy; tDeltaY = VoxelDims.y / R.dy * stepY; }
        else          { stepY = 1;  tMaxY = (V0.Max.y - R.y) / R.dy; tDeltaY = VoxelDims.y / R.dy * stepY; }
        if (R.dz < 0) { stepZ = -1; tMaxZ = (V0.Min.z - R.z) / R.dz; tDeltaZ = VoxelDims.z / R.dz * stepZ; }
        else          { stepZ = 1;  tMaxZ = (V0.Max.z - R.z) / R.dz; tDeltaZ = VoxelDims.z / R.dz * stepZ; }

        bool have = false; P3 Xpt = { 0, 0, 0 };
        double tmin = DBL_MAX; int pid = -1;
        while (true) {
            ++c.cells;
            size_t ci = cid(X, Y, Z);
            for (uint32_t k = cell_offset[ci]; k < cell_offset[ci + 1]; ++k) {
                int i = (int)cell_poly[k];
                ++c.entries;
                if (i == o1 || i == o2) continue;
                if (mailbox[i] != R.Ray_ID) {
                    mailbox[i] = R.Ray_ID;
                    ++c.tests;
                    double x, y, z, t;
                    if (intersect_fast(*T, i, R, x, y, z, t) && t > 0.0000000001) {
                        if (t < tmin) { have = true; Xpt = { x, y, z }; tmin = t; pid = i; }
                    }
                }
            }
            if (have && voxel(X, Y, Z).IsPointInBox(Xpt.x, Xpt.y, Xpt.z)) {
                ev.Hit = true; ev.X = Xpt; ev.u = 0; ev.v = 0; ev.t = tmin + t_start; ev.Poly_id = pid;
                ++c.hits;
                return 1;
            }
            if (tMaxX < tMaxY) {
                if (tMaxX < tMaxZ) { X += stepX; if (X < 0 || X >= Ct[0]) return 0; tMaxX = tMaxX + tDeltaX; }
                else               { Z += stepZ; if (Z < 0 || Z >= Ct[2]) return 0; tMaxZ = tMaxZ + tDeltaZ; }
            } else {
                if (tMaxY < tMaxZ) { Y += stepY; if (Y < 0 || Y >= Ct[1]) return 0; tMaxY = tMaxY + tDeltaY; }
                else               { Z += stepZ; if (Z < 0 || Z >= Ct[2]) return 0; tMaxZ = tMaxZ + tDeltaZ; }
            }
        }
    }
};

// ---- Octree  ("Octree - alt.cs") -------------------------------------------
struct Octree : Partition {
    struct Node { P3 Min, Max; int32_t first_child = -1; uint32_t list_off = 0, list_cnt = 0; };
    std::vector<Node> nodes;           // children of a node are 8 consecutive entries
    std::vector<uint32_t> polys;       // leaf lists, concatenated
    int maxDepth = 0, maxPolys = 0;
    uint64_t lost = 0;

    // ctor :45-89
    void build(int maxDepth_, int maxPolys_) {
        maxDepth = maxDepth_; maxPolys = maxPolys_;
        P3 mn = { INFINITY, INFINITY, INFINITY }, mx = { -INFINITY, -INFINITY, -INFINITY };
        for (const P3& v : T->Vertices_List) {
            if (v.x < mn.x) mn.x = v.x; if (v.y < mn.y) mn.y = v.y; if (v.z < mn.z) mn.z = v.z;
            if (v.x > mx.x) mx.x = v.x; if (v.y > mx.y) mx.y = v.y; if (v.z > mx.z) mx.z = v.z;
        }
        double maxdim = NetMax(mx.x - mn.x, NetMax(mx.y - mn.y, mx.z - mn.z));
        P3 center = { mx.x + mn.x / 2, mx.y + mn.y / 2, mx.z + mn.z / 2 };   // "max + min / 2" [sic] :79
        Node root;
        root.Min = { center.x - maxdim - 1e-1, center.y - maxdim - 1e-1, center.z - maxdim - 1e-1 };
        root.Max = { center.x + maxdim + 1e-1, center.y + maxdim + 1e-1, center.z + maxdim + 1e-1 };
        nodes.clear(); polys.clear();
        nodes.push_back(root);
        std::vector<int> all(T->P); std::iota(all.begin(), all.end(), 0);
        split(0, 0, all);
    }
    // BuildOctree :91-138
    void split(int ni, int depth, std::vector<int>& list) {
        if (depth >= maxDepth || (int)list.size() <= maxPolys) {
            nodes[ni].list_off = (uint32_t)polys.size(); nodes[ni].list_cnt = (uint32_t)list.size();
            polys.insert(polys.end(), list.begin(), list.end());
            return;
        }
        AABB nb(nodes[ni].Min, nodes[ni].Max);
        P3 c = nb.Center;
        int fc = (int)nodes.size();
        nodes[ni].first_child = fc;
        AABB cb[8];
        for (int i = 0; i < 8; ++i) {
            P3 mn = { ((i & 4) == 0 ? nb.Min.x : c.x) - 0.1, ((i & 2) == 0 ? nb.Min.y : c.y) - 0.1, ((i & 1) == 0 ? nb.Min.z : c.z) - 0.1 };
            P3 mx = { ((i & 4) == 0 ? c.x : nb.Max.x) + 0.1, ((i & 2) == 0 ? c.y : nb.Max.y) + 0.1, ((i & 1) == 0 ? c.z : nb.Max.z) + 0.1 };
            cb[i].set(mn, mx);
            Node ch; ch.Min = mn; ch.Max = mx; nodes.push_back(ch);
        }
        std::vector<int> cl[8];
        for (int p : list) {
            bool stored = false;
            for (int i = 0; i < 8; ++i)
                if (cb[i].PolyBoxOverlap(T->PV(p), T->vcount[p])) { cl[i].push_back(p); stored = true; }
            if (!stored) ++lost;
        }
        list.clear(); list.shrink_to_fit();
        for (int i = 0; i < 8; ++i) split(fc + i, depth + 1, cl[i]);
    }

    // Shoot :159-284 ; ComputeTraversalOrder :286-306
    int Shoot(Ray& ray, XEvent& ev, int o1, int o2, int32_t*, Counters& c) const override {
        ev = XEvent();
        double invDx = std::fabs(ray.dx) > 1e-16 ? 1.0 / ray.dx : 1e16;
        double invDy = std::fabs(ray.dy) > 1e-16 ? 1.0 / ray.dy : 1e16;
        double invDz = std::fabs(ray.dz) > 1e-16 ? 1.0 / ray.dz : 1e16;
        auto interval = [&](const Node& n, double& lo, double& hi) {
            double tx0 = (n.Min.x - ray.x) * invDx, tx1 = (n.Max.x - ray.x) * invDx;
            double ty0 = (n.Min.y - ray.y) * invDy, ty1 = (n.Max.y - ray.y) * invDy;
            double tz0 = (n.Min.z - ray.z) * invDz, tz1 = (n.Max.z - ray.z) * invDz;
            if (invDx < 0) { double s = tx0; tx0 = tx1; tx1 = s; }
            if (invDy < 0) { double s = ty0; ty0 = ty1; ty1 = s; }
            if (invDz < 0) { double s = tz0; tz0 = tz1; tz1 = s; }
            lo = NetMax(NetMax(tx0, ty0), tz0);
            hi = NetMin(NetMin(tx1, ty1), tz1);
        };
        double tmin, tmax;
        interval(nodes[0], tmin, tmax);
        if (tmax < tmin || tmax < 0) return 0;
        int order[8]; int oi = 0;
        {
            int xDir = ray.dx >= 0 ? 0 : 1, yDir = ray.dy >= 0 ? 0 : 1, zDir = ray.dz >= 0 ? 0 : 1;
            for (int ix = xDir; ix <= 1 && ix >= 0; ix += (ray.dx >= 0 ? 1 : -1))
                for (int iy = yDir; iy <= 1 && iy >= 0; iy += (ray.dy >= 0 ? 1 : -1))
                    for (int iz = zDir; iz <= 1 && iz >= 0; iz += (ray.dz >= 0 ? 1 : -1))
                        order[oi++] = (ix << 2) | (iy << 1) | iz;
        }
        struct E { int n; double a, b; };
        std::vector<E> stack; stack.reserve(64);
        stack.push_back({ 0, tmin, tmax });
        bool hit = false; double closestT = DBL_MAX;
        while (!stack.empty()) {
            E e = stack.back(); stack.pop_back();
            if (e.b < e.a || e.b < 0) continue;
            if (hit && closestT <= e.a) continue;
            const Node& n = nodes[e.n];
            ++c.cells;
            if (n.first_child < 0) {
                for (uint32_t k = n.list_off; k < n.list_off + n.list_cnt; ++k) {
                    int p = (int)polys[k];
                    ++c.entries;
                    if (p == o1 || p == o2) continue;
                    ++c.tests;
                    P3 X; double u, v, t;
                    if (intersect_slow(*T, p, ray, X, u, v, t) && t > 0.0000000001) {
                        if (t < closestT) {
                            closestT = t;
                            ev.Hit = true; ev.X = X; ev.u = u; ev.v = v; ev.t = t; ev.Poly_id = p;
                            hit = true;
                            if (closestT <= e.a) { ++c.hits; return 1; }
                        }
                    }
                }
            } else {
                for (int q = 0; q < 8; ++q) {
                    int ci = n.first_child + order[q];
                    double ca, cb_;
                    interval(nodes[ci], ca, cb_);
                    if (cb_ < ca || cb_ < 0 || ca > e.b || cb_ < e.a) continue;
                    stack.push_back({ ci, NetMax(ca, e.a), NetMin(cb_, e.b) });
                }
            }
        }
        if (hit) { ++c.hits; return 1; }
        ev = XEvent();
        return 0;
    }
};

// ---- KDTree  (KDTree.cs) -----------------------------------------------------
struct KDTree : Partition {
    struct Node { P3 Min, Max; double split = 0; int32_t axis = -1; int32_t left = -1, right = -1; uint32_t list_off = 0, list_cnt = 0; };
    std::vector<Node> nodes;
    std::vector<uint32_t> polys;
    std::vector<P3> centroid;
    int maxDepth = 0, maxPolys = 0;

    static double byint(const P3& p, int a) { return a == 0 ? p.x : (a == 1 ? p.y : p.z); }

    // ctor :51-88
    void build(int maxDepth_, int maxPolys_) {
        maxDepth = maxDepth_; maxPolys = maxPolys_;
        P3 mn = { INFINITY, INFINITY, INFINITY }, mx = { -INFINITY, -INFINITY, -INFINITY };
        for (const P3& v : T->Vertices_List) {
            if (v.x < mn.x) mn.x = v.x; if (v.y < mn.y) mn.y = v.y; if (v.z < mn.z) mn.z = v.z;
            if (v.x > mx.x) mx.x = v.x; if (v.y > mx.y) mx.y = v.y; if (v.z > mx.z) mx.z = v.z;
        }
        // Polygon_Centroid  Hare_Geometry_Topology.cs:566-575: ((0 + v0) + v1 + ...) / n
        centroid.resize(T->P);
        for (int64_t i = 0; i < T->P; ++i) {
            P3 s = { 0, 0, 0 }; const double* P = T->PV(i); int n = T->vcount[i];
            for (int k = 0; k < n; ++k) s = { s.x + P[3 * k], s.y + P[3 * k + 1], s.z + P[3 * k + 2] };
            centroid[i] = { s.x / n, s.y / n, s.z / n };
        }
        nodes.clear(); polys.clear();
        Node root; root.Min = mn; root.Max = mx; nodes.push_back(root);
        std::vector<int> all(T->P); std::iota(all.begin(), all.end(), 0);
        split(0, 0, mn, mx, all);
    }
    // BuildKDTree :90-139
    void split(int ni, int depth, P3 mn, P3 mx, std::vector<int>& list) {
        if (depth >= maxDepth || (int)list.size() <= maxPolys) {
            nodes[ni].list_off = (uint32_t)polys.size(); nodes[ni].list_cnt = (uint32_t)list.size();
            polys.insert(polys.end(), list.begin(), list.end());
            return;
        }
        int axis = depth % 3;
        std::vector<int> sorted(list);
        std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return byint(centroid[a], axis) < byint(centroid[b], axis); });
        int medianIndex = (int)sorted.size() / 2;
        double splitValue = byint(centroid[sorted[medianIndex]], axis);
        nodes[ni].axis = axis; nodes[ni].split = splitValue;
        P3 leftMax = mx, rightMin = mn;
        if (axis == 0) { leftMax.x = splitValue; rightMin.x = splitValue; }
        else if (axis == 1) { leftMax.y = splitValue; rightMin.y = splitValue; }
        else { leftMax.z = splitValue; rightMin.z = splitValue; }
        int li = (int)nodes.size(); { Node n; n.Min = mn; n.Max = leftMax; nodes.push_back(n); }
        int ri = (int)nodes.size(); { Node n; n.Min = rightMin; n.Max = mx; nodes.push_back(n); }
        nodes[ni].left = li; nodes[ni].right = ri;
        std::vector<int> L, Rr;
        for (int p : sorted) {
            const double* P = T->PV(p); int n = T->vcount[p];
            bool anyle = false, anygt = false;
            for (int k = 0; k < n; ++k) { double cv = P[3 * k + axis]; if (cv <= splitValue) anyle = true; if (cv > splitValue) anygt = true; }
            if (anyle) L.push_back(p);
            if (anygt) Rr.push_back(p);
        }
        list.clear(); list.shrink_to_fit(); sorted.clear(); sorted.shrink_to_fit();
        split(li, depth + 1, mn, leftMax, L);
        split(ri, depth + 1, rightMin, mx, Rr);
    }

    // Shoot :198-361 -- exhaustive DFS, both children always pushed.
    int Shoot(Ray& ray, XEvent& ev, int o1, int o2, int32_t* mailbox, Counters& c) const override {
        ev = XEvent();
        bool hit = false; double closestT = DBL_MAX;
        std::vector<int> stack; stack.reserve(128);
        stack.push_back(0);
        while (!stack.empty()) {
            const Node& cur = nodes[stack.back()]; stack.pop_back();
            ++c.cells;
            if (cur.left < 0 && cur.right < 0) {
                for (uint32_t k = cur.list_off; k < cur.list_off + cur.list_cnt; ++k) {
                    int p = (int)polys[k];
                    ++c.entries;
                    if (p == o1 || p == o2) continue;
                    if (mailbox[p] == ray.Ray_ID) continue;
                    mailbox[p] = ray.Ray_ID;
                    ++c.tests;
                    P3 X; double u, v, t;
                    if (intersect_slow(*T, p, ray, X, u, v, t) && t > 0.0000000001) {
                        if (t < closestT) {
                            closestT = t;
                            ev.Hit = true; ev.X = X; ev.u = u; ev.v = v; ev.t = t; ev.Poly_id = p;
                            hit = true;
                        }
                    }
                }
            } else {
                int first, second;
                const double o[3] = { ray.x, ray.y, ray.z }, d[3] = { ray.dx, ray.dy, ray.dz };
                const double mn[3] = { cur.Min.x, cur.Min.y, cur.Min.z }, mx[3] = { cur.Max.x, cur.Max.y, cur.Max.z };
                int a = cur.axis, b1 = (a == 0) ? 1 : 0, b2 = (a == 2) ? 1 : 2;   // the two other axes in x<y<z order
                double side = o[a] - cur.split;
                double tSplit = -side / d[a];
                double s1 = o[b1] + tSplit * d[b1];
                double s2 = o[b2] + tSplit * d[b2];
                bool inside = (s1 <= mx[b1] && s1 >= mn[b1] && s2 <= mx[b2] && s2 >= mn[b2]);
                if (inside) { if (side >= 0) { first = cur.right; second = cur.left; } else { first = cur.left; second = cur.right; } }
                else        { if (side >= 0) { first = cur.left; second = cur.right; } else { first = cur.right; second = cur.left; } }
                stack.push_back(second);
                stack.push_back(first);
            }
        }
        if (hit) { ++c.hits; return 1; }
        return 0;
    }
};

// ---------------------------------------------------------------------------
// Batch drivers
// ---------------------------------------------------------------------------
static void shoot_range(const Partition* part, int64_t i0, int64_t i1, double* o, const double* d,
                        const int32_t* o1, const int32_t* o2, const int32_t* ray_id,
                        double* t, double* xyz, int32_t* pid, double* uv, Counters& c) {
    std::vector<int32_t> mailbox(part->T->P, 0);   // Poly_Ray_ID slot, zero-initialised (Voxel_Grid.cs:54-62)
    for (int64_t i = i0; i < i1; ++i) {
        Ray R = { o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], ray_id ? ray_id[i] : (int32_t)(i + 1) };
        XEvent ev;
        int st = part->Shoot(R, ev, o1 ? o1[i] : -1, o2 ? o2[i] : -1, mailbox.data(), c);
        o[3 * i] = R.x; o[3 * i + 1] = R.y; o[3 * i + 2] = R.z;    // Ray is a class: caller sees the moved origin
        t[i] = ev.t; xyz[3 * i] = ev.X.x; xyz[3 * i + 1] = ev.X.y; xyz[3 * i + 2] = ev.X.z;
        pid[i] = (st == -2) ? -2 : ev.Poly_id;
        if (uv) { uv[2 * i] = ev.u; uv[2 * i + 1] = ev.v; }
    }
}

// Harness-defined specular chain (Hare itself has no reflection):
//   n = Normal[poly]; k = 2*((dx*nx)+(dy*ny)+(dz*nz)); d' = d - k*n; o' = X_Point; poly_origin1 = poly
static void chain_range(const Partition* part, int64_t i0, int64_t i1, const double* o, const double* d, int order,
                        int32_t* ev_pid, double* ev_t, double* fin_o, double* fin_d, int32_t* nb, Counters& c,
                        double* ev_xyz = nullptr, double* ev_uv = nullptr) {
    std::vector<int32_t> mailbox(part->T->P, 0);
    const Topology& T = *part->T;
    for (int64_t i = i0; i < i1; ++i) {
        Ray R = { o[3 * i], o[3 * i + 1], o[3 * i + 2], d[3 * i], d[3 * i + 1], d[3 * i + 2], 0 };
        int o1 = -1, b = 0;
        for (; b < order; ++b) {
            R.Ray_ID = (int32_t)((i * order + b) % 2147483646) + 1;
            XEvent ev;
            int st = part->Shoot(R, ev, o1, -1, mailbox.data(), c);
            if (ev_pid) ev_pid[i * order + b] = (st == -2) ? -2 : ev.Poly_id;
            if (ev_t) ev_t[i * order + b] = ev.t;
            if (ev_xyz) { double* q = ev_xyz + 3 * (i * order + b); q[0] = ev.X.x; q[1] = ev.X.y; q[2] = ev.X.z; }   // X_Point (0 on a miss)
            if (ev_uv) { double* q = ev_uv + 2 * (i * order + b); q[0] = ev.u; q[1] = ev.v; }
            if (st != 1) { ++b; break; }
            const double* N = T.NV(ev.Poly_id);
            double k = 2 * ((R.dx * N[0]) + (R.dy * N[1]) + (R.dz * N[2]));
            R.dx = R.dx - k * N[0]; R.dy = R.dy - k * N[1]; R.dz = R.dz - k * N[2];
            R.x = ev.X.x; R.y = ev.X.y; R.z = ev.X.z;
            o1 = ev.Poly_id;
        }
        if (ev_pid) for (int q = b; q < order; ++q) ev_pid[i * order + q] = -3;   // not shot
        if (ev_t) for (int q = b; q < order; ++q) ev_t[i * order + q] = 0;
        if (ev_xyz) for (int q = 3 * b; q < 3 * order; ++q) ev_xyz[3 * i * order + q] = 0;
        if (ev_uv) for (int q = 2 * b; q < 2 * order; ++q) ev_uv[2 * i * order + q] = 0;
        if (fin_o) { fin_o[3 * i] = R.x; fin_o[3 * i + 1] = R.y; fin_o[3 * i + 2] = R.z; }
        if (fin_d) { fin_d[3 * i] = R.dx; fin_d[3 * i + 1] = R.dy; fin_d[3 * i + 2] = R.dz; }
        if (nb) nb[i] = b;   // number of Shoot calls made for this chain
    }
}

template <class F>
static void par_for(int64_t N, int nthreads, Counters* total, F f) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 500) nthreads = 500;   // no_of_boxes mailbox slots (Voxel_Grid.cs:32)
    std::vector<Counters> cs(nthreads);
    std::vector<std::thread> th;
    for (int p = 0; p < nthreads; ++p) {
        int64_t i0 = N * p / nthreads, i1 = N * (p + 1) / nthreads;
        th.emplace_back([&, p, i0, i1] { f(i0, i1, cs[p]); });
    }
    for (auto& t : th) t.join();
    if (total) for (auto& c : cs) { total->cells += c.cells; total->entries += c.entries; total->tests += c.tests; total->hits += c.hits; }
}

}  // namespace ho

// ---------------------------------------------------------------------------
// C interface for ctypes (tests, smoke, bench cpu_baseline)
// ---------------------------------------------------------------------------
using namespace ho;

extern "C" {

void* ho_topology_new(const double* minpt, const double* maxpt) {
    return new Topology({ minpt[0], minpt[1], minpt[2] }, { maxpt[0], maxpt[1], maxpt[2] });
}
int ho_topology_add_polygon(void* h, const double* pts, int n) { return ((Topology*)h)->Add_Polygon(pts, n); }
// bulk: verts P x 4 x 3, vcount P
int ho_topology_add_polygons(void* h, const double* verts, const int32_t* vcount, int64_t P) {
    Topology* T = (Topology*)h;
    for (int64_t i = 0; i < P; ++i) if (T->Add_Polygon(verts + 12 * i, vcount[i]) != 0) return -1;
    return 0;
}
void ho_topology_finish(void* h) { ((Topology*)h)->Finish_Topology(); }
int64_t ho_topology_polygon_count(void* h) { return ((Topology*)h)->P; }
int64_t ho_topology_vertex_count(void* h) { return (int64_t)((Topology*)h)->Vertices_List.size(); }
void ho_topology_get(void* h, double* verts, double* normals, int32_t* vcount, double* minmax) {
    Topology* T = (Topology*)h;
    if (verts) std::memcpy(verts, T->verts.data(), T->verts.size() * 8);
    if (normals) std::memcpy(normals, T->normals.data(), T->normals.size() * 8);
    if (vcount) std::memcpy(vcount, T->vcount.data(), T->vcount.size() * 4);
    if (minmax) { minmax[0] = T->Min.x; minmax[1] = T->Min.y; minmax[2] = T->Min.z; minmax[3] = T->Max.x; minmax[4] = T->Max.y; minmax[5] = T->Max.z; }
}
void ho_topology_free(void* h) { delete (Topology*)h; }

// mode 0: literal flat ctor; 1: polygon-major equivalent; 2: hierarchical ctor (arg = MaxDomain, arg2 = Avg_polys)
void* ho_voxelgrid_new(void* topo, int mode, int arg, int arg2, int nthreads) {
    VoxelGrid* g = new VoxelGrid(); g->T = (Topology*)topo;
    if (mode == 0) g->build_flat(arg, nthreads < 1 ? 1 : nthreads);
    else if (mode == 1) g->build_flat_fast(arg);
    else g->build_hier(arg, arg2, nthreads < 1 ? 1 : nthreads);
    return (Partition*)g;
}
void ho_voxelgrid_info(void* h, double* obox6, double* voxeldims3, int32_t* ct3, int64_t* npairs) {
    VoxelGrid* g = (VoxelGrid*)(Partition*)h;
    obox6[0] = g->OBox.Min.x; obox6[1] = g->OBox.Min.y; obox6[2] = g->OBox.Min.z; obox6[3] = g->OBox.Max.x; obox6[4] = g->OBox.Max.y; obox6[5] = g->OBox.Max.z;
    voxeldims3[0] = g->VoxelDims.x; voxeldims3[1] = g->VoxelDims.y; voxeldims3[2] = g->VoxelDims.z;
    ct3[0] = g->Ct[0]; ct3[1] = g->Ct[1]; ct3[2] = g->Ct[2];
    *npairs = (int64_t)g->cell_poly.size();
}
void ho_voxelgrid_csr(void* h, uint32_t* cell_offset, uint32_t* cell_poly) {
    VoxelGrid* g = (VoxelGrid*)(Partition*)h;
    std::memcpy(cell_offset, g->cell_offset.data(), g->cell_offset.size() * 4);
    std::memcpy(cell_poly, g->cell_poly.data(), g->cell_poly.size() * 4);
}
void* ho_octree_new(void* topo, int maxDepth, int maxPolys) {
    Octree* t = new Octree(); t->T = (Topology*)topo; t->build(maxDepth, maxPolys); return (Partition*)t;
}
void ho_octree_info(void* h, int64_t* nnodes, int64_t* nlist, int64_t* lost) {
    Octree* t = (Octree*)(Partition*)h; *nnodes = (int64_t)t->nodes.size(); *nlist = (int64_t)t->polys.size(); *lost = (int64_t)t->lost;
}
void ho_octree_get(void* h, double* box /*N x 6*/, int32_t* first_child, uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys) {
    Octree* t = (Octree*)(Partition*)h;
    for (size_t i = 0; i < t->nodes.size(); ++i) {
        const auto& n = t->nodes[i];
        box[6 * i] = n.Min.x; box[6 * i + 1] = n.Min.y; box[6 * i + 2] = n.Min.z; box[6 * i + 3] = n.Max.x; box[6 * i + 4] = n.Max.y; box[6 * i + 5] = n.Max.z;
        first_child[i] = n.first_child; list_off[i] = n.list_off; list_cnt[i] = n.list_cnt;
    }
    std::memcpy(polys, t->polys.data(), t->polys.size() * 4);
}
void* ho_kdtree_new(void* topo, int maxDepth, int maxPolys) {
    KDTree* t = new KDTree(); t->T = (Topology*)topo; t->build(maxDepth, maxPolys); return (Partition*)t;
}
void ho_kdtree_info(void* h, int64_t* nnodes, int64_t* nlist) {
    KDTree* t = (KDTree*)(Partition*)h; *nnodes = (int64_t)t->nodes.size(); *nlist = (int64_t)t->polys.size();
}
void ho_kdtree_get(void* h, double* box, double* split, int32_t* axis, int32_t* left, int32_t* right, uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys) {
    KDTree* t = (KDTree*)(Partition*)h;
    for (size_t i = 0; i < t->nodes.size(); ++i) {
        const auto& n = t->nodes[i];
        box[6 * i] = n.Min.x; box[6 * i + 1] = n.Min.y; box[6 * i + 2] = n.Min.z; box[6 * i + 3] = n.Max.x; box[6 * i + 4] = n.Max.y; box[6 * i + 5] = n.Max.z;
        split[i] = n.split; axis[i] = n.axis; left[i] = n.left; right[i] = n.right; list_off[i] = n.list_off; list_cnt[i] = n.list_cnt;
    }
    std::memcpy(polys, t->polys.data(), t->polys.size() * 4);
}
void ho_partition_free(void* h) { delete (Partition*)h; }

// o is IN/OUT (origins moved by Voxel_Grid for outside starts).  counters: cells, entries, tests, hits.
void ho_shoot(void* part, int64_t N, double* o, const double* d, const int32_t* o1, const int32_t* o2, const int32_t* ray_id,
              double* t, double* xyz, int32_t* pid, double* uv, uint64_t* counters, int nthreads) {
    Counters tot;
    par_for(N, nthreads, &tot, [&](int64_t i0, int64_t i1, Counters& c) {
        shoot_range((Partition*)part, i0, i1, o, d, o1, o2, ray_id, t, xyz, pid, uv, c);
    });
    if (counters) { counters[0] = tot.cells; counters[1] = tot.entries; counters[2] = tot.tests; counters[3] = tot.hits; }
}

void ho_reflect_chain(void* part, int64_t N, const double* o, const double* d, int order,
                      int32_t* ev_pid, double* ev_t, double* fin_o, double* fin_d, int32_t* nbounce, uint64_t* counters, int nthreads) {
    Counters tot;
    par_for(N, nthreads, &tot, [&](int64_t i0, int64_t i1, Counters& c) {
        chain_range((Partition*)part, i0, i1, o, d, order, ev_pid, ev_t, fin_o, fin_d, nbounce, c);
    });
    if (counters) { counters[0] = tot.cells; counters[1] = tot.entries; counters[2] = tot.tests; counters[3] = tot.hits; }
}

// the same with the per-bounce X_Point (N x order x 3) and u, v (N x order x 2) streams
void ho_reflect_chain_events(void* part, int64_t N, const double* o, const double* d, int order,
                             int32_t* ev_pid, double* ev_t, double* ev_xyz, double* ev_uv, double* fin_o, double* fin_d, int32_t* nbounce,
                             uint64_t* counters, int nthreads) {
    Counters tot;
    par_for(N, nthreads, &tot, [&](int64_t i0, int64_t i1, Counters& c) {
        chain_range((Partition*)part, i0, i1, o, d, order, ev_pid, ev_t, fin_o, fin_d, nbounce, c, ev_xyz, ev_uv);
    });
    if (counters) { counters[0] = tot.cells; counters[1] = tot.entries; counters[2] = tot.tests; counters[3] = tot.hits; }
}

// direct predicate access for unit tests
int ho_poly_box_overlap(const double* boxmin, const double* boxmax, const double* pts, int n) {
    AABB b({ boxmin[0], boxmin[1], boxmin[2] }, { boxmax[0], boxmax[1], boxmax[2] });
    return b.PolyBoxOverlap(pts, n) ? 1 : 0;
}
double ho_round15(double x) { return Round15(x); }
int ho_intersect(void* topo, int64_t i, const double* o, const double* d, int slow, double* out /*t,x,y,z,u,v*/) {
    Topology* T = (Topology*)topo;
    Ray R = { o[0], o[1], o[2], d[0], d[1], d[2], 1 };
    if (slow) { P3 X = { 0, 0, 0 }; double u, v, t; bool h = intersect_slow(*T, i, R, X, u, v, t); out[0] = t; out[1] = X.x; out[2] = X.y; out[3] = X.z; out[4] = u; out[5] = v; return h; }
    double x, y, z, t; bool h = intersect_fast(*T, i, R, x, y, z, t); out[0] = t; out[1] = x; out[2] = y; out[3] = z; out[4] = 0; out[5] = 0; return h;
}

}  // extern "C"
