// shoot.cuh -- per-ray traversal device functions for the three Hare partitions.
//
// Each shoot_one() reproduces the result semantics of one reference Shoot():
//   VGrid   Voxel_Grid.Shoot   Voxel_Grid.cs:351-552   (3D-DDA, carried candidate, in-voxel accept)
//   OctDev  Octree.Shoot       "Octree - alt.cs":159-306 (far-first DFS, early return)
//   KdDev   KDTree.Shoot       KDTree.cs:198-361       (global closest hit; pruned walk, see below)
// All arithmetic that feeds a comparison or an output is written in the reference's
// operation order and compiled with -fmad=false.
#pragma once
#include <cstdint>
#include <cstring>
#include "hare_math.cuh"

namespace hare {

struct Event { double t, x, y, z, u, v; int32_t pid; };

struct Cnt { unsigned long long cells, entries, tests, hits; };

template <bool COUNT> struct CntT {
    unsigned int cells = 0, entries = 0, tests = 0, hits = 0;
    HD void cell() { if (COUNT) ++cells; }
    HD void entry() { if (COUNT) ++entries; }
    HD void test() { if (COUNT) ++tests; }
    HD void hit() { if (COUNT) ++hits; }
    HD void cull() {}
};

// read-only global load: LDG.E.CONSTANT on the device; a plain load when a traversal function is compiled
// for the host (tests/emu replays the wavefront scheduler of vg_wave.cuh on the CPU)
template <class T> HD T hare_ldg(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// Fetch one 128-byte polygon record as eight 128-bit read-only loads (LDG.E.128.CONSTANT),
// all independent so they are in flight together.
HD void load_poly(const PolyRec* __restrict__ polys, uint32_t i, double* P) {
    const double2* src = reinterpret_cast<const double2*>(polys + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double2 q = hare_ldg(src + k);
        P[2 * k] = q.x; P[2 * k + 1] = q.y;
    }
}

// Conservative reject: true only when the ray's supporting line stays outside the polygon's padded
// bounding sphere, i.e. when no triangle test on that polygon can succeed.  Not part of the reference;
// it never changes a result, it only spares the 128-byte fetch and the exact FP64 test.
// Evaluated in FP32 in a frame local to the current voxel: p = a point of the ray at the voxel (rounded
// to FP32), v = c - p is at most a voxel diagonal plus a polygon radius long, so |v x d|^2 = |v|^2|d|^2 -
// (v.d)^2 has no large-number cancellation.  Error budget: the FP32 evaluation is good to ~3e-7 |v|^2|d|^2,
// covered 13x by the explicit 4e-6 |v|^2 term below (it matters only when a tree leaf lists a polygon far
// from its own box); the 3e-6 m rounding of p and of the stored centre moves the line by < 1e-5 m, far
// inside the 1e-3 m + 1e-5 r by which the host pads the radius.
HD bool cull_sphere(const float4 s, float px, float py, float pz, float dx, float dy, float dz, float dd) {
    const float vx = s.x - px, vy = s.y - py, vz = s.z - pz;
    const float vd = fmaf(vx, dx, fmaf(vy, dy, vz * dz));
    const float vv = fmaf(vx, vx, fmaf(vy, vy, vz * vz));
    return fmaf(-vd, vd, vv * dd) > fmaf(4e-6f, vv, s.w * s.w) * dd;   // |v x d|^2 > (r^2 + margin) |d|^2   (NaN -> keep)
}

// Second conservative reject: true only when the ray's supporting line misses the polygon's padded axis-aligned
// bounding box (slab test).  For the wall / floor / seat rectangles of a hall the box is flat, so this is far
// tighter than the sphere.  lo/hi are the FP32 box corners, padded by 1e-3 m + 1e-5 of the extent + 1e-6 |coordinate|
// and rounded outwards (hare_box_pad); p, d as in cull_sphere (FP32 ray point near the voxel, FP32 direction).  A ray that hits the polygon at
// q has q at least 1e-3 m inside the padded box on every axis, i.e. its parameter lies >= 1e-3/|d_a| inside each
// slab interval, while the FP32 evaluation of the interval ends (lo/d - p/d, one FMA; p, d and 1/d rounded to FP32, p anywhere
// on the ray inside the model) is good to ~6e-7 |coordinate| / |d_a|: the
// intervals computed here all contain it, so it is never rejected.  A zero direction component is replaced by
// 1e-30 (cull_rcp; the slab then spans (-huge, +huge) when p is inside it and is empty when p is outside).  NaN -> keep.
HD double hare_box_pad(double l, double h) { return 1e-3 + 1e-5 * (h - l) + 1e-6 * fmax(fabs(l), fabs(h)); }

HD bool cull_box(const float4 lo, const float4 hi, float pxi, float pyi, float pzi, float ix, float iy, float iz) {   // ix = 1/dx, pxi = px/dx ...
    const float ax = fmaf(lo.x, ix, -pxi), bx = fmaf(hi.x, ix, -pxi);
    const float ay = fmaf(lo.y, iy, -pyi), by = fmaf(hi.y, iy, -pyi);
    const float az = fmaf(lo.z, iz, -pzi), bz = fmaf(hi.z, iz, -pzi);
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    return tn > tf;
}

// 1/x for the culls: one MUFU.RCP (1 ulp) instead of the IEEE division sequence; a zero or denormal x becomes 1e-30
HD float cull_rcp(float x) {
    x = fabsf(x) < 1e-30f ? 1e-30f : x;
#if defined(__CUDA_ARCH__)
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}

// bit casts between a polygon id and the float lane it rides in (VGrid::lbox)
HD uint32_t hare_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
HD float hare_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// ---------------------------------------------------------------------------------------
// Voxel_Grid
// ---------------------------------------------------------------------------------------
struct VGrid {
    double ominx, ominy, ominz, omaxx, omaxy, omaxz;   // OBox
    double vdx, vdy, vdz;                              // VoxelDims
    int nx, ny, nz;
    const uint2* __restrict__ cells;        // (offset, count) per cell, index ((x*ny+y)*nz+z)
    const uint32_t* __restrict__ cell_poly; // ascending polygon indices per cell
    const uint32_t* __restrict__ occ;       // 1 bit per cell: list non-empty
    const uint32_t* __restrict__ occp;      // 1 bit per cell of the grid padded by one voxel on every side: list non-empty, or border (vg_wave.cuh)
    const float4* __restrict__ sph;         // per polygon: padded bounding sphere (cx, cy, cz, r), see cull_sphere()
    const float4* __restrict__ lbox;        // per LIST ENTRY (cell_poly order): padded FP32 bounding box (lo.xyz, polygon id in lo.w; hi.xyz), or null -- cull_box()
};

#define HARE_EPS 0.001   /* Voxel_Grid.Epsilon, Voxel_Grid.cs:39 */

// Voxels[X,Y,Z].Min / .Max on one axis: (X*vd - eps) + omin , ((X+1)*vd + eps) + omin   Voxel_Grid.cs:283-285
HD double vox_min(int X, double vd, double omin) { return ((double)X * vd - HARE_EPS) + omin; }
HD double vox_max(int X, double vd, double omin) { return ((double)(X + 1) * vd + HARE_EPS) + omin; }

// AABB.Intersect(ref Ray, ref tmin)  AABB_Main.cs:173-260 on OBox; moves the origin.
HD bool obox_enter(const VGrid& g, Ray3& R, double& tmin) {
    tmin = 0;
    double tmax = DBL_MAX;
    const double o[3] = { R.x, R.y, R.z }, d[3] = { R.dx, R.dy, R.dz };
    const double mn[3] = { g.ominx, g.ominy, g.ominz }, mx[3] = { g.omaxx, g.omaxy, g.omaxz };
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (d[a] == 0.0) {   // Math.Abs(d) < double.Epsilon
            if (o[a] < mn[a] || o[a] > mx[a]) return false;
        } else {
            double ood = (1 / d[a]);
            double t1 = (mn[a] - o[a]) * ood;
            double t2 = (mx[a] - o[a]) * ood;
            if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
            tmin = net_max(tmin, t1);
            tmax = net_min(tmax, t2);
            if (tmin > tmax) return false;
        }
    }
    R.x = R.x + R.dx * tmin;
    R.y = R.y + R.dy * tmin;
    R.z = R.z + R.dz * tmin;
    return true;
}

// returns 1 hit, 0 miss, -2 the reference throws (entry voxel still outside the grid)
template <bool COUNT>
__device__ __forceinline__ int shoot_one(const VGrid& g, const PolyRec* __restrict__ polys, Ray3& R,
                                         int o1, int o2, bool blind, Event& ev, CntT<COUNT>& c) {
    ev.t = 0; ev.x = 0; ev.y = 0; ev.z = 0; ev.u = 0; ev.v = 0; ev.pid = -1;
    int X = floor_to_int((R.x - g.ominx) / g.vdx);
    int Y = floor_to_int((R.y - g.ominy) / g.vdy);
    int Z = floor_to_int((R.z - g.ominz) / g.vdz);
    double t_start = 0;
    if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) {
        if (!obox_enter(g, R, t_start)) return 0;
        X = floor_to_int((R.x - g.ominx + R.dx * 1E-6) / g.vdx);
        Y = floor_to_int((R.y - g.ominy + R.dy * 1E-6) / g.vdy);
        Z = floor_to_int((R.z - g.ominz + R.dz * 1E-6) / g.vdz);
        if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) { ev.pid = -2; return -2; }
    }
    int stepX, stepY, stepZ;
    double tMaxX, tMaxY, tMaxZ, tDeltaX, tDeltaY, tDeltaZ;
    if (R.dx < 0) { stepX = -1; tMaxX = (vox_min(X, g.vdx, g.ominx) - R.x) / R.dx; tDeltaX = g.vdx / R.dx * -1.0; }
    else          { stepX = 1;  tMaxX = (vox_max(X, g.vdx, g.ominx) - R.x) / R.dx; tDeltaX = g.vdx / R.dx * 1.0; }
    if (R.dy < 0) { stepY = -1; tMaxY = (vox_min(Y, g.vdy, g.ominy) - R.y) / R.dy; tDeltaY = g.vdy / R.dy * -1.0; }
    else          { stepY = 1;  tMaxY = (vox_max(Y, g.vdy, g.ominy) - R.y) / R.dy; tDeltaY = g.vdy / R.dy * 1.0; }
    if (R.dz < 0) { stepZ = -1; tMaxZ = (vox_min(Z, g.vdz, g.ominz) - R.z) / R.dz; tDeltaZ = g.vdz / R.dz * -1.0; }
    else          { stepZ = 1;  tMaxZ = (vox_max(Z, g.vdz, g.ominz) - R.z) / R.dz; tDeltaZ = g.vdz / R.dz * 1.0; }

    bool have = false;
    double tmin = DBL_MAX, bx = 0, by = 0, bz = 0;
    int pid = -1;
    uint32_t last = 0xffffffffu;   // last polygon tested: a one-entry mailbox (re-tests never change the result)
    while (true) {
        c.cell();
        const uint32_t ci = ((uint32_t)X * (uint32_t)g.ny + (uint32_t)Y) * (uint32_t)g.nz + (uint32_t)Z;
        uint2 h = __ldg(g.cells + ci);
        if (blind) h.y = 0;   // Ray_ID == 0: the zero-initialised mailbox rejects every polygon (Voxel_Grid.cs:54-62, 478-480)
        for (uint32_t k = 0; k < h.y; ++k) {
            const uint32_t i = __ldg(g.cell_poly + h.x + k);
            c.entry();
            if ((int)i == o1 || (int)i == o2) continue;
            if (i == last || (int)i == pid) continue;
            last = i;
            c.test();
            double P[16], t, u, v;
            load_poly(polys, i, P);
            if (poly_intersect<false>(P, R, t, u, v) && t > 0.0000000001) {
                if (t < tmin) {
                    bx = R.x + R.dx * t; by = R.y + R.dy * t; bz = R.z + R.dz * t;
                    tmin = t; pid = (int)i; have = true;
                }
            }
        }
        if (have) {   // Voxels[X,Y,Z].IsPointInBox(Xpt)  AABB_Main.cs:75-84
            if (!(bx < vox_min(X, g.vdx, g.ominx)) && !(by < vox_min(Y, g.vdy, g.ominy)) && !(bz < vox_min(Z, g.vdz, g.ominz)) &&
                !(bx > vox_max(X, g.vdx, g.ominx)) && !(by > vox_max(Y, g.vdy, g.ominy)) && !(bz > vox_max(Z, g.vdz, g.ominz))) {
                ev.t = tmin + t_start; ev.x = bx; ev.y = by; ev.z = bz; ev.pid = pid;
                c.hit();
                return 1;
            }
        }
        if (tMaxX < tMaxY) {
            if (tMaxX < tMaxZ) { X += stepX; if (X < 0 || X >= g.nx) return 0; tMaxX = tMaxX + tDeltaX; }
            else               { Z += stepZ; if (Z < 0 || Z >= g.nz) return 0; tMaxZ = tMaxZ + tDeltaZ; }
        } else {
            if (tMaxY < tMaxZ) { Y += stepY; if (Y < 0 || Y >= g.ny) return 0; tMaxY = tMaxY + tDeltaY; }
            else               { Z += stepZ; if (Z < 0 || Z >= g.nz) return 0; tMaxZ = tMaxZ + tDeltaZ; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// Octree
// ---------------------------------------------------------------------------------------
struct alignas(64) OctNode {
    double mnx, mny, mnz, mxx, mxy, mxz;
    int32_t first_child;      // -1: leaf
    uint32_t list_off, list_cnt;
    uint32_t pad;
};

struct OctDev {
    const OctNode* __restrict__ nodes;
    const uint32_t* __restrict__ lists;
    const float4* __restrict__ sph;   // padded bounding spheres, see cull_sphere()
    const float4* __restrict__ csph;  // one sphere per run of HARE_OCT_CHUNK consecutive leaf-list entries (leaf.pad = first chunk)
    const float4* __restrict__ cbox;  // the same runs' padded FP32 bounding boxes (lo, hi), see cull_box()
    const float4* __restrict__ gbox;  // boxes of groups of 8 runs (a leaf's first run is a multiple of 8): gbox[run / 8]
    const float4* __restrict__ lbox;  // per leaf-list entry: its polygon's padded box, polygon id in lo.w (or null, see pbox)
    const float4* __restrict__ pbox;  // per polygon: padded box (lo, hi), indexed by polygon id
    const float4* __restrict__ nbox;  // per node: padded FP32 box of every polygon listed below it (lo, hi); a ray whose line
                                      // misses it cannot be affected by the subtree, which is then not entered
    int depth;   // deepest level (root = 0)
};

#define HARE_OCT_MAXLVL 20
#define HARE_OCT_CHUNK 8

__device__ __forceinline__ void oct_interval(const OctNode* __restrict__ n, const Ray3& R, double ix, double iy, double iz,
                                             double& lo, double& hi) {
    const double2* q = reinterpret_cast<const double2*>(n);
    const double2 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2);   // mnx,mny | mnz,mxx | mxy,mxz
    double tx0 = (a.x - R.x) * ix, tx1 = (b.y - R.x) * ix;
    double ty0 = (a.y - R.y) * iy, ty1 = (cc.x - R.y) * iy;
    double tz0 = (b.x - R.z) * iz, tz1 = (cc.y - R.z) * iz;
    if (ix < 0) { double s = tx0; tx0 = tx1; tx1 = s; }
    if (iy < 0) { double s = ty0; ty0 = ty1; ty1 = s; }
    if (iz < 0) { double s = tz0; tz0 = tz1; tz1 = s; }
    lo = net_max(net_max(tx0, ty0), tz0);
    hi = net_min(net_min(tx1, ty1), tz1);
}

// Same arithmetic for a ray whose six components are finite: the node boxes and the guarded reciprocals
// are finite and non-zero, so no NaN can arise and Math.Max/Min reduce to DMNMX (which orders -0 < +0 as
// .NET does).  Rays with a NaN/Inf component take the fully guarded version above.
__device__ __forceinline__ void oct_interval_finite(const OctNode* __restrict__ n, const Ray3& R, double ix, double iy, double iz,
                                                    double& lo, double& hi) {
    const double2* q = reinterpret_cast<const double2*>(n);
    const double2 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2);
    double tx0 = (a.x - R.x) * ix, tx1 = (b.y - R.x) * ix;
    double ty0 = (a.y - R.y) * iy, ty1 = (cc.x - R.y) * iy;
    double tz0 = (b.x - R.z) * iz, tz1 = (cc.y - R.z) * iz;
    if (ix < 0) { double s = tx0; tx0 = tx1; tx1 = s; }
    if (iy < 0) { double s = ty0; ty0 = ty1; ty1 = s; }
    if (iz < 0) { double s = tz0; tz0 = tz1; tz1 = s; }
    lo = fmax(fmax(tx0, ty0), tz0);
    hi = fmin(fmin(tx1, ty1), tz1);
}

template <bool COUNT>
__device__ __forceinline__ int shoot_one(const OctDev& T, const PolyRec* __restrict__ polys, Ray3& R,
                                         int o1, int o2, bool /*blind: the octree mailbox is commented out, :221-222*/, Event& ev, CntT<COUNT>& c) {
    ev.t = 0; ev.x = 0; ev.y = 0; ev.z = 0; ev.u = 0; ev.v = 0; ev.pid = -1;
    const double ix = fabs(R.dx) > 1e-16 ? 1.0 / R.dx : 1e16;
    const double iy = fabs(R.dy) > 1e-16 ? 1.0 / R.dy : 1e16;
    const double iz = fabs(R.dz) > 1e-16 ? 1.0 / R.dz : 1e16;
    double ca, cb;
    oct_interval(T.nodes, R, ix, iy, iz, ca, cb);
    if (cb < ca || cb < 0) return 0;
    // ComputeTraversalOrder :286-306 -- order[q] = near->far octant sequence; with
    // s = (dx<0)<<2 | (dy<0)<<1 | (dz<0) it is simply order[q] = q ^ s.
    const int s = (R.dx >= 0 ? 0 : 4) | (R.dy >= 0 ? 0 : 2) | (R.dz >= 0 ? 0 : 1);

    // The reference's LIFO of (node, tmin, tmax) pops a node's pushed children in reverse push
    // order, depth first.  A frame per level (first child, parent interval, next q to "pop")
    // replays that order lazily; the push-time filter depends only on the parent interval.
    int fchild[HARE_OCT_MAXLVL]; double fa[HARE_OCT_MAXLVL], fb[HARE_OCT_MAXLVL]; int fq[HARE_OCT_MAXLVL];
    const float fdx = (float)R.dx, fdy = (float)R.dy, fdz = (float)R.dz, fdd = fmaf(fdx, fdx, fmaf(fdy, fdy, fdz * fdz));
    uint32_t last = 0xffffffffu;
    int sp = -1;
    int cur = 0;
    bool have_cur = true, hit = false;
    double closest = DBL_MAX;
    while (true) {
        if (have_cur) {
            have_cur = false;
            if (!(cb < ca || cb < 0) && !(hit && closest <= ca)) {
                c.cell();
                const OctNode* n = T.nodes + cur;
                const uint4 m = __ldg(reinterpret_cast<const uint4*>(n) + 3);   // first_child, list_off, list_cnt, pad
                if ((int)m.x < 0) {
                    // voxel-local FP32 frame for the conservative sphere reject (cull_sphere): ray point at the leaf entry
                    const double te = ca > 0.0 ? ca : 0.0;
                    const float px = (float)fma(R.dx, te, R.x), py = (float)fma(R.dy, te, R.y), pz = (float)fma(R.dz, te, R.z);
                    for (uint32_t k = 0; k < m.z; ++k) {
                        const uint32_t i = __ldg(T.lists + m.y + k);
                        c.entry();
                        if ((int)i == o1 || (int)i == o2) continue;
                        // a polygon already tested for this ray (it sits in several leaves) cannot change anything:
                        // its t is not below closestT any more, so neither the update nor the early return fires
                        if (i == last || (int)i == ev.pid) continue;
                        if (cull_sphere(__ldg(T.sph + i), px, py, pz, fdx, fdy, fdz, fdd)) continue;
                        last = i;
                        c.test();
                        double P[16], t, u, v;
                        load_poly(polys, i, P);
                        if (poly_intersect<true>(P, R, t, u, v) && t > 0.0000000001) {
                            if (t < closest) {
                                closest = t; hit = true;
                                ev.t = t; ev.u = u; ev.v = v; ev.pid = (int)i;
                                ev.x = R.x + R.dx * t; ev.y = R.y + R.dy * t; ev.z = R.z + R.dz * t;
                                if (closest <= ca) { c.hit(); return 1; }   // early return :233-237
                            }
                        }
                    }
                } else if (sp + 1 < HARE_OCT_MAXLVL) {
                    ++sp; fchild[sp] = (int)m.x; fa[sp] = ca; fb[sp] = cb; fq[sp] = 7;
                }
            }
        }
        if (sp < 0) break;
        if (fq[sp] < 0) { --sp; continue; }
        const int q = fq[sp]--;
        const int child = fchild[sp] + (q ^ s);
        double lo, hi;
        oct_interval(T.nodes + child, R, ix, iy, iz, lo, hi);
        const double pa = fa[sp], pb = fb[sp];
        if (hi < lo || hi < 0 || lo > pb || hi < pa) continue;
        cur = child; ca = net_max(lo, pa); cb = net_min(hi, pb); have_cur = true;
    }
    if (hit) { c.hit(); return 1; }
    return 0;
}

// ---------------------------------------------------------------------------------------
// KDTree
// ---------------------------------------------------------------------------------------
struct alignas(64) KdNode {
    double mnx, mny, mnz, mxx, mxy, mxz;
    double split;             // internal: SplitValue; leaf: bits = (list_off | list_cnt << 32)
    int32_t left;             // internal: index of Left (Right = left + 1); leaf: -1
    int32_t axis;             // 0,1,2
};

struct KdDev {
    const KdNode* __restrict__ nodes;
    const uint32_t* __restrict__ lists;
    const float4* __restrict__ sph;   // padded bounding spheres, see cull_sphere()
    const float4* __restrict__ lbox;  // per leaf-list entry: its polygon's padded box, polygon id in lo.w (cull_box)
    const float4* __restrict__ pbox;  // per polygon: padded box (lo, hi), indexed by polygon id
    int depth;
};

#define HARE_KD_MAXSTACK 64
#define HARE_KD_PAD 1e-5   /* spatial inflation of node boxes for the conservative prune */

// The reference visits every leaf (both children always pushed, KDTree.cs:355-356) and keeps the
// strict minimum of t over all polygons: its result is the global closest hit.  This walk goes
// near child first and skips a subtree when the ray's parameter interval inside the node's box,
// inflated by HARE_KD_PAD, lies wholly beyond the current closest hit or behind the origin -- such
// a subtree cannot hold a polygon hit at t <= closest.  Results are identical except that among
// polygons hit at exactly equal t the reference keeps the first in its exhaustive DFS order (the
// documented exact-edge ties).
__device__ __forceinline__ bool kd_box_reachable(const double2 a, const double2 b, const double2 cc, const Ray3& R,
                                                 const double* inv, double closest, double& lo) {
    double hi = closest;
    lo = 0.0;
    const double mn[3] = { a.x - HARE_KD_PAD, a.y - HARE_KD_PAD, b.x - HARE_KD_PAD };
    const double mx[3] = { b.y + HARE_KD_PAD, cc.x + HARE_KD_PAD, cc.y + HARE_KD_PAD };
    const double o[3] = { R.x, R.y, R.z }, d[3] = { R.dx, R.dy, R.dz };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0) {
            if (o[k] < mn[k] || o[k] > mx[k]) return false;
        } else {
            double t1 = (mn[k] - o[k]) * inv[k], t2 = (mx[k] - o[k]) * inv[k];
            if (t1 > t2) { double s = t1; t1 = t2; t2 = s; }
            if (t1 > lo) lo = t1;
            if (t2 < hi) hi = t2;
        }
    }
    return !(lo > hi);
}

template <bool COUNT>
__device__ __forceinline__ int shoot_one(const KdDev& T, const PolyRec* __restrict__ polys, Ray3& R,
                                         int o1, int o2, bool blind, Event& ev, CntT<COUNT>& c) {
    ev.t = 0; ev.x = 0; ev.y = 0; ev.z = 0; ev.u = 0; ev.v = 0; ev.pid = -1;
    int stack[HARE_KD_MAXSTACK];
    int sp = 0;
    stack[sp++] = 0;
    bool hit = false;
    double closest = DBL_MAX;
    uint32_t last = 0xffffffffu;
    // reciprocal used only by the conservative prune (its rounding is far inside HARE_KD_PAD)
    const double inv[3] = { 1.0 / R.dx, 1.0 / R.dy, 1.0 / R.dz };
    const float fdx = (float)R.dx, fdy = (float)R.dy, fdz = (float)R.dz, fdd = fmaf(fdx, fdx, fmaf(fdy, fdy, fdz * fdz));
    if (blind) return 0;   // Ray_ID == 0: zero-initialised mailbox rejects every polygon (KDTree.cs:58-66, 224-229)
    while (sp > 0) {
        const int ni = stack[--sp];
        const double2* q = reinterpret_cast<const double2*>(T.nodes + ni);
        const double2 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2), dd = __ldg(q + 3);
        double t_in;
        if (!kd_box_reachable(a, b, cc, R, inv, closest, t_in)) continue;
        c.cell();
        const int left = __double2loint(dd.y), axis = __double2hiint(dd.y);
        if (left < 0) {
            const uint32_t off = (uint32_t)__double2loint(dd.x), cnt = (uint32_t)__double2hiint(dd.x);
            const float px = (float)fma(R.dx, t_in, R.x), py = (float)fma(R.dy, t_in, R.y), pz = (float)fma(R.dz, t_in, R.z);
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t i = __ldg(T.lists + off + k);
                c.entry();
                if ((int)i == o1 || (int)i == o2) continue;
                if (i == last || (int)i == ev.pid) continue;   // mailbox: each polygon counts once
                if (cull_sphere(__ldg(T.sph + i), px, py, pz, fdx, fdy, fdz, fdd)) continue;
                last = i;
                c.test();
                double P[16], t, u, v;
                load_poly(polys, i, P);
                if (poly_intersect<true>(P, R, t, u, v) && t > 0.0000000001) {
                    if (t < closest) {
                        closest = t; hit = true;
                        ev.t = t; ev.u = u; ev.v = v; ev.pid = (int)i;
                        ev.x = R.x + R.dx * t; ev.y = R.y + R.dy * t; ev.z = R.z + R.dz * t;
                    }
                }
            }
        } else {
            // The reference's first/second rule (KDTree.cs:249-353) only fixes the order in which its exhaustive
            // walk meets the leaves; the result is the global minimum of t either way.  Here the child on the
            // origin's side goes first so that the prune above can cut the far side as early as possible.
            const double oa = axis == 0 ? R.x : (axis == 1 ? R.y : R.z);
            const bool right_first = oa > dd.x;
            const int first = right_first ? left + 1 : left, second = right_first ? left : left + 1;
            if (sp + 2 <= HARE_KD_MAXSTACK) { stack[sp++] = second; stack[sp++] = first; }
        }
    }
    if (hit) { c.hit(); return 1; }
    return 0;
}

}  // namespace hare
