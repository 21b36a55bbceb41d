// shoot.cuh -- device-side layout of the three Hare partitions and the pieces their traversal kernels share:
//   VGrid   Voxel_Grid   Voxel_Grid.cs:351-552      (vg_wave.cuh)
//   OctDev  Octree       "Octree - alt.cs":159-306  (oct_wave.cuh)
//   KdDev   KDTree       KDTree.cs:198-361          (kd_wave.cuh)
// polygon record fetch, the conservative padded-box cull, counters and the kernels' output block.
// All arithmetic that feeds a comparison or an output is written in the reference's
// operation order and compiled with -fmad=false.
#pragma once
#include <cstdint>
#include <cstring>
#include "hare_math.cuh"

namespace hare {

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

// Conservative reject (not part of the reference; it never changes a result, it only spares the 128-byte fetch and the exact FP64
// test): true only when the ray's supporting line misses the polygon's padded axis-aligned bounding box (slab test).  For the
// wall / floor / seat rectangles of a hall the box is flat, so the test is nearly as sharp as the exact one.  lo/hi are the FP32 box corners, padded by 1e-3 m + 1e-5 of the extent + 1e-6 |coordinate|
// and rounded outwards (hare_box_pad); p = an FP32 ray point inside the model, d = the FP32 direction.  A ray that hits the polygon at
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
    const float4* __restrict__ cbox;  // per run of HARE_OCT_CHUNK consecutive leaf-list entries (leaf.pad = first run): padded FP32 bounding box (lo, hi), see cull_box()
    const float4* __restrict__ gbox;  // boxes of groups of 8 runs (a leaf's first run is a multiple of 8): gbox[run / 8]
    const float4* __restrict__ pbox;  // per polygon: padded box (lo, hi), indexed by polygon id
    const float4* __restrict__ nbox;  // per node: padded FP32 box of every polygon listed below it (lo, hi); a ray whose line
                                      // misses it cannot be affected by the subtree, which is then not entered
    int depth;   // deepest level (root = 0)
    int regular; // 1: every child box is the reference's function of its parent's box ("Octree - alt.cs":99-114), see oct_child_filter()
};

#define HARE_OCT_MAXLVL 20
#define HARE_OCT_CHUNK 8

// ---------------------------------------------------------------------------------------
// KDTree
// ---------------------------------------------------------------------------------------
struct alignas(64) KdNode {
    double mnx, mny, mnz, mxx, mxy, mxz;
    double split;             // internal: SplitValue; leaf: bits = (list_off | list_cnt << 32)
    int32_t left;             // internal: index of Left (Right = left + 1); leaf: -1
    int32_t axis;             // 0,1,2
};

// Hot form of a kd node, 32 B: the walk only needs a conservative box, a split to order the children by and the child / list
// links.  FP32, box rounded outwards and padded (hare_box_pad): the pruned walk may visit nodes in any order and may only skip a
// subtree that provably holds no hit at t <= closest -- exactness lives in the polygon tests and in the tie rule, which read the
// FP64 records (KdNode, ref_box).
//   internal: a = split as float bits, b = left << 2 | axis (0..2)        leaf: a = list offset, b = list count << 2 | 3
struct alignas(32) KdNodeC { float mnx, mny, mnz, mxx, mxy, mxz; uint32_t a, b; };

// The walk's record: a kd node together with its (up to four) grandchildren -- the two levels below it collapsed into one 128-byte
// fetch, so that a ray pays one dependent memory round trip per TWO levels of the reference's binary tree.  Entry k: padded box,
//   b & 3 == 0: internal node, b >> 2 = its index (and that of its own KdWide record);  b & 3 == 3: leaf, a = list offset, b >> 2 = count;
//   b & 3 == 2: unused entry.  Indexed by the binary node's index (only internal nodes have a record).
struct alignas(128) KdWide { KdNodeC e[4]; };

struct KdDev {
    const KdWide* __restrict__ wide;  // the walk's records (kd_wave.cuh)
    const KdNodeC* __restrict__ hot;  // FP32 form of single nodes (root look-up)
    const KdNode* __restrict__ nodes;
    const uint32_t* __restrict__ lists;
    const float4* __restrict__ lbox;  // per leaf-list entry: its polygon's padded box, polygon id in lo.w (cull_box)
    int depth;
    const double* __restrict__ ref_box;   // per node: the REFERENCE's node box (Min xyz, Max xyz; KDTree.cs:68-83, 107-121) -- the device nodes hold
                                          // it intersected with the content box; only the tie rule (kd_dfs_before) reads this cold table
};

#define HARE_KD_MAXSTACK 64
#define HARE_KD_PAD 1e-5   /* spatial inflation of node boxes for the conservative prune */

// output block of the traversal kernels: batched Shoot (t .. omoved) or reflection chains (ev_pid .. total_shots)
struct WalkOut {
    double* __restrict__ t; double* __restrict__ xyz; int32_t* __restrict__ pid; double* __restrict__ uv; double* __restrict__ omoved;   // Shoot
    int32_t* __restrict__ ev_pid; double* __restrict__ ev_t; double* __restrict__ fin_o; double* __restrict__ fin_d;                   // chain
    int32_t* __restrict__ nshots; unsigned long long* __restrict__ total_shots;
    unsigned long long* __restrict__ counters;
    double* __restrict__ ev_xyz; double* __restrict__ ev_uv;   // chain: per-bounce X_Point (N x order x 3) and u, v (N x order x 2), optional
};

// Per-bounce X_Point / (u, v) rows of a chain.  Row `bounce` of chain `ray`; a miss writes zeros like X_Event() does; rows of Shoots
// that never happen (the chain ended) are zeroed by chain_rows_clear.
HD void chain_row_xyz(const WalkOut& out, long long ray, int order, uint32_t bounce, bool h, double x, double y, double z) {
    if (!out.ev_xyz) return;
    double* q = out.ev_xyz + 3 * (ray * order + (long long)bounce);
    q[0] = h ? x : 0.0; q[1] = h ? y : 0.0; q[2] = h ? z : 0.0;
}
HD void chain_row_uv(const WalkOut& out, long long ray, int order, uint32_t bounce, double u, double v) {
    if (!out.ev_uv) return;
    double* q = out.ev_uv + 2 * (ray * order + (long long)bounce);
    q[0] = u; q[1] = v;
}
HD void chain_rows_clear(const WalkOut& out, long long ray, int order, int from) {
    for (int q = from; q < order; ++q) { chain_row_xyz(out, ray, order, (uint32_t)q, false, 0, 0, 0); chain_row_uv(out, ray, order, (uint32_t)q, 0.0, 0.0); }
}

#if defined(__CUDACC__)
// per-warp reduction of the walk counters into counters[0..3] (cells or nodes, entries, tests, hits)
template <bool COUNT>
__device__ __forceinline__ void flush_counters(const CntT<COUNT>& c, unsigned long long* __restrict__ counters) {
    if (!COUNT) return;
    unsigned int v[4] = { c.cells, c.entries, c.tests, c.hits };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned int s = v[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(counters + k, (unsigned long long)s);
    }
}
#endif

}  // namespace hare
