// oct_wave.cuh -- K3 (and K5 on an Octree): Hare's Octree.Shoot ("Octree - alt.cs":159-306) under the per-warp
// WAVEFRONT scheduler of vg_wave.cuh.
//
// The first-generation kernel tied one ray to one thread: every trip the warp ran its S / N / C / T phases with
// whichever lanes happened to be in each -- ncu showed 6 of 32 lanes per issued instruction and 6.9 long-scoreboard
// stall cycles per issue (profiles/r1_ncu_octree_oct4_summary.txt).  Here the ray state lives in SHARED MEMORY: every
// warp owns a pool of SLOTS ray slots (structure-of-arrays, 169 bytes per slot) with a one-byte phase tag each; per trip
// the warp counts its slots per phase, picks the fullest phase, compacts up to 32 of its slots onto the lanes and runs
// that ONE phase converged on 32 different rays.
//
// The reference's LIFO of (node, tmin, tmax) is replayed lazily by one frame per level -- first child, the 8-bit mask of
// octants still to pop (near->far order, popped far->near), and the parent's clipped interval (a, b).  Only the TOP
// frame lives in the slot; pushing a level spills the previous top to a per-slot scratch area in global memory (L2
// resident, 24 bytes per level) and popping reads it back, so the slot size does not depend on the tree depth.
//
// Phases (the arithmetic that feeds a comparison or an output is the reference's, operation for operation):
//   SF  finish a Shoot (event out; a chain reflects), fetch a ray, reciprocals + root interval (:165-190), enter the root
//   N   pop octants until a leaf with a non-empty list is entered (<= N_MAX candidates per execution): content-box cull,
//       child interval (:252-266), push-time filter (:268), pop-time prunes (:207-211)
//   G   next group of 64 leaf-list entries: group box, then its 8 chunk boxes            (conservative culls,
//   C   the 8 entries of the lowest surviving chunk: poly_origin / duplicate skip, box    never change a result)
//   T   one exact slow-path test (u, v): strict t < closestT, early `return` when closestT <= nodeTmin (:224-237)
// Entries are taken in stored order and survivors are tested lowest first, so the sequence of closestT updates -- and
// with it the early return and the pop-time prune -- is the reference's.
//
// Every per-slot function is `HD`: tests/emu/ compiles them for the host and replays the scheduler against the oracle.
#pragma once
#include "shoot.cuh"
#include "vg_wave.cuh"   // FIN_*, RayFeed

namespace hare {

enum : uint32_t { OP_SF = 0, OP_N = 1, OP_G = 2, OP_C = 3, OP_T = 4, OP_DONE = 5, OP_COUNT = 5 };

// slot flags: NORAY | fin(2) | hit | sgn(3) | sp(5) | bounce(16)
enum : uint32_t {
    OFL_NORAY = 1u, OFL_FIN_SHIFT = 1, OFL_FIN_MASK = 3u << 1, OFL_HIT = 8u,
    OFL_SGN_SHIFT = 4, OFL_SGN_MASK = 7u << 4, OFL_SP_SHIFT = 7, OFL_SP_MASK = 31u << 7, OFL_BOUNCE_SHIFT = 16
};
// OU_MASKS: qmask (octants of the top frame still to pop) | emask << 8 (surviving chunks of the current group) | bmask << 16 (surviving entries of the
// current chunk) | kc << 24 (which chunk of the group that is) | OM_PEND (OU_PEND holds the id of the lowest surviving entry)
enum : uint32_t { OM_KC_SHIFT = 24, OM_KC_MASK = 7u << 24, OM_PEND = 1u << 27 };
enum { OD_OX, OD_OY, OD_OZ, OD_DX, OD_DY, OD_DZ, OD_IX, OD_IY, OD_IZ, OD_CLOSEST, OD_CA, OD_FA, OD_FB, OD_COUNT };
// OU_LAST doubles as the id of the lowest surviving entry between C and T (OM_PEND): T makes that polygon the last one tested anyway.
// The event's u, v are not kept in the slot: T writes them to the output row whenever closestT improves (the last write is the winner's).
enum { OU_FLAGS, OU_RAY, OU_PID, OU_OR1, OU_OR2, OU_LAST, OU_LPOS, OU_LEND, OU_CIDX, OU_CPOS, OU_FCHILD, OU_MASKS, OU_COUNT };
enum { OF_PX, OF_PY, OF_PZ, OF_COUNT };   // cull_box frame point, already divided by d (FP32)

template <int SLOTS>
struct OctPool {
    double* dbl; uint32_t* u32; float* f32; uint8_t* tag; uint8_t* sel;
    static_assert(SLOTS <= 255, "phase counts are packed into bytes");
    static constexpr size_t BYTES = (size_t)SLOTS * (OD_COUNT * 8 + OU_COUNT * 4 + OF_COUNT * 4 + 1) + 32;
    static constexpr size_t STRIDE = (BYTES + 15) & ~(size_t)15;
    HD void bind(unsigned char* base) {
        dbl = reinterpret_cast<double*>(base);
        u32 = reinterpret_cast<uint32_t*>(base + (size_t)SLOTS * OD_COUNT * 8);
        f32 = reinterpret_cast<float*>(base + (size_t)SLOTS * (OD_COUNT * 8 + OU_COUNT * 4));
        tag = base + (size_t)SLOTS * (OD_COUNT * 8 + OU_COUNT * 4 + OF_COUNT * 4);
        sel = tag + SLOTS;
    }
    HD double& D(int f, int s) const { return dbl[f * SLOTS + s]; }
    HD uint32_t& U(int f, int s) const { return u32[f * SLOTS + s]; }
    HD float& F(int f, int s) const { return f32[f * SLOTS + s]; }
};

// spilled frames: per slot, `depth` levels of (a, b) and (first child, qmask)
struct OctFrames { double2* ab; uint2* cq; int depth; };

HD int hare_ffs(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x);
#else
    return x ? __builtin_ctz(x) + 1 : 0;
#endif
}
HD int hare_fls(uint32_t x) {   // index of the highest set bit (x != 0)
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

HD uint32_t oct_tag(uint32_t fl, uint32_t masks, uint32_t lpos, uint32_t lend) {
    if (fl & OFL_FIN_MASK) return OP_SF;
    if (masks & 0xff0000u) return OP_T;
    if (masks & 0x00ff00u) return OP_C;
    return lpos < lend ? OP_G : OP_N;
}

// the phase with the most ready slots; ties go to the later phase (drain before refill)
HD int oct_pick(const int n[OP_COUNT]) {
    int best = OP_T, bn = n[OP_T];
    if (n[OP_C] > bn) { best = OP_C; bn = n[OP_C]; }
    if (n[OP_G] > bn) { best = OP_G; bn = n[OP_G]; }
    if (n[OP_N] > bn) { best = OP_N; bn = n[OP_N]; }
    if (n[OP_SF] > bn) { best = OP_SF; bn = n[OP_SF]; }
    return bn > 0 ? best : -1;
}

// 16-byte read-only loads of a node record, issued where they are written: the compiler must not sink the rarely used words
// (first_child / list range) below the culls, that would cost the entering lanes a second L2 round trip
HD double2 oct_ld_d2(const double2* q) {
#if defined(__CUDA_ARCH__)
    double2 r; asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(q)); return r;
#else
    return *q;
#endif
}
HD uint4 oct_ld_u4(const uint4* q) {
#if defined(__CUDA_ARCH__)
    uint4 r; asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(q)); return r;
#else
    return *q;
#endif
}
HD uint32_t oct_ld_u32(const uint32_t* q) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(q)); return r;
#else
    return *q;
#endif
}
HD float4 oct_ld_f4(const float4* q) {
#if defined(__CUDA_ARCH__)
    float4 r; asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(q)); return r;
#else
    return *q;
#endif
}

// child interval of a ray with finite components ("Octree - alt.cs":252-266; see oct_interval_finite in shoot.cuh)
HD void oct_interval_of(const double2 a, const double2 b, const double2 cc, double ox, double oy, double oz, double ix, double iy, double iz, double& lo, double& hi) {
    double tx0 = (a.x - ox) * ix, tx1 = (b.y - ox) * ix;      // a = mnx,mny | b = mnz,mxx | cc = mxy,mxz
    double ty0 = (a.y - oy) * iy, ty1 = (cc.x - oy) * iy;
    double tz0 = (b.x - oz) * iz, tz1 = (cc.y - oz) * iz;
    if (ix < 0) { double s = tx0; tx0 = tx1; tx1 = s; }
    if (iy < 0) { double s = ty0; ty0 = ty1; ty1 = s; }
    if (iz < 0) { double s = tz0; tz0 = tz1; tz1 = s; }
    lo = fmax(fmax(tx0, ty0), tz0);   // finite operands: Math.Max/Min == fmax/fmin (DMNMX orders -0 < +0 like .NET)
    hi = fmin(fmin(tx1, ty1), tz1);
}

// near->far permutation of a node's content mask: bit q = octant (q ^ sgn) holds polygons
HD uint32_t oct_perm_mask(uint32_t content, int sgn) {
    // q -> q ^ sgn permutes the bits by up to three swaps: neighbours (sgn & 1), pairs (sgn & 2), nibbles (sgn & 4)
    uint32_t m = content & 0xffu;
    m = (sgn & 1) ? (((m & 0xaau) >> 1) | ((m & 0x55u) << 1)) : m;
    m = (sgn & 2) ? (((m & 0xccu) >> 2) | ((m & 0x33u) << 2)) : m;
    m = (sgn & 4) ? (((m & 0xf0u) >> 4) | ((m & 0x0fu) << 4)) : m;
    return m;
}

// Push-time filter of all eight children at once ("Octree - alt.cs":252-268), from the PARENT's box: BuildOctree derives a child's
// box from its parent's (:99-114: centre = (Max + Min) / 2; Min = (bit ? centre : Min) - 0.1, Max = (bit ? Max : centre) + 0.1), so per
// axis there are only four planes and the eight child intervals share twelve slab parameters.  The expressions are the builder's and
// the interval code's, so every value equals what oct_interval_of() gets from the child's stored box (OctDev::regular: checked
// against the stored boxes when the tree is uploaded; an irregular tree skips this filter and relies on the per-child one).
// Returns the mask in near->far order: bit q = child (q ^ sgn) passes.
HD uint32_t oct_child_filter(const double2 a, const double2 b, const double2 cc, double ox, double oy, double oz, double ix, double iy, double iz,
                             double fa, double fb, int sgn) {
    const double mn[3] = { a.x, a.y, b.x }, mx[3] = { b.y, cc.x, cc.y }, o[3] = { ox, oy, oz }, inv[3] = { ix, iy, iz };
    double p[3][2], q[3][2];   // [axis][half]: entry / exit parameter of the lower (0) and upper (1) half's slab
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double mid = (mx[k] + mn[k]) / 2;
        const double l0 = mn[k] - 0.1, h0 = mid + 0.1, l1 = mid - 0.1, h1 = mx[k] + 0.1;
        double p0 = (l0 - o[k]) * inv[k], q0 = (h0 - o[k]) * inv[k], p1 = (l1 - o[k]) * inv[k], q1 = (h1 - o[k]) * inv[k];
        if (inv[k] < 0) { double t = p0; p0 = q0; q0 = t; t = p1; p1 = q1; q1 = t; }
        p[k][0] = p0; q[k][0] = q0; p[k][1] = p1; q[k][1] = q1;
    }
    // A child is skipped when hi < lo || hi < 0 || lo > b || hi < a with lo = max of its three entries p, hi = min of its three exits q
    // (all finite): that is, when some exit lies below some entry of another axis, below 0 or below a, or some entry above b.  Each of
    // these comparisons is shared by the two or four children on that side, so the eight verdicts take 42 comparisons and no min/max.
    const uint32_t side[3][2] = { { 0x0fu, 0xf0u }, { 0x33u, 0xccu }, { 0x55u, 0xaau } };   // children in the lower / upper half of x, y, z
    uint32_t m = 0xffu;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int h = 0; h < 2; ++h)
            m &= (q[k][h] < 0 || q[k][h] < fa || p[k][h] > fb) ? ~side[k][h] : 0xffu;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int j = (k + 1) % 3;
#pragma unroll
        for (int hk = 0; hk < 2; ++hk)
#pragma unroll
            for (int hj = 0; hj < 2; ++hj)
                m &= (q[k][hk] < p[j][hj] || q[j][hj] < p[k][hk]) ? ~(side[k][hk] & side[j][hj]) : 0xffu;
    }
    return oct_perm_mask(m, sgn);
}

// ---- SF, part 1: the Shoot in slot s is over -> write its event; a chain reflects and goes on, or ends
template <bool CHAIN, bool COUNT, int SLOTS>
HD void octw_finish(const PolyRec* __restrict__ polys, const OctPool<SLOTS>& p, int s, int order, const WalkOut& out,
                    unsigned int& shots, CntT<COUNT>& c) {
    uint32_t fl = p.U(OU_FLAGS, s);
    const uint32_t fin = (fl & OFL_FIN_MASK) >> OFL_FIN_SHIFT;
    if (fin == FIN_RUN) return;
    const bool h = fin == FIN_HIT;
    const double closest = p.D(OD_CLOSEST, s);
    Ray3 R = { p.D(OD_OX, s), p.D(OD_OY, s), p.D(OD_OZ, s), p.D(OD_DX, s), p.D(OD_DY, s), p.D(OD_DZ, s) };
    const int pid = (int)p.U(OU_PID, s);
    const long long ray = (long long)p.U(OU_RAY, s);
    const double bx = R.x + R.dx * closest, by = R.y + R.dy * closest, bz = R.z + R.dz * closest;   // X_Point, Polygons.cs:749
    if (h) c.hit();
    fl &= ~(OFL_FIN_MASK | OFL_HIT);
    if (CHAIN) {
        uint32_t bounce = fl >> OFL_BOUNCE_SHIFT;
        ++shots;
        if (out.ev_pid) out.ev_pid[ray * order + bounce] = h ? pid : -1;
        if (out.ev_t) out.ev_t[ray * order + bounce] = h ? closest : 0.0;
        chain_row_xyz(out, ray, order, bounce, h, bx, by, bz);
        if (!h) chain_row_uv(out, ray, order, bounce, 0.0, 0.0);               // a hit's u, v were written by the test that found it
        ++bounce;
        bool go_on = false;
        if (h) {
            const double* P = polys[pid].v;
            const double nx = hare_ldg(P + 12), ny = hare_ldg(P + 13), nz = hare_ldg(P + 14);
            const double k = 2 * ((R.dx * nx) + (R.dy * ny) + (R.dz * nz));
            R.dx = R.dx - k * nx; R.dy = R.dy - k * ny; R.dz = R.dz - k * nz;
            R.x = bx; R.y = by; R.z = bz;
            p.D(OD_OX, s) = R.x; p.D(OD_OY, s) = R.y; p.D(OD_OZ, s) = R.z;
            p.D(OD_DX, s) = R.dx; p.D(OD_DY, s) = R.dy; p.D(OD_DZ, s) = R.dz;
            p.U(OU_OR1, s) = (uint32_t)pid;
            go_on = (int)bounce < order;
        }
        fl = (fl & 0xffffu) | (bounce << OFL_BOUNCE_SHIFT);
        if (!go_on) {
            for (int q = (int)bounce; q < order; ++q) {
                if (out.ev_pid) out.ev_pid[ray * order + q] = -3;
                if (out.ev_t) out.ev_t[ray * order + q] = 0;
            }
            chain_rows_clear(out, ray, order, (int)bounce);
            if (out.fin_o) { out.fin_o[3 * ray] = R.x; out.fin_o[3 * ray + 1] = R.y; out.fin_o[3 * ray + 2] = R.z; }
            if (out.fin_d) { out.fin_d[3 * ray] = R.dx; out.fin_d[3 * ray + 1] = R.dy; out.fin_d[3 * ray + 2] = R.dz; }
            if (out.nshots) out.nshots[ray] = (int32_t)bounce;
            fl |= OFL_NORAY;
        }
    } else {
        out.pid[ray] = h ? pid : -1;
        if (out.t) out.t[ray] = h ? closest : 0.0;
        if (out.xyz) { out.xyz[3 * ray] = h ? bx : 0.0; out.xyz[3 * ray + 1] = h ? by : 0.0; out.xyz[3 * ray + 2] = h ? bz : 0.0; }
        if (out.uv && !h) { out.uv[2 * ray] = 0.0; out.uv[2 * ray + 1] = 0.0; }     // a hit's u, v were written by the test that found it
        if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }   // the Octree never moves a ray
        fl |= OFL_NORAY;
    }
    p.U(OU_FLAGS, s) = fl;
}

// ---- SF, part 2: put ray number `ray` into slot s
template <int SLOTS>
HD void octw_fetch(const OctPool<SLOTS>& p, int s, long long ray, const double* __restrict__ o, const double* __restrict__ d,
                   const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a) {
    p.D(OD_OX, s) = o[3 * ray]; p.D(OD_OY, s) = o[3 * ray + 1]; p.D(OD_OZ, s) = o[3 * ray + 2];
    p.D(OD_DX, s) = d[3 * ray]; p.D(OD_DY, s) = d[3 * ray + 1]; p.D(OD_DZ, s) = d[3 * ray + 2];
    p.U(OU_OR1, s) = (uint32_t)(o1a ? o1a[ray] : -1);
    p.U(OU_OR2, s) = (uint32_t)(o2a ? o2a[ray] : -1);
    p.U(OU_RAY, s) = (uint32_t)ray;
    p.U(OU_FLAGS, s) = 0u;   // bounce 0, running
}

// ---- SF, part 3: reciprocals, root interval, enter the root   "Octree - alt.cs":165-205.  Returns the slot's new tag.
template <bool COUNT, int SLOTS>
HD uint32_t octw_setup(const OctDev& T, const OctPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(OU_FLAGS, s) & (0xffffu << OFL_BOUNCE_SHIFT);
    const Ray3 R = { p.D(OD_OX, s), p.D(OD_OY, s), p.D(OD_OZ, s), p.D(OD_DX, s), p.D(OD_DY, s), p.D(OD_DZ, s) };
    const double ix = fabs(R.dx) > 1e-16 ? 1.0 / R.dx : 1e16;
    const double iy = fabs(R.dy) > 1e-16 ? 1.0 / R.dy : 1e16;
    const double iz = fabs(R.dz) > 1e-16 ? 1.0 / R.dz : 1e16;
    p.D(OD_IX, s) = ix; p.D(OD_IY, s) = iy; p.D(OD_IZ, s) = iz;
    double ca, cb;
    uint4 m;
    double2 ra, rb, rc;
    {   // root interval with .NET Math.Max / Math.Min (NaN-propagating): the ray may hold anything here
        const double2* q = reinterpret_cast<const double2*>(T.nodes);
        const double2 a = hare_ldg(q), b = hare_ldg(q + 1), cc = hare_ldg(q + 2);
        ra = a; rb = b; rc = cc;
        m = hare_ldg(reinterpret_cast<const uint4*>(T.nodes) + 3);
        double tx0 = (a.x - R.x) * ix, tx1 = (b.y - R.x) * ix;
        double ty0 = (a.y - R.y) * iy, ty1 = (cc.x - R.y) * iy;
        double tz0 = (b.x - R.z) * iz, tz1 = (cc.y - R.z) * iz;
        if (ix < 0) { double t = tx0; tx0 = tx1; tx1 = t; }
        if (iy < 0) { double t = ty0; ty0 = ty1; ty1 = t; }
        if (iz < 0) { double t = tz0; tz0 = tz1; tz1 = t; }
        ca = net_max(net_max(tx0, ty0), tz0);
        cb = net_min(net_min(tx1, ty1), tz1);
    }
    uint32_t fin = FIN_RUN;
    if (cb < ca || cb < 0) fin = FIN_MISS;                       // :185-190
    // A ray with a NaN/Inf component is reported as a miss (documented deviation): the walk uses
    // fmax/fmin, which assume finite operands.  (In the reference every comparison on such a ray is false and every t NaN,
    // which also ends in a miss.)
    if (!(isfinite(R.x) && isfinite(R.y) && isfinite(R.z) && isfinite(R.dx) && isfinite(R.dy) && isfinite(R.dz))) fin = FIN_MISS;
    const int sgn = (R.dx >= 0 ? 0 : 4) | (R.dy >= 0 ? 0 : 2) | (R.dz >= 0 ? 0 : 1);   // ComputeTraversalOrder :286-306: order[q] = q ^ sgn
    fl |= (uint32_t)sgn << OFL_SGN_SHIFT;
    p.D(OD_CLOSEST, s) = DBL_MAX;
    p.U(OU_PID, s) = 0xffffffffu; p.U(OU_LAST, s) = 0xffffffffu;
    uint32_t lpos = 0, lend = 0, masks = 0;
    if (fin == FIN_RUN) {
        // cull_box frame: p = the point where the ray enters the root cube (the origin itself when it starts inside), formed in
        // FP64 and then rounded -- a ray shot from far outside the model must not lose the millimetres the padding allows
        const double te = ca > 0.0 ? ca : 0.0;
        const float fix = cull_rcp((float)R.dx), fiy = cull_rcp((float)R.dy), fiz = cull_rcp((float)R.dz);
        p.F(OF_PX, s) = (float)fma(R.dx, te, R.x) * fix; p.F(OF_PY, s) = (float)fma(R.dy, te, R.y) * fiy; p.F(OF_PZ, s) = (float)fma(R.dz, te, R.z) * fiz;
        // the root is "popped" first (:205): the pop-time prunes cannot fire on it (interval checked above, no hit yet)
        c.cell();
        if ((int)m.x < 0) {       // the root is a leaf
            lpos = m.y; lend = m.y + m.z;
            p.U(OU_CIDX, s) = m.w; p.D(OD_CA, s) = ca;
            p.U(OU_FCHILD, s) = 0;
            if (lpos >= lend) fin = FIN_MISS;     // empty tree
        } else {
            p.U(OU_FCHILD, s) = m.x;
            masks = oct_perm_mask(m.w, sgn);
            if (T.regular) masks &= oct_child_filter(ra, rb, rc, R.x, R.y, R.z, ix, iy, iz, ca, cb, sgn);
            p.D(OD_FA, s) = ca; p.D(OD_FB, s) = cb;
        }
    }
    fl |= fin << OFL_FIN_SHIFT;   // sp = 0
    p.U(OU_FLAGS, s) = fl; p.U(OU_LPOS, s) = lpos; p.U(OU_LEND, s) = lend; p.U(OU_MASKS, s) = masks;
    return oct_tag(fl, masks, lpos, lend);
}

// ---- N: pop octants until a leaf with a non-empty list is entered, the stack runs empty, or N_MAX candidates were looked at
template <bool COUNT, int SLOTS, int N_MAX>
HD uint32_t octw_node(const OctDev& T, const OctFrames& F, size_t gslot, const OctPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(OU_FLAGS, s);
    const double ox = p.D(OD_OX, s), oy = p.D(OD_OY, s), oz = p.D(OD_OZ, s);
    const double ix = p.D(OD_IX, s), iy = p.D(OD_IY, s), iz = p.D(OD_IZ, s);
    const double closest = p.D(OD_CLOSEST, s);
    const bool hit = (fl & OFL_HIT) != 0;
    const int sgn = (int)((fl & OFL_SGN_MASK) >> OFL_SGN_SHIFT);
    int sp = (int)((fl & OFL_SP_MASK) >> OFL_SP_SHIFT);
    double fa = p.D(OD_FA, s), fb = p.D(OD_FB, s);
    uint32_t fchild = p.U(OU_FCHILD, s), qmask = p.U(OU_MASKS, s) & 0xffu;
    const float fdx = (float)p.D(OD_DX, s), fdy = (float)p.D(OD_DY, s), fdz = (float)p.D(OD_DZ, s);
    const float fix = cull_rcp(fdx), fiy = cull_rcp(fdy), fiz = cull_rcp(fdz);
    const float fpx = p.F(OF_PX, s), fpy = p.F(OF_PY, s), fpz = p.F(OF_PZ, s);
    uint32_t fin = FIN_RUN, lpos = 0, lend = 0;
    double2* fab = F.ab + gslot * (size_t)F.depth; uint2* fcq = F.cq + gslot * (size_t)F.depth;
#pragma unroll 1
    for (int guard = 0; guard < N_MAX; ++guard) {
        if (qmask == 0) {
            if (sp == 0) { fin = hit ? FIN_HIT : FIN_MISS; break; }          // stack empty :276-283
            --sp;                                                            // this level is exhausted: back to its parent's frame
            const double2 ab = fab[sp]; const uint2 cq = fcq[sp];
            fa = ab.x; fb = ab.y; fchild = cq.x; qmask = cq.y;
            continue;
        }
        // The next TWO octants of the frame (pushed near->far, popped far->near) are fetched together -- content box and node record
        // of each, one round trip: more than half of the candidates are rejected, and the one behind a rejected candidate is then
        // already here.  (If the first one is entered, the second stays in the frame and is fetched again when the walk returns.)
        const int q0 = hare_fls(qmask);
        const uint32_t rest = qmask & ~(1u << q0);
        const int q1 = rest ? hare_fls(rest) : q0;
        const uint32_t child0 = fchild + (uint32_t)(q0 ^ sgn), child1 = fchild + (uint32_t)(q1 ^ sgn);
        float4 nlo[2], nhi[2]; double2 na[2], nb[2], nc[2]; uint4 m[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t child = k ? child1 : child0;
            const float4* e = T.nbox + 2 * (size_t)child;
            const double2* nq = reinterpret_cast<const double2*>(T.nodes + child);
            nlo[k] = oct_ld_f4(e); nhi[k] = oct_ld_f4(e + 1);
            na[k] = oct_ld_d2(nq); nb[k] = oct_ld_d2(nq + 1); nc[k] = oct_ld_d2(nq + 2);
            m[k] = oct_ld_u4(reinterpret_cast<const uint4*>(nq) + 3);      // first_child, list_off, list_cnt, pad
        }
        // when this is the frame's last octant the next thing needed -- unless the child opens a level of its own -- is the parent's
        // frame: fetch it in the same round trip
        const bool last_q = rest == 0 && sp > 0;
        double2 pab = make_double2(0.0, 0.0); uint2 pcq = make_uint2(0u, 0u);
        if (last_q) { pab = fab[sp - 1]; pcq = fcq[sp - 1]; }
        bool enter = false; int k = 0;
        double ca = 0, cb = 0;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
            if (!enter && (kk == 0 || rest != 0)) {
                // the ray's line missing everything listed below the child means entering it could change nothing (only a successful
                // test updates closestT or returns)
                double lo, hi;
                oct_interval_of(na[kk], nb[kk], nc[kk], ox, oy, oz, ix, iy, iz, lo, hi);
                ca = fmax(lo, fa); cb = fmin(hi, fb);
                enter = !cull_box(nlo[kk], nhi[kk], fpx, fpy, fpz, fix, fiy, fiz) &&
                        !(hi < lo || hi < 0 || lo > fb || hi < fa) &&         // push-time filter :268
                        !(cb < ca || cb < 0) && !(hit && closest <= ca) &&    // pop-time prunes :207-211
                        !((int)m[kk].x < 0 && m[kk].z == 0);                  // (an empty leaf changes nothing)
                k = kk;
                qmask = kk ? (rest & ~(1u << q1)) : rest;                     // this octant is consumed, entered or not
            }
        }
        const uint4 mm = k ? m[1] : m[0];
        if (enter && (int)mm.x >= 0) {
            // an internal node opens a level; the current top goes to the spill area -- unless it is exhausted: nothing would ever
            // be read from it again
            c.cell();
            if (qmask != 0) { fab[sp] = make_double2(fa, fb); fcq[sp] = make_uint2(fchild, qmask); ++sp; }
            fchild = mm.x; qmask = oct_perm_mask(mm.w, sgn); fa = ca; fb = cb;
            if (T.regular) qmask &= oct_child_filter(k ? na[1] : na[0], k ? nb[1] : nb[0], k ? nc[1] : nc[0], ox, oy, oz, ix, iy, iz, fa, fb, sgn);
            continue;
        }
        if (last_q) { --sp; fa = pab.x; fb = pab.y; fchild = pcq.x; qmask = pcq.y; }     // (rest == 0: the frame is exhausted either way)
        if (enter) {                                                         // a leaf with a list
            c.cell();
            lpos = mm.y; lend = mm.y + mm.z;
            p.U(OU_CIDX, s) = mm.w; p.D(OD_CA, s) = ca;
            break;
        }
    }
    fl = (fl & ~(OFL_SP_MASK | OFL_FIN_MASK)) | ((uint32_t)sp << OFL_SP_SHIFT) | (fin << OFL_FIN_SHIFT);
    p.U(OU_FLAGS, s) = fl;
    p.D(OD_FA, s) = fa; p.D(OD_FB, s) = fb;
    p.U(OU_FCHILD, s) = fchild; p.U(OU_MASKS, s) = qmask;
    p.U(OU_LPOS, s) = lpos; p.U(OU_LEND, s) = lend;
    return oct_tag(fl, qmask, lpos, lend);
}

// ---- G: the next group of (up to) 64 leaf-list entries: its box, then its (up to) eight chunk boxes
template <bool COUNT, int SLOTS>
HD uint32_t octw_group(const OctDev& T, const OctPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    static_assert(HARE_OCT_CHUNK == 8, "eight entries per chunk, eight chunks per group");
    uint32_t lpos = p.U(OU_LPOS, s);
    const uint32_t lend = p.U(OU_LEND, s);
    uint32_t cidx = p.U(OU_CIDX, s);
    const float fdx = (float)p.D(OD_DX, s), fdy = (float)p.D(OD_DY, s), fdz = (float)p.D(OD_DZ, s);
    const float fix = cull_rcp(fdx), fiy = cull_rcp(fdy), fiz = cull_rcp(fdz);
    const float fpx = p.F(OF_PX, s), fpy = p.F(OF_PY, s), fpz = p.F(OF_PZ, s);
    const uint32_t left = lend - lpos, nch = (left + 7u) / 8u < 8u ? (left + 7u) / 8u : 8u;
    uint32_t em = 0;
    // the group box and its (up to) eight chunk boxes are fetched together: one round trip (the registers are there: 16 warps per SM)
    const float4* ge = T.gbox + 2 * (size_t)(cidx >> 3);
    const float4 glo = oct_ld_f4(ge), ghi = oct_ld_f4(ge + 1);
    float4 lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4* e = T.cbox + 2 * (size_t)(cidx + (j < (int)nch ? j : 0));
        lo[j] = oct_ld_f4(e); hi[j] = oct_ld_f4(e + 1);
    }
    if (!cull_box(glo, ghi, fpx, fpy, fpz, fix, fiy, fiz)) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            em |= (j < (int)nch && !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz)) ? (1u << j) : 0u;
    }
    const uint32_t adv = left < 64u ? left : 64u;
    p.U(OU_CPOS, s) = lpos; p.U(OU_CIDX, s) = cidx + nch;
    lpos += adv;
    if (COUNT) c.entries += adv;
    p.U(OU_LPOS, s) = lpos;
    const uint32_t masks = (p.U(OU_MASKS, s) & 0xffu) | (em << 8);
    p.U(OU_MASKS, s) = masks;
    return em ? (uint32_t)OP_C : (lpos < lend ? (uint32_t)OP_G : (uint32_t)OP_N);
}

// ---- C: the (up to) eight entries of the lowest surviving chunk
template <bool COUNT, int SLOTS>
HD uint32_t octw_cull(const OctDev& T, const OctPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t masks = p.U(OU_MASKS, s);
    uint32_t emask = (masks >> 8) & 0xffu;
    const uint32_t lend = p.U(OU_LEND, s);
    const int kc = hare_ffs(emask) - 1;
    emask &= emask - 1u;
    const uint32_t base = p.U(OU_CPOS, s) + (uint32_t)kc * 8u;
    const uint32_t n = (lend - base) < 8u ? (lend - base) : 8u;
    const float fdx = (float)p.D(OD_DX, s), fdy = (float)p.D(OD_DY, s), fdz = (float)p.D(OD_DZ, s);
    const float fix = cull_rcp(fdx), fiy = cull_rcp(fdy), fiz = cull_rcp(fdz);
    const float fpx = p.F(OF_PX, s), fpy = p.F(OF_PY, s), fpz = p.F(OF_PZ, s);
    const int or1 = (int)p.U(OU_OR1, s), or2 = (int)p.U(OU_OR2, s), pid = (int)p.U(OU_PID, s);
    const uint32_t last = p.U(OU_LAST, s);
    uint32_t bm = 0, first_id = 0;
    // poly_origin skip (:218); a polygon already tested for this ray (it sits in several leaves) cannot change anything:
    // its t is not below closestT any more, so neither the update nor the early return fires
    uint32_t ids[8]; float4 lo[8], hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ids[j] = oct_ld_u32(T.lists + base + (j < (int)n ? j : 0));                                   // round trip 1: the ids
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float4* e = T.pbox + 2 * (size_t)ids[j]; lo[j] = oct_ld_f4(e); hi[j] = oct_ld_f4(e + 1); }   // round trip 2: their boxes
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t i = ids[j];
        const bool keep = (j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                          !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz);
        first_id = (keep && bm == 0) ? i : first_id;
        bm |= keep ? (1u << j) : 0u;
    }
    // the lowest survivor's id rides in the slot, so that the test phase fetches its record without re-reading the list
    masks = (masks & 0xffu) | (emask << 8) | (bm << 16) | ((uint32_t)kc << OM_KC_SHIFT) | (bm ? (uint32_t)OM_PEND : 0u);
    p.U(OU_MASKS, s) = masks;
    if (bm) p.U(OU_LAST, s) = first_id;
    return oct_tag(0, masks, p.U(OU_LPOS, s), lend);
}

// ---- T: one exact FP64 test (slow path: u, v) of the lowest surviving entry
template <bool CHAIN, bool COUNT, int SLOTS>
HD uint32_t octw_test(const OctDev& T, const PolyRec* __restrict__ polys, const OctPool<SLOTS>& p, int s, int order, const WalkOut& out, CntT<COUNT>& c) {
    uint32_t masks = p.U(OU_MASKS, s);
    uint32_t bmask = (masks >> 16) & 0xffu;
    const int k = hare_ffs(bmask) - 1;                // lowest survivor first: stored list order
    bmask &= bmask - 1u;
    const uint32_t kc = (masks & OM_KC_MASK) >> OM_KC_SHIFT;
    const uint32_t pend = (masks & OM_PEND) ? p.U(OU_LAST, s) : hare_ldg(T.lists + p.U(OU_CPOS, s) + kc * 8u + (uint32_t)k);
    masks = (masks & (0xffffu | OM_KC_MASK)) | (bmask << 16);
    c.test();
    const Ray3 R = { p.D(OD_OX, s), p.D(OD_OY, s), p.D(OD_OZ, s), p.D(OD_DX, s), p.D(OD_DY, s), p.D(OD_DZ, s) };
    double P[16], t = 0, u = 0, v = 0;
    load_poly(polys, pend, P);
    // Polygon.Ray_Side picks the winding (Hare_Geometry_Polygons.cs:601-606, 662-688, 731-782):
    //   side ? (P0,P1,P2) then (P2,P3,P0) : (P2,P1,P0) then (P0,P3,P2)
    const bool side = !(dot3(R.dx, R.dy, R.dz, P[12], P[13], P[14]) < 0);
    const double ax = side ? P[0] : P[6], ay = side ? P[1] : P[7], az = side ? P[2] : P[8];
    const double cx = side ? P[6] : P[0], cy = side ? P[7] : P[1], cz = side ? P[8] : P[2];
    bool h = ray_x_tri_slow1(R, ax, ay, az, P[3], P[4], P[5], cx, cy, cz, t, u, v);
    if (!h && P[15] == 4.0) h = ray_x_tri_slow1(R, cx, cy, cz, P[9], P[10], P[11], ax, ay, az, t, u, v);
    p.U(OU_LAST, s) = pend;
    uint32_t fl = p.U(OU_FLAGS, s);
    if (h && t > 0.0000000001 && t < p.D(OD_CLOSEST, s)) {
        p.D(OD_CLOSEST, s) = t; p.U(OU_PID, s) = pend;
        if (out.uv) { const long long ray = (long long)p.U(OU_RAY, s); out.uv[2 * ray] = u; out.uv[2 * ray + 1] = v; }
        if (CHAIN && out.ev_uv) chain_row_uv(out, (long long)p.U(OU_RAY, s), order, p.U(OU_FLAGS, s) >> OFL_BOUNCE_SHIFT, u, v);
        fl |= OFL_HIT;
        if (t <= p.D(OD_CA, s)) fl |= FIN_HIT << OFL_FIN_SHIFT;            // early return :233-237 (CA = this leaf's nodeTmin)
        p.U(OU_FLAGS, s) = fl;
    }
    p.U(OU_MASKS, s) = masks;
    return oct_tag(fl, masks, p.U(OU_LPOS, s), p.U(OU_LEND, s));
}

#if defined(__CUDACC__)

#ifndef HARE_OCTW_WARPS
#define HARE_OCTW_WARPS 16   /* a multiple of 4 (one warp set per SM sub-partition: 17 / 18 warps leave one scheduler with 5 warps and 96 registers: 608 / 643 vs 727 Mrays/s); 16 x 10.6 KB pools = 170 KB: the 196 KB shared-memory carve-out, ~60 KB of L1 left for the tree's upper levels (19 warps / 28 KB L1: 461 vs 595 at the time) */
#endif

template <bool CHAIN, bool COUNT, int SLOTS, int N_MAX>
__global__ void __launch_bounds__(HARE_OCTW_WARPS * 32, 1)
oct_wave_kernel(const OctDev T, const OctFrames F, const PolyRec* __restrict__ polys,
                const double* __restrict__ o, const double* __restrict__ d,
                const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a,
                long long N, int order, const uint32_t* __restrict__ perm /* ray order of ray_bin.cuh, or null */, const RayFeedArgs feed, const WalkOut out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    OctPool<SLOTS> p;
    p.bind(s_raw + (size_t)warp * OctPool<SLOTS>::STRIDE);
    constexpr int GROUPS = (SLOTS + 31) / 32;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const int s = k * 32 + lane;
        if (s < SLOTS) { p.U(OU_FLAGS, s) = OFL_NORAY; p.U(OU_LPOS, s) = 0; p.U(OU_LEND, s) = 0; p.U(OU_MASKS, s) = 0; p.tag[s] = (uint8_t)OP_SF; }
    }
    __syncwarp();

    CntT<COUNT> c;
    unsigned int shots = 0;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    const size_t gslot0 = (size_t)gw * SLOTS;
    RayFeed f = { gw * feed.block, 0, 0 };   // see RayFeed in vg_wave.cuh
    if (lane == 0) f.b1 = feed_claim(feed);
    const unsigned lt = (1u << lane) - 1u;

    while (true) {
        // 1. count the slots per phase: REDUX over byte-packed per-lane counts (SF, N, G, C) and a second one for T
        uint32_t tg[GROUPS], packed = 0, nt_ = 0;
#pragma unroll
        for (int k = 0; k < GROUPS; ++k) {
            const int s = k * 32 + lane;
            tg[k] = (s < SLOTS) ? p.tag[s] : (uint32_t)OP_DONE;
            packed += (tg[k] < (uint32_t)OP_T) ? (1u << (8 * tg[k])) : 0u;
            nt_ += (tg[k] == (uint32_t)OP_T) ? 1u : 0u;
        }
        packed = __reduce_add_sync(0xffffffffu, packed);
        nt_ = __reduce_add_sync(0xffffffffu, nt_);
        const int n[OP_COUNT] = { (int)(packed & 255u), (int)((packed >> 8) & 255u), (int)((packed >> 16) & 255u), (int)(packed >> 24), (int)nt_ };
        // 2. pick one
        const int ph = oct_pick(n);
        if (ph < 0) break;
        // 3. compact up to 32 of its slots, one per lane
        int base = 0;
#pragma unroll
        for (int k = 0; k < GROUPS; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, tg[k] == (uint32_t)ph);
            const int r = base + __popc(m & lt);
            if (tg[k] == (uint32_t)ph && r < 32) p.sel[r] = (uint8_t)(k * 32 + lane);
            base += __popc(m);
        }
        __syncwarp();
        const int cnt = base < 32 ? base : 32;
        const bool act = lane < cnt;
        const int s = act ? (int)p.sel[lane] : 0;
        // 4. run it
        uint32_t nt = OP_DONE;
        if (ph == OP_T) {
            if (act) nt = octw_test<CHAIN, COUNT, SLOTS>(T, polys, p, s, order, out, c);
        } else if (ph == OP_C) {
            if (act) nt = octw_cull<COUNT, SLOTS>(T, p, s, c);
        } else if (ph == OP_G) {
            if (act) nt = octw_group<COUNT, SLOTS>(T, p, s, c);
        } else if (ph == OP_N) {
            if (act) nt = octw_node<COUNT, SLOTS, N_MAX>(T, F, gslot0 + (size_t)s, p, s, c);
        } else {
            if (act) octw_finish<CHAIN, COUNT, SLOTS>(polys, p, s, order, out, shots, c);
            const bool noray = act && (p.U(OU_FLAGS, s) & OFL_NORAY);
            const unsigned want = __ballot_sync(0xffffffffu, noray);
            bool ready = act;
            const long long b1 = __shfl_sync(0xffffffffu, f.b1, 0);
            if (noray) {
                const long long ray = feed_ray(f, feed, b1, __popc(want & lt));
                if (ray < N) octw_fetch<SLOTS>(p, s, perm ? (long long)__ldg(perm + ray) : ray, o, d, o1a, o2a);
                else ready = false;
            }
            if (feed_advance(f, feed, __popc(want), b1) && lane == 0) f.b1 = feed_claim(feed);
            if (ready) nt = octw_setup<COUNT, SLOTS>(T, p, s, c);
        }
        if (act) p.tag[s] = (uint8_t)nt;
        __syncwarp();
    }
    if (CHAIN) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) shots += __shfl_xor_sync(0xffffffffu, shots, off);
        if (lane == 0 && shots) atomicAdd(out.total_shots, (unsigned long long)shots);
    }
    flush_counters<COUNT>(c, out.counters);
}

#endif  // __CUDACC__

}  // namespace hare
