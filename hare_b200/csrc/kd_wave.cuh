// kd_wave.cuh -- K4 (and K5 on a KDTree): Hare's KDTree.Shoot (KDTree.cs:198-361) under the per-warp WAVEFRONT scheduler of
// vg_wave.cuh / oct_wave.cuh: ray state in shared-memory pools (133 bytes per slot), one converged phase per trip.
//
// The reference pushes both children of every node (:355-356): it visits every leaf and keeps the strict minimum of t, i.e.
// the global closest hit.  Here the walk goes near child first and drops a subtree when the ray's parameter interval inside
// the node's box (content-tightened, inflated by HARE_KD_PAD) lies wholly beyond the current closest hit or behind the origin
// (kd_box_reachable).  t and hit/miss are the reference's; among polygons hit at exactly equal t the reference keeps the first
// of its exhaustive DFS order, which kd_dfs_before() reconstructs from the reference's own (untightened) node boxes, so
// Poly_id, u and v match as well.
//
// The walk reads 4-wide records (KdWide, shoot.cuh): a node together with its (up to four) grandchildren, FP32 padded boxes, one
// 128-byte fetch per two levels of the reference's binary tree; the four entries are ordered by entry parameter in registers (a
// 5-comparator network).  The stack of pending entries lives in a per-slot scratch area in global memory (16-byte entries: node or
// leaf list, entry parameter; 3 * (depth / 2 + 2) + 4 per slot, L2 resident): a push is a store, an entry a closer hit has overtaken
// is dropped at pop time from its parameter alone; the record being descended into stays in the slot.
//
// Phases:  SF finish / fetch / set-up;  N descend until a reachable leaf with a non-empty list (<= N_MAX records per execution);
//          C cull the next (up to) eight leaf entries (poly_origin / duplicate skip, padded-box reject);  T one exact slow-path
//          (u, v) Moller-Trumbore test, strict t < closestT.
#pragma once
#include "shoot.cuh"
#include "vg_wave.cuh"
#include "oct_wave.cuh"   // hare_ffs

namespace hare {

enum : uint32_t { KP_SF = 0, KP_N = 1, KP_C = 2, KP_T = 3, KP_DONE = 4, KP_COUNT = 4 };
// slot flags: NORAY | fin(2) | hit | blind | bmask(8) << 8 | bounce(16) << 16
enum : uint32_t { KFL_NORAY = 1u, KFL_FIN_SHIFT = 1, KFL_FIN_MASK = 3u << 1, KFL_HIT = 8u, KFL_BLIND = 16u, KFL_BMASK_SHIFT = 8, KFL_BMASK_MASK = 255u << 8, KFL_BOUNCE_SHIFT = 16 };
enum { KD_OX, KD_OY, KD_OZ, KD_DX, KD_DY, KD_DZ, KD_CLOSEST, KD_EU, KD_EV, KD_TE /* ray parameter of the FP32 frame point */, KD_COUNT };
enum { KU_FLAGS, KU_RAY, KU_PID, KU_OR1, KU_OR2, KU_LAST, KU_LPOS, KU_LEND, KU_CUR, KU_SP, KU_COUNT };
enum { KF_PX, KF_PY, KF_PZ, KF_COUNT };
#define HARE_KD_NONE 0xffffffffu

template <int SLOTS>
struct KdPool {
    double* dbl; uint32_t* u32; float* f32; uint8_t* tag; uint8_t* sel;
    static_assert(SLOTS <= 255, "phase counts are packed into bytes");
    static constexpr size_t BYTES = (size_t)SLOTS * (KD_COUNT * 8 + KU_COUNT * 4 + KF_COUNT * 4 + 1) + 32;
    static constexpr size_t STRIDE = (BYTES + 15) & ~(size_t)15;
    HD void bind(unsigned char* base) {
        dbl = reinterpret_cast<double*>(base);
        u32 = reinterpret_cast<uint32_t*>(base + (size_t)SLOTS * KD_COUNT * 8);
        f32 = reinterpret_cast<float*>(base + (size_t)SLOTS * (KD_COUNT * 8 + KU_COUNT * 4));
        tag = base + (size_t)SLOTS * (KD_COUNT * 8 + KU_COUNT * 4 + KF_COUNT * 4);
        sel = tag + SLOTS;
    }
    HD double& D(int f, int s) const { return dbl[f * SLOTS + s]; }
    HD uint32_t& U(int f, int s) const { return u32[f * SLOTS + s]; }
    HD float& F(int f, int s) const { return f32[f * SLOTS + s]; }
};

struct KdStacks { uint4* st; int depth; };   // per slot: `depth` pending entries (x, y, entry parameter as float bits, -): internal y = ~0, x = node; leaf x = list offset, y = count

HD uint32_t kd_tag(uint32_t fl, uint32_t lpos, uint32_t lend) {
    if (fl & KFL_FIN_MASK) return KP_SF;
    if (fl & KFL_BMASK_MASK) return KP_T;
    return lpos < lend ? KP_C : KP_N;
}

HD int kd_pick(const int n[KP_COUNT]) {
    int best = KP_T, bn = n[KP_T];
    if (n[KP_C] > bn) { best = KP_C; bn = n[KP_C]; }
    if (n[KP_N] > bn) { best = KP_N; bn = n[KP_N]; }
    if (n[KP_SF] > bn) { best = KP_SF; bn = n[KP_SF]; }
    return bn > 0 ? best : -1;
}

HD int kd_lo32(double x) { unsigned long long b; memcpy(&b, &x, 8); return (int)(uint32_t)(b & 0xffffffffull); }
HD int kd_hi32(double x) { unsigned long long b; memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }

// conservative prune: can a polygon of this node be hit at 0 <= t <= closest inside the node's (tightened, padded) box?
HD bool kd_box_reachable_hd(const double2 a, const double2 b, const double2 cc, double ox, double oy, double oz, double dx, double dy, double dz,
                            double ix, double iy, double iz, double closest, double& lo) {
    double hi = closest;
    lo = 0.0;
    const double mn[3] = { a.x - HARE_KD_PAD, a.y - HARE_KD_PAD, b.x - HARE_KD_PAD };
    const double mx[3] = { b.y + HARE_KD_PAD, cc.x + HARE_KD_PAD, cc.y + HARE_KD_PAD };
    const double o[3] = { ox, oy, oz }, d[3] = { dx, dy, dz }, inv[3] = { ix, iy, iz };
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.0) {
            if (o[k] < mn[k] || o[k] > mx[k]) return false;
        } else {
            double t1 = (mn[k] - o[k]) * inv[k], t2 = (mx[k] - o[k]) * inv[k];
            if (t1 > t2) { double t = t1; t1 = t2; t2 = t; }
            if (t1 > lo) lo = t1;
            if (t2 < hi) hi = t2;
        }
    }
    return !(lo > hi);
}

// Exact-t tie between two different polygons: the reference keeps the one its exhaustive DFS meets first (strict t < closestT,
// mailbox = first occurrence only).  That order is reconstructed analytically: walk down from the root choosing first/second by
// the reference's rule (KDTree.cs:249-353) evaluated on the REFERENCE's node boxes (KdDev::ref_box, never the tightened device
// boxes); a polygon belongs to the Left subtree iff one of its vertices is <= split on the node's axis, to the Right iff one is
// > split (:123-133).  Where the two polygons part ways the one in `first` wins; in a common leaf the earlier list entry wins.
// Rare (rays through shared edges or vertices of a structured mesh).
HD bool kd_dfs_before_hd(const KdDev& T, const PolyRec* __restrict__ polys, const Ray3& R, uint32_t pa, uint32_t pb) {
    const double* A = polys[pa].v; const double* B = polys[pb].v;
    const int na = (A[15] == 4.0) ? 4 : 3, nb = (B[15] == 4.0) ? 4 : 3;
    uint32_t ni = 0;
    for (int depth = 0; depth < HARE_KD_MAXSTACK; ++depth) {
        const double2 dd = hare_ldg(reinterpret_cast<const double2*>(T.nodes + ni) + 3);
        const int left = kd_lo32(dd.y), axis = kd_hi32(dd.y);
        if (left < 0) {   // common leaf: stored list order
            const uint32_t off = (uint32_t)kd_lo32(dd.x), cnt = (uint32_t)kd_hi32(dd.x);
            for (uint32_t k = 0; k < cnt; ++k) { const uint32_t i = hare_ldg(T.lists + off + k); if (i == pa) return true; if (i == pb) return false; }
            return pa < pb;
        }
        const double split = dd.x;
        bool aL = false, aR = false, bL = false, bR = false;
        for (int k = 0; k < na; ++k) { const double c = A[3 * k + axis]; aL |= (c <= split); aR |= (c > split); }
        for (int k = 0; k < nb; ++k) { const double c = B[3 * k + axis]; bL |= (c <= split); bR |= (c > split); }
        // first / second exactly as the reference computes them (:249-353), on currentNode.Min / .Max
        const double* rb = T.ref_box + 6 * (size_t)ni;
        const double mn[3] = { rb[0], rb[1], rb[2] }, mx[3] = { rb[3], rb[4], rb[5] };
        const double o[3] = { R.x, R.y, R.z }, d[3] = { R.dx, R.dy, R.dz };
        const int b1 = (axis == 0) ? 1 : 0, b2 = (axis == 2) ? 1 : 2;
        const double side = o[axis] - split;
        const double tSplit = -side / d[axis];
        const double s1 = o[b1] + tSplit * d[b1], s2 = o[b2] + tSplit * d[b2];
        const bool inside = (s1 <= mx[b1] && s1 >= mn[b1] && s2 <= mx[b2] && s2 >= mn[b2]);
        const bool right_first = inside ? (side >= 0) : !(side >= 0);
        const bool aF = right_first ? aR : aL, bF = right_first ? bR : bL;   // membership in the subtree visited first
        if (aF != bF) return aF;
        ni = (uint32_t)(aF ? (right_first ? left + 1 : left) : (right_first ? left : left + 1));
    }
    return pa < pb;
}
#if defined(__CUDA_ARCH__)
__device__ __noinline__ bool kd_dfs_before_w(const KdDev T, const PolyRec* __restrict__ polys, const Ray3 R, uint32_t pa, uint32_t pb) { return kd_dfs_before_hd(T, polys, R, pa, pb); }
#else
inline bool kd_dfs_before_w(const KdDev& T, const PolyRec* polys, const Ray3& R, uint32_t pa, uint32_t pb) { return kd_dfs_before_hd(T, polys, R, pa, pb); }
#endif

template <bool CHAIN, bool COUNT, int SLOTS>
HD void kdw_finish(const PolyRec* __restrict__ polys, const KdPool<SLOTS>& p, int s, int order, const WalkOut& out, unsigned int& shots, CntT<COUNT>& c) {
    uint32_t fl = p.U(KU_FLAGS, s);
    const uint32_t fin = (fl & KFL_FIN_MASK) >> KFL_FIN_SHIFT;
    if (fin == FIN_RUN) return;
    const bool h = fin == FIN_HIT;
    const double closest = p.D(KD_CLOSEST, s);
    Ray3 R = { p.D(KD_OX, s), p.D(KD_OY, s), p.D(KD_OZ, s), p.D(KD_DX, s), p.D(KD_DY, s), p.D(KD_DZ, s) };
    const int pid = (int)p.U(KU_PID, s);
    const long long ray = (long long)p.U(KU_RAY, s);
    const double bx = R.x + R.dx * closest, by = R.y + R.dy * closest, bz = R.z + R.dz * closest;   // X_Point, Polygons.cs:749
    if (h) c.hit();
    fl &= ~(KFL_FIN_MASK | KFL_HIT | KFL_BMASK_MASK);
    if (CHAIN) {
        uint32_t bounce = fl >> KFL_BOUNCE_SHIFT;
        ++shots;
        if (out.ev_pid) out.ev_pid[ray * order + bounce] = h ? pid : -1;
        if (out.ev_t) out.ev_t[ray * order + bounce] = h ? closest : 0.0;
        chain_row_xyz(out, ray, order, bounce, h, bx, by, bz);
        chain_row_uv(out, ray, order, bounce, h ? p.D(KD_EU, s) : 0.0, h ? p.D(KD_EV, s) : 0.0);
        ++bounce;
        bool go_on = false;
        if (h) {
            const double* P = polys[pid].v;
            const double nx = hare_ldg(P + 12), ny = hare_ldg(P + 13), nz = hare_ldg(P + 14);
            const double k = 2 * ((R.dx * nx) + (R.dy * ny) + (R.dz * nz));
            R.dx = R.dx - k * nx; R.dy = R.dy - k * ny; R.dz = R.dz - k * nz;
            R.x = bx; R.y = by; R.z = bz;
            p.D(KD_OX, s) = R.x; p.D(KD_OY, s) = R.y; p.D(KD_OZ, s) = R.z;
            p.D(KD_DX, s) = R.dx; p.D(KD_DY, s) = R.dy; p.D(KD_DZ, s) = R.dz;
            p.U(KU_OR1, s) = (uint32_t)pid;
            go_on = (int)bounce < order;
        }
        fl = (fl & 0xffffu) | (bounce << KFL_BOUNCE_SHIFT);
        if (!go_on) {
            for (int q = (int)bounce; q < order; ++q) {
                if (out.ev_pid) out.ev_pid[ray * order + q] = -3;
                if (out.ev_t) out.ev_t[ray * order + q] = 0;
            }
            chain_rows_clear(out, ray, order, (int)bounce);
            if (out.fin_o) { out.fin_o[3 * ray] = R.x; out.fin_o[3 * ray + 1] = R.y; out.fin_o[3 * ray + 2] = R.z; }
            if (out.fin_d) { out.fin_d[3 * ray] = R.dx; out.fin_d[3 * ray + 1] = R.dy; out.fin_d[3 * ray + 2] = R.dz; }
            if (out.nshots) out.nshots[ray] = (int32_t)bounce;
            fl |= KFL_NORAY;
        }
    } else {
        out.pid[ray] = h ? pid : -1;
        if (out.t) out.t[ray] = h ? closest : 0.0;
        if (out.xyz) { out.xyz[3 * ray] = h ? bx : 0.0; out.xyz[3 * ray + 1] = h ? by : 0.0; out.xyz[3 * ray + 2] = h ? bz : 0.0; }
        if (out.uv) { out.uv[2 * ray] = h ? p.D(KD_EU, s) : 0.0; out.uv[2 * ray + 1] = h ? p.D(KD_EV, s) : 0.0; }
        if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }   // the KDTree never moves a ray
        fl |= KFL_NORAY;
    }
    p.U(KU_FLAGS, s) = fl;
}

template <int SLOTS>
HD void kdw_fetch(const KdPool<SLOTS>& p, int s, long long ray, const double* __restrict__ o, const double* __restrict__ d,
                  const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid) {
    p.D(KD_OX, s) = o[3 * ray]; p.D(KD_OY, s) = o[3 * ray + 1]; p.D(KD_OZ, s) = o[3 * ray + 2];
    p.D(KD_DX, s) = d[3 * ray]; p.D(KD_DY, s) = d[3 * ray + 1]; p.D(KD_DZ, s) = d[3 * ray + 2];
    p.U(KU_OR1, s) = (uint32_t)(o1a ? o1a[ray] : -1);
    p.U(KU_OR2, s) = (uint32_t)(o2a ? o2a[ray] : -1);
    p.U(KU_RAY, s) = (uint32_t)ray;
    p.U(KU_FLAGS, s) = (rid && rid[ray] == 0) ? KFL_BLIND : 0u;
}

template <bool COUNT, int SLOTS>
HD uint32_t kdw_setup(const KdDev& T, const KdPool<SLOTS>& p, int s, CntT<COUNT>&) {
    uint32_t fl = p.U(KU_FLAGS, s) & (KFL_BLIND | (0xffffu << KFL_BOUNCE_SHIFT));
    const double ox = p.D(KD_OX, s), oy = p.D(KD_OY, s), oz = p.D(KD_OZ, s), dx = p.D(KD_DX, s), dy = p.D(KD_DY, s), dz = p.D(KD_DZ, s);
    p.D(KD_CLOSEST, s) = DBL_MAX; p.D(KD_EU, s) = 0; p.D(KD_EV, s) = 0;
    p.U(KU_PID, s) = 0xffffffffu; p.U(KU_LAST, s) = 0xffffffffu;
    p.U(KU_LPOS, s) = 0; p.U(KU_LEND, s) = 0; p.U(KU_CUR, s) = 0; p.U(KU_SP, s) = 0;
    // Frame of the FP32 walk: the point where the ray enters the root's box (the exact vertex bounds; the origin itself when it starts
    // inside), formed in FP64 -- a ray shot from far outside the model must not lose the millimetres the padding allows to FP32.  A ray
    // that misses the root box hits nothing.
    const double2* q = reinterpret_cast<const double2*>(T.nodes);
    double te = 0.0;
    const bool in = kd_box_reachable_hd(hare_ldg(q), hare_ldg(q + 1), hare_ldg(q + 2), ox, oy, oz, dx, dy, dz, 1.0 / dx, 1.0 / dy, 1.0 / dz, DBL_MAX, te);
    p.D(KD_TE, s) = te;
    const float fix = cull_rcp((float)dx), fiy = cull_rcp((float)dy), fiz = cull_rcp((float)dz);
    p.F(KF_PX, s) = (float)fma(dx, te, ox) * fix; p.F(KF_PY, s) = (float)fma(dy, te, oy) * fiy; p.F(KF_PZ, s) = (float)fma(dz, te, oz) * fiz;
    // Ray_ID == 0 against a fresh mailbox: every polygon is rejected (KDTree.cs:58-66, 224-229)
    if ((fl & KFL_BLIND) || !in) fl |= FIN_MISS << KFL_FIN_SHIFT;
    uint32_t lpos = 0, lend = 0;
    {   // a tree that is a single leaf has no KdWide record: its list is scanned directly
        const uint32_t rb = hare_ldg(&T.hot[0].b);
        if ((rb & 3u) == 3u) {
            p.U(KU_CUR, s) = HARE_KD_NONE;
            if (!(fl & KFL_FIN_MASK)) { lpos = hare_ldg(&T.hot[0].a); lend = lpos + (rb >> 2); }
        }
    }
    p.U(KU_LPOS, s) = lpos; p.U(KU_LEND, s) = lend;
    p.U(KU_FLAGS, s) = fl;
    return kd_tag(fl, lpos, lend);
}

// closest - te as a float that is not below the exact difference (the prune compares a box's entry parameter against it)
HD float kd_upper_float(double x) {
    float f = (float)x;
    if ((double)f < x) f = f > 0.0f ? f * 1.0000002f : f * 0.9999998f;      // one step up (x is finite and far from the float range's ends)
    return f;
}

// ---- N: descend until a reachable leaf with a non-empty list is found, the stack runs empty, or N_MAX records were looked at.
// One KdWide record = a node's (up to four) grandchildren = two levels of the reference's binary tree per dependent fetch.
// FP32 on padded boxes: slab test of the ray's line against each entry's box in the ray's frame (as cull_box), then the ray's own
// extent: a box wholly behind the ray's start (tf < -te) or wholly beyond the closest hit (tn > closest - te) holds no polygon that
// could still improve or tie the result.  Boxes are padded by >= 1e-3 m in space while the FP32 evaluation is good to ~1e-5 m, so the
// computed interval contains every hit parameter inside the unpadded box: nothing reachable is ever skipped.  The reference's
// first/second rule (KDTree.cs:249-353) only fixes the order in which its exhaustive walk meets the leaves -- the result is the
// global minimum of t either way --, so the nearest reachable entry goes first and the others wait on the stack WITH their entry
// parameter: an entry that a closer hit has overtaken in the meantime is dropped at pop time without touching memory.
template <bool COUNT, int SLOTS, int N_MAX>
HD uint32_t kdw_node(const KdDev& T, const KdStacks& S, size_t gslot, const KdPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(KU_FLAGS, s);
    const float fdx = (float)p.D(KD_DX, s), fdy = (float)p.D(KD_DY, s), fdz = (float)p.D(KD_DZ, s);
    const float fix = cull_rcp(fdx), fiy = cull_rcp(fdy), fiz = cull_rcp(fdz);
    const float fpx = p.F(KF_PX, s), fpy = p.F(KF_PY, s), fpz = p.F(KF_PZ, s);
    const double closest = p.D(KD_CLOSEST, s), te = p.D(KD_TE, s);
    const float crel = (fl & KFL_HIT) ? kd_upper_float(closest - te) : 3.0e38f;
    const float tback = -kd_upper_float(te);                      // the ray's own start, seen from the frame point
    uint32_t cur = p.U(KU_CUR, s), sp = p.U(KU_SP, s);
    uint4* st = S.st + gslot * (size_t)S.depth;
    uint32_t fin = FIN_RUN, lpos = 0, lend = 0;
#pragma unroll 1
    for (int guard = 0; guard < N_MAX; ++guard) {
        if (cur == HARE_KD_NONE) {
            if (sp == 0) { fin = (fl & KFL_HIT) ? FIN_HIT : FIN_MISS; break; }
            const uint4 e = st[--sp];
            if (hare_u2f(e.z) > crel) continue;                   // overtaken by a closer hit since it was pushed
            if (e.y != 0xffffffffu) { lpos = e.x; lend = e.x + e.y; break; }
            cur = e.x;
        }
        const float4* q = reinterpret_cast<const float4*>(T.wide + cur);
        float4 lo[4], hi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { lo[k] = hare_ldg(q + 2 * k); hi[k] = hare_ldg(q + 2 * k + 1); }     // mnx mny mnz mxx | mxy mxz a b
        // per entry: sort key (entry parameter, or +huge when it cannot matter) and the stack code (x, y)
        float key[4]; uint32_t ex[4], ey[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float ax = fmaf(lo[k].x, fix, -fpx), bx = fmaf(lo[k].w, fix, -fpx);
            const float ay = fmaf(lo[k].y, fiy, -fpy), by = fmaf(hi[k].x, fiy, -fpy);
            const float az = fmaf(lo[k].z, fiz, -fpz), bz = fmaf(hi[k].y, fiz, -fpz);
            const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
            const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
            const uint32_t b = hare_f2u(hi[k].w);
            const bool leaf = (b & 3u) == 3u;
            const bool ok = !(tn > tf || tf < tback || tn > crel) && (b & 3u) != 2u && !(leaf && (b >> 2) == 0u);
            if (ok) c.cell();
            key[k] = ok ? fminf(tn, 1.0e38f) : 3.0e38f;
            ex[k] = leaf ? hare_f2u(hi[k].z) : (b >> 2);
            ey[k] = leaf ? (b >> 2) : 0xffffffffu;
        }
        // sorting network (5 compare-exchanges, registers only): key[0] <= key[1] <= key[2] <= key[3]
#define HARE_KD_CX(i, j) { const bool sw = key[j] < key[i]; const float tk = sw ? key[j] : key[i]; key[j] = sw ? key[i] : key[j]; key[i] = tk; \
                           const uint32_t tx = sw ? ex[j] : ex[i]; ex[j] = sw ? ex[i] : ex[j]; ex[i] = tx; \
                           const uint32_t ty = sw ? ey[j] : ey[i]; ey[j] = sw ? ey[i] : ey[j]; ey[i] = ty; }
        HARE_KD_CX(0, 1) HARE_KD_CX(2, 3) HARE_KD_CX(0, 2) HARE_KD_CX(1, 3) HARE_KD_CX(1, 2)
#undef HARE_KD_CX
        // the nearest reachable entry is taken now, the others are pushed farthest first
        cur = HARE_KD_NONE;
        if (!(key[0] < 2.0e38f)) continue;
#pragma unroll
        for (int k = 3; k >= 1; --k)
            if (key[k] < 2.0e38f && (int)sp < S.depth) st[sp++] = make_uint4(ex[k], ey[k], hare_f2u(key[k]), 0u);
        if (ey[0] != 0xffffffffu) { lpos = ex[0]; lend = ex[0] + ey[0]; break; }
        cur = ex[0];
    }
    fl |= fin << KFL_FIN_SHIFT;
    p.U(KU_FLAGS, s) = fl; p.U(KU_CUR, s) = cur; p.U(KU_SP, s) = sp;
    p.U(KU_LPOS, s) = lpos; p.U(KU_LEND, s) = lend;
    return kd_tag(fl, lpos, lend);
}

// ---- C: cull the next (up to) eight leaf entries; every list entry carries its polygon's padded box and its id (lo.w)
template <bool COUNT, int SLOTS>
HD uint32_t kdw_cull(const KdDev& T, const KdPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    const uint32_t lpos = p.U(KU_LPOS, s), lend = p.U(KU_LEND, s);
    const uint32_t n = (lend - lpos) < 8u ? (lend - lpos) : 8u;
    if (COUNT) c.entries += n;
    const float fix = cull_rcp((float)p.D(KD_DX, s)), fiy = cull_rcp((float)p.D(KD_DY, s)), fiz = cull_rcp((float)p.D(KD_DZ, s));
    const float fpx = p.F(KF_PX, s), fpy = p.F(KF_PY, s), fpz = p.F(KF_PZ, s);
    const int or1 = (int)p.U(KU_OR1, s), or2 = (int)p.U(KU_OR2, s), pid = (int)p.U(KU_PID, s);
    const uint32_t last = p.U(KU_LAST, s);
    uint32_t bm = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float4 lo[4], hi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4* e = T.lbox + 2 * (size_t)(lpos + (4 * h + j < (int)n ? 4 * h + j : 0));
            lo[j] = hare_ldg(e); hi[j] = hare_ldg(e + 1);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t i = hare_f2u(lo[j].w);   // poly_origin skip (:220); mailbox: a polygon counts once (:224-229)
            const bool keep = (4 * h + j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                              !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz);
            bm |= keep ? (1u << (4 * h + j)) : 0u;
        }
    }
    uint32_t fl = p.U(KU_FLAGS, s);
    if (bm) { fl |= bm << KFL_BMASK_SHIFT; p.U(KU_FLAGS, s) = fl; return KP_T; }   // lpos stays on the batch until T has consumed it
    p.U(KU_LPOS, s) = lpos + n;
    return lpos + n < lend ? (uint32_t)KP_C : (uint32_t)KP_N;
}

// ---- T: one exact FP64 test (slow path: u, v) of the lowest surviving entry
template <bool COUNT, int SLOTS>
HD uint32_t kdw_test(const KdDev& T, const PolyRec* __restrict__ polys, const KdPool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(KU_FLAGS, s);
    uint32_t bmask = (fl & KFL_BMASK_MASK) >> KFL_BMASK_SHIFT;
    const int k = hare_ffs(bmask) - 1;
    bmask &= bmask - 1u;
    fl = (fl & ~KFL_BMASK_MASK) | (bmask << KFL_BMASK_SHIFT);
    uint32_t lpos = p.U(KU_LPOS, s);
    const uint32_t lend = p.U(KU_LEND, s);
    const uint32_t pend = hare_ldg(T.lists + lpos + (uint32_t)k);
    if (!bmask) { lpos += (lend - lpos) < 8u ? (lend - lpos) : 8u; p.U(KU_LPOS, s) = lpos; }
    c.test();
    const Ray3 R = { p.D(KD_OX, s), p.D(KD_OY, s), p.D(KD_OZ, s), p.D(KD_DX, s), p.D(KD_DY, s), p.D(KD_DZ, s) };
    double P[16], t = 0, u = 0, v = 0;
    load_poly(polys, pend, P);
    const bool side = !(dot3(R.dx, R.dy, R.dz, P[12], P[13], P[14]) < 0);
    const double ax = side ? P[0] : P[6], ay = side ? P[1] : P[7], az = side ? P[2] : P[8];
    const double cx = side ? P[6] : P[0], cy = side ? P[7] : P[1], cz = side ? P[8] : P[2];
    bool h = ray_x_tri_slow1(R, ax, ay, az, P[3], P[4], P[5], cx, cy, cz, t, u, v);
    if (!h && P[15] == 4.0) h = ray_x_tri_slow1(R, cx, cy, cz, P[9], P[10], P[11], ax, ay, az, t, u, v);
    p.U(KU_LAST, s) = pend;
    if (h && t > 0.0000000001) {
        const double closest = p.D(KD_CLOSEST, s);
        const uint32_t pid = p.U(KU_PID, s);
        if (t < closest) {
            p.D(KD_CLOSEST, s) = t; p.D(KD_EU, s) = u; p.D(KD_EV, s) = v; p.U(KU_PID, s) = pend; fl |= KFL_HIT;
        } else if (t == closest && pend != pid && kd_dfs_before_w(T, polys, R, pend, pid)) {
            p.D(KD_EU, s) = u; p.D(KD_EV, s) = v; p.U(KU_PID, s) = pend;
        }
    }
    p.U(KU_FLAGS, s) = fl;
    return kd_tag(fl, lpos, lend);
}

#if defined(__CUDACC__)

#ifndef HARE_KDW_WARPS
#define HARE_KDW_WARPS 20
#endif

template <bool CHAIN, bool COUNT, int SLOTS, int N_MAX>
__global__ void __launch_bounds__(HARE_KDW_WARPS * 32, 1)
kd_wave_kernel(const KdDev T, const KdStacks S, const PolyRec* __restrict__ polys,
               const double* __restrict__ o, const double* __restrict__ d,
               const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid,
               long long N, int order, const uint32_t* __restrict__ perm /* ray order of ray_bin.cuh, or null */, const RayFeedArgs feed, const WalkOut out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    KdPool<SLOTS> p;
    p.bind(s_raw + (size_t)warp * KdPool<SLOTS>::STRIDE);
    constexpr int GROUPS = (SLOTS + 31) / 32;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const int s = k * 32 + lane;
        if (s < SLOTS) { p.U(KU_FLAGS, s) = KFL_NORAY; p.U(KU_LPOS, s) = 0; p.U(KU_LEND, s) = 0; p.tag[s] = (uint8_t)KP_SF; }
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
        uint32_t tg[GROUPS], packed = 0;
#pragma unroll
        for (int k = 0; k < GROUPS; ++k) {
            const int s = k * 32 + lane;
            tg[k] = (s < SLOTS) ? p.tag[s] : (uint32_t)KP_DONE;
            packed += (tg[k] < (uint32_t)KP_COUNT) ? (1u << (8 * tg[k])) : 0u;
        }
        packed = __reduce_add_sync(0xffffffffu, packed);
        const int n[KP_COUNT] = { (int)(packed & 255u), (int)((packed >> 8) & 255u), (int)((packed >> 16) & 255u), (int)(packed >> 24) };
        const int ph = kd_pick(n);
        if (ph < 0) break;
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
        uint32_t nt = KP_DONE;
        if (ph == KP_T) {
            if (act) nt = kdw_test<COUNT, SLOTS>(T, polys, p, s, c);
        } else if (ph == KP_C) {
            if (act) nt = kdw_cull<COUNT, SLOTS>(T, p, s, c);
        } else if (ph == KP_N) {
            if (act) nt = kdw_node<COUNT, SLOTS, N_MAX>(T, S, gslot0 + (size_t)s, p, s, c);
        } else {
            if (act) kdw_finish<CHAIN, COUNT, SLOTS>(polys, p, s, order, out, shots, c);
            const bool noray = act && (p.U(KU_FLAGS, s) & KFL_NORAY);
            const unsigned want = __ballot_sync(0xffffffffu, noray);
            bool ready = act;
            const long long b1 = __shfl_sync(0xffffffffu, f.b1, 0);
            if (noray) {
                const long long ray = feed_ray(f, feed, b1, __popc(want & lt));
                if (ray < N) kdw_fetch<SLOTS>(p, s, perm ? (long long)__ldg(perm + ray) : ray, o, d, o1a, o2a, rid);
                else ready = false;
            }
            if (feed_advance(f, feed, __popc(want), b1) && lane == 0) f.b1 = feed_claim(feed);
            if (ready) nt = kdw_setup<COUNT, SLOTS>(T, p, s, c);
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
