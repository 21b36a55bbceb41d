// kd_walk.cuh -- K4 (and K5 on a KDTree): persistent, warp-synchronous phased traversal of Hare's
// KDTree (KDTree.cs:198-361), built like vg_walk.cuh / oct_walk.cuh.
//
// The reference pushes both children of every node (:355-356): it visits every leaf and keeps the strict
// minimum of t, i.e. the global closest hit.  Here the walk goes near child first and drops a subtree when
// the ray's parameter interval inside the node's box (inflated by HARE_KD_PAD) lies wholly beyond the
// current closest hit or behind the origin (kd_box_reachable).  t and hit/miss are the reference's; among
// polygons hit at exactly equal t the reference keeps the first of its exhaustive DFS order, which
// kd_dfs_before() reconstructs, so Poly_id, u and v match as well.
//
// Phases per trip round the main loop (all lanes of the warp together):
//   S  finish / fetch / set-up (batched);   N  pop nodes until a reachable leaf is found;
//   C  cull the next (up to) eight leaf entries: poly_origin and duplicate skip, FP32 sphere reject;
//   T  one exact slow-path (u, v) Moller-Trumbore test; strict t < closestT.
#pragma once
#include "shoot.cuh"
#include "vg_walk.cuh"   // WalkOut, ST_* states

namespace hare {

#ifndef HARE_KD_THREADS
#define HARE_KD_THREADS 640
#endif
#define HARE_KD_CB 8
// 1: leaf entries are culled on padded FP32 bounding boxes (cull_box; KdDev::lbox); 0: on spheres
// 1: per-entry boxes come from the per-POLYGON table (KdDev::pbox, by id) instead of a per-entry copy (see oct_walk.cuh)
#ifndef HARE_KD_ENTRY_PBOX
#define HARE_KD_ENTRY_PBOX 0
#endif
#ifndef HARE_KD_BOX
#define HARE_KD_BOX 1
#endif

// Exact-t tie between two different polygons: the reference keeps the one its exhaustive DFS meets first
// (strict t < closestT, mailbox = first occurrence only).  That order is reconstructed analytically: walk down
// from the root choosing first/second by the reference's rule (KDTree.cs:249-353); a polygon belongs to the Left
// subtree iff one of its vertices is <= split on the node's axis, to the Right iff one is > split (:123-133).
// Where the two polygons part ways the one in `first` wins; in a common leaf the earlier list entry wins.
// Out of line and rare (axis-aligned rays through shared edges or vertices of a structured mesh).
__device__ __noinline__ bool kd_dfs_before(const KdDev T, const PolyRec* __restrict__ polys, const Ray3 R, uint32_t pa, uint32_t pb) {
    const double* A = polys[pa].v; const double* B = polys[pb].v;
    const int na = (A[15] == 4.0) ? 4 : 3, nb = (B[15] == 4.0) ? 4 : 3;
    int ni = 0;
    for (int depth = 0; depth < HARE_KD_MAXSTACK; ++depth) {
        const double2* q = reinterpret_cast<const double2*>(T.nodes + ni);
        const double2 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2), dd = __ldg(q + 3);
        const int left = __double2loint(dd.y), axis = __double2hiint(dd.y);
        if (left < 0) {   // common leaf: stored list order
            const uint32_t off = (uint32_t)__double2loint(dd.x), cnt = (uint32_t)__double2hiint(dd.x);
            for (uint32_t k = 0; k < cnt; ++k) { const uint32_t i = __ldg(T.lists + off + k); if (i == pa) return true; if (i == pb) return false; }
            return pa < pb;
        }
        const double split = dd.x;
        bool aL = false, aR = false, bL = false, bR = false;
        for (int k = 0; k < na; ++k) { const double c = A[3 * k + axis]; aL |= (c <= split); aR |= (c > split); }
        for (int k = 0; k < nb; ++k) { const double c = B[3 * k + axis]; bL |= (c <= split); bR |= (c > split); }
        // first / second exactly as the reference computes them
        const double mn[3] = { a.x, a.y, b.x }, mx[3] = { b.y, cc.x, cc.y };
        const double o[3] = { R.x, R.y, R.z }, d[3] = { R.dx, R.dy, R.dz };
        const int b1 = (axis == 0) ? 1 : 0, b2 = (axis == 2) ? 1 : 2;
        const double side = o[axis] - split;
        const double tSplit = -side / d[axis];
        const double s1 = o[b1] + tSplit * d[b1], s2 = o[b2] + tSplit * d[b2];
        const bool inside = (s1 <= mx[b1] && s1 >= mn[b1] && s2 <= mx[b2] && s2 >= mn[b2]);
        const bool right_first = inside ? (side >= 0) : !(side >= 0);
        const bool aF = right_first ? aR : aL, bF = right_first ? bR : bL;   // membership in the subtree visited first
        if (aF != bF) return aF;
        ni = aF ? (right_first ? left + 1 : left) : (right_first ? left : left + 1);
    }
    return pa < pb;
}

template <bool CHAIN, bool COUNT, int S_BATCH, int N_MAX, int N_BATCH, int T_BATCH>
__global__ void __launch_bounds__(HARE_KD_THREADS, 1)
kd_walk_kernel(const KdDev T, const PolyRec* __restrict__ polys,
               const double* __restrict__ o, const double* __restrict__ d,
               const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid,
               long long N, int order, const WalkOut out) {
    CntT<COUNT> c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long next = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long ray = -1;

    Ray3 R = { 0, 0, 0, 0, 0, 0 };
    double inv[3] = { 0, 0, 0 };                     // reciprocals for the conservative prune only
    double closest = DBL_MAX, eu = 0, ev = 0;
    float fdx = 0, fdy = 0, fdz = 0, fdd = 0, fpx = 0, fpy = 0, fpz = 0;
#if HARE_KD_BOX
    float fix = 0, fiy = 0, fiz = 0;   // FP32 reciprocal direction (cull_box); with it fpx.. hold p/d instead of p
#endif
    int stack[HARE_KD_MAXSTACK];
    int sp = 0;
    int pid = -1, or1 = -1, or2 = -1, bounce = 0;
    uint32_t lpos = 0, lend = 0, last = 0xffffffffu;
    uint32_t bid[HARE_KD_CB] = { 0 }, bmask = 0;
    bool hit = false, blind = false;
    int state = ST_NEED_RAY;
    int fin = 2;   // 2 = running; 1 hit, 0 miss
    unsigned int shots = 0;

    while (true) {
        const unsigned want = __ballot_sync(0xffffffffu, (state == ST_NEED_RAY || state == ST_NEED_SETUP) || (state == ST_WALK && fin != 2));
        const unsigned busy = __ballot_sync(0xffffffffu, state == ST_WALK && fin == 2);
        if (want == 0 && busy == 0) break;
        const bool needN = state == ST_WALK && fin == 2 && bmask == 0 && lpos >= lend;
        const bool needC = state == ST_WALK && fin == 2 && bmask == 0 && lpos < lend;
        const bool needT = state == ST_WALK && fin == 2 && bmask != 0;
        const int nN = __popc(__ballot_sync(0xffffffffu, needN)), nC = __popc(__ballot_sync(0xffffffffu, needC));
        const int nT = __popc(__ballot_sync(0xffffffffu, needT));
        const bool doT = nT > 0 && (nT >= T_BATCH || nC == 0);
        const bool doN = nN > 0 && (nN >= N_BATCH || (nC == 0 && !doT));
        // ------------------------------------------------------------------ S phase
        if (want && (__popc(want) >= S_BATCH || busy == 0 || (nC == 0 && !doT && !doN))) {
            if (state == ST_WALK && fin != 2) {
                const bool h = fin == 1;
                const double bx = R.x + R.dx * closest, by = R.y + R.dy * closest, bz = R.z + R.dz * closest;   // X_Point, Polygons.cs:749
                if (h) c.hit();
                state = ST_NEED_RAY;
                if (CHAIN) {
                    ++shots;
                    if (out.ev_pid) out.ev_pid[ray * order + bounce] = h ? pid : -1;
                    if (out.ev_t) out.ev_t[ray * order + bounce] = h ? closest : 0.0;
                    ++bounce;
                    if (h) {
                        const double* P = polys[pid].v;
                        const double nx = __ldg(P + 12), ny = __ldg(P + 13), nz = __ldg(P + 14);
                        const double k = 2 * ((R.dx * nx) + (R.dy * ny) + (R.dz * nz));
                        R.dx = R.dx - k * nx; R.dy = R.dy - k * ny; R.dz = R.dz - k * nz;
                        R.x = bx; R.y = by; R.z = bz;
                        or1 = pid;
                        if (bounce < order) state = ST_NEED_SETUP;
                    }
                    if (state == ST_NEED_RAY) {
                        for (int q = bounce; q < order; ++q) {
                            if (out.ev_pid) out.ev_pid[ray * order + q] = -3;
                            if (out.ev_t) out.ev_t[ray * order + q] = 0;
                        }
                        if (out.fin_o) { out.fin_o[3 * ray] = R.x; out.fin_o[3 * ray + 1] = R.y; out.fin_o[3 * ray + 2] = R.z; }
                        if (out.fin_d) { out.fin_d[3 * ray] = R.dx; out.fin_d[3 * ray + 1] = R.dy; out.fin_d[3 * ray + 2] = R.dz; }
                        if (out.nshots) out.nshots[ray] = bounce;
                    }
                } else {
                    out.pid[ray] = h ? pid : -1;
                    if (out.t) out.t[ray] = h ? closest : 0.0;
                    if (out.xyz) { out.xyz[3 * ray] = h ? bx : 0.0; out.xyz[3 * ray + 1] = h ? by : 0.0; out.xyz[3 * ray + 2] = h ? bz : 0.0; }
                    if (out.uv) { out.uv[2 * ray] = h ? eu : 0.0; out.uv[2 * ray + 1] = h ? ev : 0.0; }
                    if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }   // the KDTree never moves a ray
                }
                fin = 2;
            }
            if (state == ST_NEED_RAY) {
                if (next < N) {
                    ray = next; next += stride;
                    R.x = o[3 * ray]; R.y = o[3 * ray + 1]; R.z = o[3 * ray + 2];
                    R.dx = d[3 * ray]; R.dy = d[3 * ray + 1]; R.dz = d[3 * ray + 2];
                    or1 = o1a ? o1a[ray] : -1; or2 = o2a ? o2a[ray] : -1;
                    blind = rid ? (rid[ray] == 0) : false;
                    bounce = 0;
                    state = ST_NEED_SETUP;
                } else {
                    state = ST_DONE;
                }
            }
            if (state == ST_NEED_SETUP) {
                state = ST_WALK; fin = 2;
                hit = false; closest = DBL_MAX; pid = -1; eu = 0; ev = 0; last = 0xffffffffu;
                lpos = 0; lend = 0; bmask = 0;
                sp = 0; stack[sp++] = 0;
                inv[0] = 1.0 / R.dx; inv[1] = 1.0 / R.dy; inv[2] = 1.0 / R.dz;
                fdx = (float)R.dx; fdy = (float)R.dy; fdz = (float)R.dz;
                fdd = fmaf(fdx, fdx, fmaf(fdy, fdy, fdz * fdz));
#if HARE_KD_BOX
                fix = cull_rcp(fdx); fiy = cull_rcp(fdy); fiz = cull_rcp(fdz);
#endif
                if (blind) fin = 0;   // Ray_ID == 0 against a fresh mailbox: every polygon is rejected (KDTree.cs:58-66, 224-229)
            }
        }
        // ------------------------------------------------------------------ N phase: pop until a reachable, non-empty leaf
        if (doN && needN) {
#pragma unroll 1
            for (int guard = 0; guard < N_MAX; ++guard) {
                if (sp == 0) { fin = hit ? 1 : 0; break; }
                const int ni = stack[--sp];
                const double2* q = reinterpret_cast<const double2*>(T.nodes + ni);
                const double2 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2), dd = __ldg(q + 3);
                double t_in;
                if (!kd_box_reachable(a, b, cc, R, inv, closest, t_in)) continue;
                c.cell();
                const int left = __double2loint(dd.y), axis = __double2hiint(dd.y);
                if (left < 0) {
                    const uint32_t off = (uint32_t)__double2loint(dd.x), cnt = (uint32_t)__double2hiint(dd.x);
                    lpos = off; lend = off + cnt;
                    if (cnt) {
                        fpx = (float)fma(R.dx, t_in, R.x); fpy = (float)fma(R.dy, t_in, R.y); fpz = (float)fma(R.dz, t_in, R.z);
#if HARE_KD_BOX
                        fpx *= fix; fpy *= fiy; fpz *= fiz;
#endif
                        break;
                    }
                } else {
                    const double oa = axis == 0 ? R.x : (axis == 1 ? R.y : R.z);
                    const bool right_first = oa > dd.x;                 // child on the origin's side first
                    if (sp + 2 <= HARE_KD_MAXSTACK) { stack[sp++] = right_first ? left : left + 1; stack[sp++] = right_first ? left + 1 : left; }
                }
            }
        }
        // ------------------------------------------------------------------ C phase: cull a batch of leaf entries
        if (state == ST_WALK && fin == 2 && bmask == 0 && lpos < lend) {
            const uint32_t n = min((uint32_t)HARE_KD_CB, lend - lpos);
            uint32_t m = 0;
#if HARE_KD_BOX
            // every list entry carries its polygon's padded box and its id (lo.w): 32 contiguous bytes, no dependent load
            static_assert(HARE_KD_CB == 8, "two groups of four");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 lo[4], hi[4];
#if HARE_KD_ENTRY_PBOX
                uint32_t ids[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ids[j] = __ldg(T.lists + lpos + (4 * h + j < (int)n ? 4 * h + j : 0));
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float4* e = T.pbox + 2 * (size_t)ids[j]; lo[j] = __ldg(e); hi[j] = __ldg(e + 1); }
#else
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4* e = T.lbox + 2 * (size_t)(lpos + (4 * h + j < (int)n ? 4 * h + j : 0));
                    lo[j] = __ldg(e); hi[j] = __ldg(e + 1);
                }
#endif
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#if HARE_KD_ENTRY_PBOX
                    const uint32_t i = ids[j];   // poly_origin skip (:220); mailbox: a polygon counts once (:224-229)
#else
                    const uint32_t i = __float_as_uint(lo[j].w);   // poly_origin skip (:220); mailbox: a polygon counts once (:224-229)
#endif
                    bid[4 * h + j] = i;
                    const bool keep = (4 * h + j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                                      !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz);
                    m |= keep ? (1u << (4 * h + j)) : 0u;
                }
            }
            lpos += n;
            if (COUNT) c.entries += n;
#else
#pragma unroll
            for (int j = 0; j < HARE_KD_CB; ++j) bid[j] = __ldg(T.lists + lpos + (j < (int)n ? j : 0));
            float4 s[HARE_KD_CB];
#pragma unroll
            for (int j = 0; j < HARE_KD_CB; ++j) s[j] = __ldg(T.sph + bid[j]);
            lpos += n;
            if (COUNT) c.entries += n;
#pragma unroll
            for (int j = 0; j < HARE_KD_CB; ++j) {
                const uint32_t i = bid[j];   // poly_origin skip (:220); mailbox: a polygon counts once (:224-229)
                const bool keep = (j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                                  !cull_sphere(s[j], fpx, fpy, fpz, fdx, fdy, fdz, fdd);
                m |= keep ? (1u << j) : 0u;
            }
#endif
            bmask = m;
        }
        // ------------------------------------------------------------------ T phase: the exact FP64 test (slow path: u, v)
        if (doT && state == ST_WALK && fin == 2 && bmask != 0) {
            const int kk = __ffs(bmask) - 1;
            uint32_t pend = bid[0];
#pragma unroll
            for (int j = 1; j < HARE_KD_CB; ++j) pend = (kk == j) ? bid[j] : pend;
            bmask &= bmask - 1u;
            last = pend;
            c.test();
            double P[16], t, u, v;
            load_poly(polys, pend, P);
            if (poly_intersect<true>(P, R, t, u, v) && t > 0.0000000001) {
                if (t < closest) { closest = t; hit = true; pid = (int)pend; eu = u; ev = v; }
                else if (t == closest && (int)pend != pid && kd_dfs_before(T, polys, R, pend, (uint32_t)pid)) { pid = (int)pend; eu = u; ev = v; }
            }
        }
    }
    if (CHAIN) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) shots += __shfl_xor_sync(0xffffffffu, shots, off);
        if ((threadIdx.x & 31) == 0 && shots) atomicAdd(out.total_shots, (unsigned long long)shots);
    }
    flush_counters<COUNT>(c, out.counters);
}

}  // namespace hare
