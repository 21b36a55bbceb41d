// oct_walk.cuh -- K3 (and K5 on an Octree): persistent, warp-synchronous phased traversal of
// Hare's Octree ("Octree - alt.cs":159-306), built like vg_walk.cuh.
//
// A thread owns one ray (or one reflection chain) at a time.  Each trip round the main loop the
// warp goes through four phases together:
//
//   S  (batched)  finish the Shoot that ended (event out; in a chain, reflect), fetch the next ray,
//                 reciprocals and root interval (:165-190);
//   N  (cheap)    replay the reference's LIFO walk until a leaf with a non-empty list is reached:
//                 one frame per level (first child, parent interval, next octant) reproduces the
//                 far-first pop order and the push-time filter (:245-272) lazily, the pop-time prunes
//                 (:207-211) are applied when a node is entered;
//   C  (cheap)    next (up to) eight entries of the leaf list: poly_origin and duplicate skip, then the
//                 conservative FP32 sphere reject (cull_sphere) in a frame local to the leaf;
//   T  (dense)    one exact test: 128-byte record, slow-path Moller-Trumbore with u, v (:224), strict
//                 t < closestT, early `return` when closestT <= nodeTmin (:233-237).
//
// Entries are taken in stored order and survivors are tested lowest first, so the sequence of
// closestT updates -- and with it the early return and the pop-time prune -- is the reference's.
#pragma once
#include "shoot.cuh"
#include "vg_walk.cuh"   // WalkOut, ST_* states

namespace hare {

#ifndef HARE_OCT_CB
#define HARE_OCT_CB 8   /* leaf entries culled per C round */
#endif
// 1: chunks and leaf entries are culled on padded FP32 bounding boxes (cull_box; OctDev::cbox / lbox); 0: on spheres
#ifndef HARE_OCT_BOX
#define HARE_OCT_BOX 1
#endif
// 1: the per-entry box is read from the per-POLYGON table (OctDev::pbox, by id) instead of a per-entry copy: a dependent load, but
// only for entries of the few chunks that survive, and the 32 bytes per list entry no longer compete for L2
#ifndef HARE_OCT_ENTRY_PBOX
#define HARE_OCT_ENTRY_PBOX 1
#endif
#ifndef HARE_OCT_THREADS
#define HARE_OCT_THREADS 640
#endif

template <bool CHAIN, bool COUNT, int S_BATCH, int N_MAX, int N_BATCH, int T_BATCH>
__global__ void __launch_bounds__(HARE_OCT_THREADS, 1)
oct_walk_kernel(const OctDev T, const PolyRec* __restrict__ polys,
                const double* __restrict__ o, const double* __restrict__ d,
                const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a,
                long long N, int order, const WalkOut out) {
    CntT<COUNT> c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long next = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long ray = -1;

    Ray3 R = { 0, 0, 0, 0, 0, 0 };
    double ix = 0, iy = 0, iz = 0;                 // guarded reciprocals :165-167
    double closest = DBL_MAX, eu = 0, ev = 0;      // closestT and the u, v of the best event
    double ca = 0, cb = 0;                         // interval of the node being entered / of the current leaf
    float fdx = 0, fdy = 0, fdz = 0, fdd = 0, fpx = 0, fpy = 0, fpz = 0;
#if HARE_OCT_BOX
    float fix = 0, fiy = 0, fiz = 0;   // FP32 reciprocal direction (cull_box); with it fpx.. hold p/d instead of p
#endif
    int fchild[HARE_OCT_MAXLVL]; double fa[HARE_OCT_MAXLVL], fb[HARE_OCT_MAXLVL]; uint32_t fq[HARE_OCT_MAXLVL];
    int sp = -1, cur = 0, sgn = 0;
    int pid = -1, or1 = -1, or2 = -1, bounce = 0;
    uint32_t lpos = 0, lend = 0, last = 0xffffffffu;
    uint32_t bid[HARE_OCT_CB] = { 0 }, bmask = 0;
    uint32_t emask = 0, cpos = 0, cidx = 0;   // surviving chunks of the current group of 8 chunks, its list position, next chunk index   // batch of up to 8 leaf entries; bit k = entry k survived the cull
    bool have_cur = false, hit = false;
    int state = ST_NEED_RAY;
    int fin = 2;   // 2 = running; 1 hit, 0 miss
    unsigned int shots = 0;

    while (true) {
        // ------------------------------------------------------------------ S phase
        const unsigned want = __ballot_sync(0xffffffffu, (state == ST_NEED_RAY || state == ST_NEED_SETUP) || (state == ST_WALK && fin != 2));
        const unsigned busy = __ballot_sync(0xffffffffu, state == ST_WALK && fin == 2);
        if (want == 0 && busy == 0) break;
        // N and T are dear and usually wanted by few lanes while C (cheap, long leaf lists) is wanted by most:
        // they run once N_BATCH / T_BATCH lanes wait for them, or when nothing cheaper is left to do
        const bool needN = state == ST_WALK && fin == 2 && bmask == 0 && emask == 0 && lpos >= lend;
        const bool needC = state == ST_WALK && fin == 2 && bmask == 0 && (emask != 0 || lpos < lend);
        const bool needT = state == ST_WALK && fin == 2 && bmask != 0;
        const int nN = __popc(__ballot_sync(0xffffffffu, needN)), nC = __popc(__ballot_sync(0xffffffffu, needC));
        const int nT = __popc(__ballot_sync(0xffffffffu, needT));
        const bool doT = nT > 0 && (nT >= T_BATCH || nC == 0);
        const bool doN = nN > 0 && (nN >= N_BATCH || (nC == 0 && !doT));
        if (want && (__popc(want) >= S_BATCH || busy == 0 || (nC == 0 && !doT && !doN))) {
            if (state == ST_WALK && fin != 2) {
                // ---- the Shoot is over
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
                    if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }   // the Octree never moves a ray
                }
                fin = 2;
            }
            if (state == ST_NEED_RAY) {
                if (next < N) {
                    ray = next; next += stride;
                    R.x = o[3 * ray]; R.y = o[3 * ray + 1]; R.z = o[3 * ray + 2];
                    R.dx = d[3 * ray]; R.dy = d[3 * ray + 1]; R.dz = d[3 * ray + 2];
                    or1 = o1a ? o1a[ray] : -1; or2 = o2a ? o2a[ray] : -1;
                    bounce = 0;
                    state = ST_NEED_SETUP;
                } else {
                    state = ST_DONE;
                }
            }
            if (state == ST_NEED_SETUP) {
                state = ST_WALK; fin = 2;
                hit = false; closest = DBL_MAX; pid = -1; eu = 0; ev = 0; last = 0xffffffffu;
                lpos = 0; lend = 0; bmask = 0; emask = 0; sp = -1;
                ix = fabs(R.dx) > 1e-16 ? 1.0 / R.dx : 1e16;
                iy = fabs(R.dy) > 1e-16 ? 1.0 / R.dy : 1e16;
                iz = fabs(R.dz) > 1e-16 ? 1.0 / R.dz : 1e16;
                oct_interval(T.nodes, R, ix, iy, iz, ca, cb);
                if (cb < ca || cb < 0) fin = 0;                       // :185-190
                // A ray with a NaN/Inf component is reported as a miss (documented deviation): the
                // walk below uses oct_interval_finite, which assumes finite operands.  (In the reference such a ray
                // makes every comparison false and every t NaN, which also ends in a miss.)
                if (!(isfinite(R.x) && isfinite(R.y) && isfinite(R.z) && isfinite(R.dx) && isfinite(R.dy) && isfinite(R.dz))) fin = 0;
                sgn = (R.dx >= 0 ? 0 : 4) | (R.dy >= 0 ? 0 : 2) | (R.dz >= 0 ? 0 : 1);   // ComputeTraversalOrder: order[q] = q ^ sgn
                cur = 0; have_cur = true;
                fdx = (float)R.dx; fdy = (float)R.dy; fdz = (float)R.dz;
                fdd = fmaf(fdx, fdx, fmaf(fdy, fdy, fdz * fdz));
#if HARE_OCT_BOX
                // cull_box frame: p = the point where the ray enters the root cube (the origin itself when it starts inside),
                // formed in FP64 and then rounded -- a ray shot from far outside the model must not lose the millimetres
                // the padding allows to FP32; kept as p/d
                {
                    const double te = ca > 0.0 ? ca : 0.0;
                    fix = cull_rcp(fdx); fiy = cull_rcp(fdy); fiz = cull_rcp(fdz);
                    fpx = (float)fma(R.dx, te, R.x) * fix; fpy = (float)fma(R.dy, te, R.y) * fiy; fpz = (float)fma(R.dz, te, R.z) * fiz;
                }
#endif
            }
        }
        // ------------------------------------------------------------------ N phase: walk to the next leaf
        if (doN && needN) {
#pragma unroll 1
            for (int guard = 0; guard < N_MAX; ++guard) {
                if (have_cur) {
                    have_cur = false;
                    if (!(cb < ca || cb < 0) && !(hit && closest <= ca)) {            // pop-time prunes :207-211
                        c.cell();
                        const uint4 m = __ldg(reinterpret_cast<const uint4*>(T.nodes + cur) + 3);   // first_child, list_off, list_cnt
                        if ((int)m.x < 0) {
                            lpos = m.y; lend = m.y + m.z; cidx = m.w;
                            if (lpos < lend) {
#if !HARE_OCT_BOX
                                // leaf-local FP32 frame for cull_sphere: the ray point where the leaf is entered
                                const double te = ca > 0.0 ? ca : 0.0;
                                fpx = (float)fma(R.dx, te, R.x); fpy = (float)fma(R.dy, te, R.y); fpz = (float)fma(R.dz, te, R.z);
#endif
                                break;
                            }
                        } else if (sp + 1 < HARE_OCT_MAXLVL) {
                            // frame: bit q = the q-th octant in near->far order (child q ^ sgn) still has to be popped; octants
                            // whose subtree holds no polygon (node.pad, built at upload) are never entered: they cannot change
                            // closestT, so neither the result nor the walk after them depends on them
                            uint32_t pm = 0;
#pragma unroll
                            for (int q = 0; q < 8; ++q) pm |= ((m.w >> (q ^ sgn)) & 1u) << q;
                            ++sp; fchild[sp] = (int)m.x; fa[sp] = ca; fb[sp] = cb; fq[sp] = pm;
                        }
                    }
                    continue;
                }
                if (sp < 0) { fin = hit ? 1 : 0; break; }                              // stack empty :276-283
                const uint32_t pm = fq[sp];
                if (pm == 0) { --sp; continue; }
                const int q = 31 - __clz(pm);                                         // pushed near->far, popped far->near
                fq[sp] = pm & ~(1u << q);
                const int child = fchild[sp] + (q ^ sgn);
#if HARE_OCT_BOX
                {   // the ray's line misses everything listed below this child: entering it could change nothing
                    const float4* e = T.nbox + 2 * (size_t)child;
                    if (cull_box(__ldg(e), __ldg(e + 1), fpx, fpy, fpz, fix, fiy, fiz)) continue;
                }
#endif
                double lo, hi;
                oct_interval_finite(T.nodes + child, R, ix, iy, iz, lo, hi);
                const double pa = fa[sp], pb = fb[sp];
                if (hi < lo || hi < 0 || lo > pb || hi < pa) continue;                // push-time filter :268
                cur = child; ca = fmax(lo, pa); cb = fmin(hi, pb); have_cur = true;
            }
        }
        // ------------------------------------------------------------------ C phase: cull a batch of leaf entries
        if (state == ST_WALK && fin == 2 && bmask == 0 && (emask != 0 || lpos < lend)) {
            static_assert(HARE_OCT_CB == HARE_OCT_CHUNK, "one C round culls one chunk");
            if (emask == 0) {
                // next group of (up to) eight chunks = 64 list entries: cull the chunks first
                const uint32_t left = lend - lpos, nch = min(8u, (left + HARE_OCT_CHUNK - 1) / HARE_OCT_CHUNK);
                uint32_t em = 0;
#if HARE_OCT_BOX
                // the whole group of 64 entries first: most groups of a long leaf list are nowhere near the ray
                const float4* ge = T.gbox + 2 * (size_t)(cidx >> 3);
                if (!cull_box(__ldg(ge), __ldg(ge + 1), fpx, fpy, fpz, fix, fiy, fiz))
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 lo[4], hi[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4* e = T.cbox + 2 * (size_t)(cidx + (4 * h + j < (int)nch ? 4 * h + j : 0));
                        lo[j] = __ldg(e); hi[j] = __ldg(e + 1);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        em |= (4 * h + j < (int)nch && !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz)) ? (1u << (4 * h + j)) : 0u;
                }
#else
                float4 cs[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) cs[j] = __ldg(T.csph + cidx + (j < (int)nch ? j : 0));
#pragma unroll
                for (int j = 0; j < 8; ++j) em |= (j < (int)nch && !cull_sphere(cs[j], fpx, fpy, fpz, fdx, fdy, fdz, fdd)) ? (1u << j) : 0u;
#endif
                emask = em; cpos = lpos; cidx += nch;
                const uint32_t adv = min(left, 8u * HARE_OCT_CHUNK);
                lpos += adv;
                if (COUNT) c.entries += adv;
            }
            if (emask != 0) {
                // lowest surviving chunk: its (up to) eight entries
                const int kc = __ffs(emask) - 1;
                emask &= emask - 1u;
                const uint32_t base = cpos + (uint32_t)kc * HARE_OCT_CHUNK;
                const uint32_t n = min((uint32_t)HARE_OCT_CB, lend - base);
                uint32_t m = 0;
                // poly_origin skip (:218); a polygon already tested for this ray (it sits in several leaves) cannot
                // change anything: its t is not below closestT any more, so neither the update nor the early return fires
#if HARE_OCT_BOX
                // every list entry carries its polygon's padded box and its id (lo.w): 32 contiguous bytes, no dependent load
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float4 lo[4], hi[4];
#if HARE_OCT_ENTRY_PBOX
                    uint32_t ids[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) ids[j] = __ldg(T.lists + base + (4 * h + j < (int)n ? 4 * h + j : 0));
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float4* e = T.pbox + 2 * (size_t)ids[j]; lo[j] = __ldg(e); hi[j] = __ldg(e + 1); }
#else
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4* e = T.lbox + 2 * (size_t)(base + (4 * h + j < (int)n ? 4 * h + j : 0));
                        lo[j] = __ldg(e); hi[j] = __ldg(e + 1);
                    }
#endif
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#if HARE_OCT_ENTRY_PBOX
                        const uint32_t i = ids[j];
#else
                        const uint32_t i = __float_as_uint(lo[j].w);
#endif
                        bid[4 * h + j] = i;
                        const bool keep = (4 * h + j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                                          !cull_box(lo[j], hi[j], fpx, fpy, fpz, fix, fiy, fiz);
                        m |= keep ? (1u << (4 * h + j)) : 0u;
                    }
                }
#else
#pragma unroll
                for (int j = 0; j < HARE_OCT_CB; ++j) bid[j] = __ldg(T.lists + base + (j < (int)n ? j : 0));
                float4 s[HARE_OCT_CB];
#pragma unroll
                for (int j = 0; j < HARE_OCT_CB; ++j) s[j] = __ldg(T.sph + bid[j]);
#pragma unroll
                for (int j = 0; j < HARE_OCT_CB; ++j) {
                    const uint32_t i = bid[j];
                    const bool keep = (j < (int)n) && !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) &&
                                      !cull_sphere(s[j], fpx, fpy, fpz, fdx, fdy, fdz, fdd);
                    m |= keep ? (1u << j) : 0u;
                }
#endif
                bmask = m;
            }
        }
        // ------------------------------------------------------------------ T phase: the exact FP64 test (slow path: u, v)
        if (doT && state == ST_WALK && fin == 2 && bmask != 0) {
            const int kk = __ffs(bmask) - 1;            // lowest survivor first: stored list order
            uint32_t pend = bid[0];
#pragma unroll
            for (int j = 1; j < HARE_OCT_CB; ++j) pend = (kk == j) ? bid[j] : pend;
            bmask &= bmask - 1u;
            last = pend;
            c.test();
            double P[16], t, u, v;
            load_poly(polys, pend, P);
            if (poly_intersect<true>(P, R, t, u, v) && t > 0.0000000001) {
                if (t < closest) {
                    closest = t; hit = true; pid = (int)pend; eu = u; ev = v;
                    if (closest <= ca) fin = 1;                                         // early return :233-237 (ca = this leaf's nodeTmin)
                }
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
