// vg_walk.cuh -- K1/K5 for Voxel_Grid: persistent, warp-synchronous phased traversal.
//
// A thread owns one ray (or one reflection chain) at a time and runs a small state machine.
// Each trip round the main loop the whole warp goes through three phases together:
//
//   S  (batched)  lanes that need a new ray, or a DDA set-up for the next bounce, do it --
//                 but only once S_BATCH lanes want it (or nobody can do anything else), so the
//                 nine divides of the set-up are paid by several lanes at once;
//   W  (cheap)    every lane whose voxel list is exhausted advances its 3D-DDA -- skipping empty
//                 voxels on the occupancy bitmap (staged in shared memory) without touching the
//                 cell table -- until it stands in a non-empty voxel, has accepted its carried
//                 candidate, or left the grid (at most W_MAX voxels per trip);
//   T  (dense)    every lane with a list entry left takes the next one, fetches the 128-byte
//                 polygon record (8 x LDG.128) and runs the FP64 Moller-Trumbore test.
//
// The FP64 pipe is the busiest unit of this path (ncu, profiles/), and the polygon test is
// ~10x the cost of a voxel step; one-ray-per-thread "while-while" code left 2.6 of 32 lanes
// active per issued instruction.  Phasing keeps the expensive T phase converged.
//
// Result semantics are exactly Voxel_Grid.Shoot's (Voxel_Grid.cs:351-552): same voxel sequence,
// ascending list order, strict t < tmin, carried candidate accepted only inside the current
// inflated voxel, leaving the grid is a miss.
#pragma once
#include "shoot.cuh"

namespace hare {

struct WalkOut {
    double* __restrict__ t; double* __restrict__ xyz; int32_t* __restrict__ pid; double* __restrict__ uv; double* __restrict__ omoved;   // Shoot
    int32_t* __restrict__ ev_pid; double* __restrict__ ev_t; double* __restrict__ fin_o; double* __restrict__ fin_d;                   // chain
    int32_t* __restrict__ nshots; unsigned long long* __restrict__ total_shots;
    unsigned long long* __restrict__ counters;
};

enum : int { ST_NEED_RAY = 0, ST_NEED_SETUP = 1, ST_WALK = 2, ST_DONE = 3 };

#ifndef HARE_VG_THREADS
#define HARE_VG_THREADS 640
#endif

// OCC_SMEM: the occupancy bitmap (1 bit per voxel) is staged in shared memory once per CTA, so an
// empty-voxel step costs a 29-cycle LDS instead of an L1/L2 round trip.  One 512-thread CTA per SM.
template <bool CHAIN, bool COUNT, bool OCC_SMEM, int S_BATCH, int W_MAX>
__global__ void __launch_bounds__(HARE_VG_THREADS, 1)
vg_walk_kernel(const VGrid g, const PolyRec* __restrict__ polys,
               const double* __restrict__ o, const double* __restrict__ d,
               const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid,
               long long N, int order, const WalkOut out) {
    extern __shared__ uint32_t s_occ[];
    if (OCC_SMEM) {
        const uint32_t words = ((uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz + 31u) >> 5;
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) s_occ[w] = __ldg(g.occ + w);
        __syncthreads();
    }
    CntT<COUNT> c;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long next = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long ray = -1;

    Ray3 R = { 0, 0, 0, 0, 0, 0 };
    double tMaxX = 0, tMaxY = 0, tMaxZ = 0, tDeltaX = 0, tDeltaY = 0, tDeltaZ = 0;
    double tmin = DBL_MAX, t_start = 0;
    float fdx = 0, fdy = 0, fdz = 0, fdd = 0, fpx = 0, fpy = 0, fpz = 0;   // FP32 copy of d, |d|^2 and of a ray point at the current voxel (cull only)
    int X = 0, Y = 0, Z = 0, stepX = 1, stepY = 1, stepZ = 1;
    int pid = -1, or1 = -1, or2 = -1, bounce = 0;
    uint32_t lpos = 0, lend = 0, last = 0xffffffffu, ci = 0;
    uint32_t bid0 = 0, bid1 = 0, bid2 = 0, bid3 = 0, bmask = 0;   // batch of up to 4 list entries; bit k = entry k survived the cull
    bool have = false, blind = false;
    int state = ST_NEED_RAY;
    int fin = 2;              // 2 = Shoot still running; otherwise its status: 1 hit, 0 miss, -2 fault
    unsigned int shots = 0;
    const int strideX = g.ny * g.nz, strideY = g.nz;

    // list range of voxel ci (empty voxels never touch the cell table) and its first polygon index
    auto enter_cell = [&](double t_in) {   // t_in: ray parameter at which this voxel was entered
        c.cell();
        lpos = 0; lend = 0;
        const uint32_t word = OCC_SMEM ? s_occ[ci >> 5] : __ldg(g.occ + (ci >> 5));
        if (!blind && ((word >> (ci & 31)) & 1u)) {
            const uint2 h = __ldg(g.cells + ci);
            lpos = h.x; lend = h.x + h.y;
            fpx = (float)fma(R.dx, t_in, R.x); fpy = (float)fma(R.dy, t_in, R.y); fpz = (float)fma(R.dz, t_in, R.z);
        }
    };

    while (true) {
        // ------------------------------------------------------------------ S phase
        const unsigned want = __ballot_sync(0xffffffffu, state == ST_NEED_RAY || state == ST_NEED_SETUP);
        const unsigned busy = __ballot_sync(0xffffffffu, state == ST_WALK);
        if (want == 0 && busy == 0) break;
        if (want && (__popc(want) >= S_BATCH || busy == 0)) {
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
            if (state == ST_NEED_SETUP) {   // Voxel_Grid.cs:357-422
                state = ST_WALK; fin = 2;
                have = false; bmask = 0; tmin = DBL_MAX; pid = -1; last = 0xffffffffu; t_start = 0; lpos = 0; lend = 0;
                X = floor_to_int((R.x - g.ominx) / g.vdx);
                Y = floor_to_int((R.y - g.ominy) / g.vdy);
                Z = floor_to_int((R.z - g.ominz) / g.vdz);
                if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) {
                    if (!obox_enter(g, R, t_start)) fin = 0;
                    else {
                        X = floor_to_int((R.x - g.ominx + R.dx * 1E-6) / g.vdx);
                        Y = floor_to_int((R.y - g.ominy + R.dy * 1E-6) / g.vdy);
                        Z = floor_to_int((R.z - g.ominz + R.dz * 1E-6) / g.vdz);
                        if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) fin = -2;
                    }
                }
                if (fin == 2) {
                    const bool nx_ = R.dx < 0, ny_ = R.dy < 0, nz_ = R.dz < 0;
                    stepX = nx_ ? -1 : 1; stepY = ny_ ? -1 : 1; stepZ = nz_ ? -1 : 1;
                    tMaxX = ((nx_ ? vox_min(X, g.vdx, g.ominx) : vox_max(X, g.vdx, g.ominx)) - R.x) / R.dx;
                    tMaxY = ((ny_ ? vox_min(Y, g.vdy, g.ominy) : vox_max(Y, g.vdy, g.ominy)) - R.y) / R.dy;
                    tMaxZ = ((nz_ ? vox_min(Z, g.vdz, g.ominz) : vox_max(Z, g.vdz, g.ominz)) - R.z) / R.dz;
                    tDeltaX = g.vdx / R.dx * (nx_ ? -1.0 : 1.0);
                    tDeltaY = g.vdy / R.dy * (ny_ ? -1.0 : 1.0);
                    tDeltaZ = g.vdz / R.dz * (nz_ ? -1.0 : 1.0);
                    ci = ((uint32_t)X * (uint32_t)g.ny + (uint32_t)Y) * (uint32_t)g.nz + (uint32_t)Z;
                    fdx = (float)R.dx; fdy = (float)R.dy; fdz = (float)R.dz;
                    fdd = fmaf(fdx, fdx, fmaf(fdy, fdy, fdz * fdz));
                    enter_cell(0.0);
                }
            }
        }
        // ------------------------------------------------------------------ W phase: voxel steps
        if (state == ST_WALK && fin == 2 && bmask == 0 && lpos >= lend) {
#pragma unroll 1
            for (int guard = 0; guard < W_MAX; ++guard) {
                // list exhausted: Voxels[X,Y,Z].IsPointInBox(candidate)?   Voxel_Grid.cs:496-500
                if (have) {
                    const double bx = R.x + R.dx * tmin, by = R.y + R.dy * tmin, bz = R.z + R.dz * tmin;
                    const bool in = !(bx < vox_min(X, g.vdx, g.ominx)) & !(by < vox_min(Y, g.vdy, g.ominy)) & !(bz < vox_min(Z, g.vdz, g.ominz)) &
                                    !(bx > vox_max(X, g.vdx, g.ominx)) & !(by > vox_max(Y, g.vdy, g.ominy)) & !(bz > vox_max(Z, g.vdz, g.ominz));
                    if (in) { fin = 1; break; }
                }
                // next voxel   Voxel_Grid.cs:504-550: X only if strictly below both, Y only if below Z, else Z
                const bool xy = tMaxX < tMaxY, xz = tMaxX < tMaxZ, yz = tMaxY < tMaxZ;
                const bool goX = xy & xz, goY = (!xy) & yz;
                const bool goZ = !(goX | goY);
                const double t_in = goX ? tMaxX : (goY ? tMaxY : tMaxZ);
                const double nX = tMaxX + tDeltaX, nY = tMaxY + tDeltaY, nZ = tMaxZ + tDeltaZ;
                tMaxX = goX ? nX : tMaxX; tMaxY = goY ? nY : tMaxY; tMaxZ = goZ ? nZ : tMaxZ;
                X += goX ? stepX : 0; Y += goY ? stepY : 0; Z += goZ ? stepZ : 0;
                ci += (uint32_t)(goX ? stepX * strideX : (goY ? stepY * strideY : stepZ));
                if ((unsigned)X >= (unsigned)g.nx || (unsigned)Y >= (unsigned)g.ny || (unsigned)Z >= (unsigned)g.nz) { fin = 0; break; }
                enter_cell(t_in);
                if (lpos < lend) break;
            }
        }
        // ------------------------------------------------------------------ C phase: cull a batch of list entries
        if (state == ST_WALK && fin == 2 && bmask == 0 && lpos < lend) {
            // next (up to) four list entries, ascending polygon index: ids and bounding spheres are fetched as
            // two groups of independent loads, then culled in FP32; survivors wait in bmask for the T phase
            const uint32_t n = min(4u, lend - lpos);
            bid0 = __ldg(g.cell_poly + lpos);
            bid1 = (n > 1) ? __ldg(g.cell_poly + lpos + 1) : bid0;
            bid2 = (n > 2) ? __ldg(g.cell_poly + lpos + 2) : bid0;
            bid3 = (n > 3) ? __ldg(g.cell_poly + lpos + 3) : bid0;
            const float4 s0 = __ldg(g.sph + bid0), s1 = __ldg(g.sph + bid1), s2 = __ldg(g.sph + bid2), s3 = __ldg(g.sph + bid3);
            lpos += n;
            if (COUNT) c.entries += n;
            // poly_origin skip (Voxel_Grid.cs:477); a polygon already tested for this ray cannot change the result
            auto keep = [&](uint32_t i, const float4& s) {
                return !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid) && !cull_sphere(s, fpx, fpy, fpz, fdx, fdy, fdz, fdd);
            };
            bmask = (keep(bid0, s0) ? 1u : 0u) | ((n > 1 && keep(bid1, s1)) ? 2u : 0u) |
                    ((n > 2 && keep(bid2, s2)) ? 4u : 0u) | ((n > 3 && keep(bid3, s3)) ? 8u : 0u);
        }
        // ------------------------------------------------------------------ F phase: the Shoot is over
        if (fin != 2) {
            const double ev_t = (fin == 1) ? tmin + t_start : 0.0;
            const int ev_p = (fin == 1) ? pid : (fin == -2 ? -2 : -1);
            // X_Point = R + d*t of the winning test (Hare_Geometry_Polygons.cs:802): same bits whenever it is formed
            const double bx = R.x + R.dx * tmin, by = R.y + R.dy * tmin, bz = R.z + R.dz * tmin;
            if (fin == 1) c.hit();
            state = ST_NEED_RAY;
            if (CHAIN) {
                ++shots;
                if (out.ev_pid) out.ev_pid[ray * order + bounce] = ev_p;
                if (out.ev_t) out.ev_t[ray * order + bounce] = ev_t;
                ++bounce;
                if (fin == 1) {
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
                const bool h = fin == 1;
                out.pid[ray] = ev_p;
                if (out.t) out.t[ray] = ev_t;
                if (out.xyz) { out.xyz[3 * ray] = h ? bx : 0.0; out.xyz[3 * ray + 1] = h ? by : 0.0; out.xyz[3 * ray + 2] = h ? bz : 0.0; }
                if (out.uv) { out.uv[2 * ray] = 0.0; out.uv[2 * ray + 1] = 0.0; }
                if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }
            }
            fin = 2; lpos = 0; lend = 0; bmask = 0;
        }
        // ------------------------------------------------------------------ T phase: the exact FP64 test
        if (bmask) {
            const uint32_t pend = (bmask & 1u) ? bid0 : ((bmask & 2u) ? bid1 : ((bmask & 4u) ? bid2 : bid3));   // lowest survivor first
            bmask &= bmask - 1u;
            last = pend;
            c.test();
            double P[16], t = 0;
            load_poly(polys, pend, P);
            // Polygon.Ray_Side picks the winding (Hare_Geometry_Polygons.cs:601-606, 637-660, 784-823):
            //   side ? (P0,P1,P2) then (P2,P3,P0) : (P2,P1,P0) then (P0,P3,P2)
            const bool side = !(dot3(R.dx, R.dy, R.dz, P[12], P[13], P[14]) < 0);
            const double ax = side ? P[0] : P[6], ay = side ? P[1] : P[7], az = side ? P[2] : P[8];
            const double cx = side ? P[6] : P[0], cy = side ? P[7] : P[1], cz = side ? P[8] : P[2];
            bool hit = ray_x_tri_fast1(R, ax, ay, az, P[3], P[4], P[5], cx, cy, cz, t);
            if (!hit && P[15] == 4.0) hit = ray_x_tri_fast1(R, cx, cy, cz, P[9], P[10], P[11], ax, ay, az, t);
            if (hit && t > 0.0000000001 && t < tmin) { tmin = t; pid = (int)pend; have = true; }
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
