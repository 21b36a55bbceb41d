// vg_wave.cuh -- K1/K5 for Voxel_Grid, second generation: a per-warp WAVEFRONT scheduler.
//
// The first generation tied one ray to one thread; every trip round its loop the warp ran the S, W, C, F and T
// phases one after the other, each with the subset of lanes whose ray happened to be in that phase: ncu
// showed ~10 of 32 lanes active per issued instruction (profiles/r1_ncu_v7_summary.txt).
//
// Here the ray state lives in SHARED MEMORY instead of registers.  Every warp owns a private pool of SLOTS
// (> 32) ray slots, stored structure-of-arrays so that any lane can work on any slot.  Each slot carries a
// one-byte phase tag.  Each trip the warp
//   1. counts its slots per phase (ballots over the tags),
//   2. picks the phase with the most ready slots,
//   3. compacts up to 32 slot indices of that phase, one per lane (popc-ranked, via a 32-byte smem list),
//   4. runs that ONE phase converged on (up to) 32 different rays, and writes the state and new tags back.
// No inter-warp communication exists (only __syncwarp), so there is nothing to deadlock on.
//
// Phases (the arithmetic of each is Voxel_Grid.Shoot's, operation for operation; Voxel_Grid.cs:351-552):
//   SF  finish a Shoot (write the event, reflect), fetch a new ray if the slot is empty, DDA set-up
//   W   up to W_MAX voxel steps on the border-padded occupancy bitmap in shared memory (no coordinates, no bounds checks)
//   C   next <= 4 list entries: (padded box, id) records in list order, FP32 conservative box cull
//   T   one exact FP64 polygon test (128-byte record, Ray_Side, Moller-Trumbore)
//
// Every per-slot function below is `HD`: tests/emu/ compiles the same functions for the host and replays
// the scheduler on the CPU against the oracle (logic check without a GPU, and lane-utilisation statistics).
#pragma once
#include "shoot.cuh"

namespace hare {

enum : uint32_t { PH_SF = 0, PH_W = 1, PH_C = 2, PH_T = 3, PH_DONE = 4, PH_COUNT = 4 };

// slot flags
enum : uint32_t {
    WF_NEGX = 1u, WF_NEGY = 2u, WF_NEGZ = 4u, WF_HAVE = 8u, WF_BLIND = 16u,
    WF_NORAY = 32u,                        // the slot holds no ray: SF fetches one
    WF_FIN_SHIFT = 6, WF_FIN_MASK = 3u << 6,   // 0 running, 1 hit, 2 miss, 3 fault (the reference throws)
    WF_BMASK_SHIFT = 8, WF_BMASK_MASK = 15u << 8,
    WF_BOUNCE_SHIFT = 16
};
enum : uint32_t { FIN_RUN = 0, FIN_HIT = 1, FIN_MISS = 2, FIN_FAULT = 3 };

// field indices of the structure-of-arrays pool
enum { D_OX, D_OY, D_OZ, D_DX, D_DY, D_DZ, D_TMX, D_TMY, D_TMZ, D_TDX, D_TDY, D_TDZ, D_TMIN, D_TSTART, D_COUNT };
enum { U_XYZ /* padded voxel index cp */, U_FLAGS, U_LPOS, U_LEND, U_PID, U_OR1, U_OR2, U_LAST, U_RAY, U_COUNT };

template <int SLOTS>
struct WavePool {
    double* dbl; uint32_t* u32; uint8_t* tag; uint8_t* sel;
    static_assert(SLOTS <= 255, "phase counts are packed into bytes");
    static constexpr size_t BYTES = (size_t)SLOTS * (D_COUNT * 8 + U_COUNT * 4 + 1) + 32;
    static constexpr size_t STRIDE = (BYTES + 15) & ~(size_t)15;
    HD void bind(unsigned char* base) {
        dbl = reinterpret_cast<double*>(base);
        u32 = reinterpret_cast<uint32_t*>(base + (size_t)SLOTS * D_COUNT * 8);
        tag = base + (size_t)SLOTS * (D_COUNT * 8 + U_COUNT * 4);
        sel = tag + SLOTS;
    }
    HD double& D(int f, int s) const { return dbl[f * SLOTS + s]; }
    HD uint32_t& U(int f, int s) const { return u32[f * SLOTS + s]; }
};

// which phase a slot waits for, from its state
HD uint32_t wave_tag(uint32_t fl, uint32_t lpos, uint32_t lend) {
    if (fl & WF_FIN_MASK) return PH_SF;
    if (fl & WF_BMASK_MASK) return PH_T;
    return lpos < lend ? PH_C : PH_W;
}

// scheduling policy: the phase with the most ready slots; ties go to the later phase (drain before refill)
HD int wave_pick(const int n[PH_COUNT]) {
    int best = PH_T, bn = n[PH_T];
    if (n[PH_C] > bn) { best = PH_C; bn = n[PH_C]; }
    if (n[PH_W] > bn) { best = PH_W; bn = n[PH_W]; }
    if (n[PH_SF] > bn) { best = PH_SF; bn = n[PH_SF]; }
    return bn > 0 ? best : -1;
}

// The walk addresses voxels by their index in a grid PADDED by one voxel on every side, cp = ((X+1)*(ny+2) + (Y+1))*(nz+2) + (Z+1),
// and tests one bit per padded voxel (VGrid::occp: list non-empty, or border).  Leaving the grid then shows up as a set bit, so a
// voxel step needs neither the three coordinates nor their bounds checks; coordinates are decoded from cp only where they are
// needed (an occupied voxel, the border, the in-voxel test of a carried candidate).
struct WaveGeom { uint32_t d1, pz; double inv1, inv2; };   // d1 = (ny+2)*(nz+2), pz = nz+2 and their reciprocals

HD WaveGeom wave_geom(const VGrid& g) {
    WaveGeom w;
    w.pz = (uint32_t)g.nz + 2u; w.d1 = ((uint32_t)g.ny + 2u) * w.pz;
    w.inv1 = 1.0 / (double)w.d1; w.inv2 = 1.0 / (double)w.pz;
    return w;
}

HD uint32_t wave_cp(const VGrid& g, int X, int Y, int Z) {
    return ((uint32_t)(X + 1) * ((uint32_t)g.ny + 2u) + (uint32_t)(Y + 1)) * ((uint32_t)g.nz + 2u) + (uint32_t)(Z + 1);
}

// cp -> unpadded coordinates (-1 or n on the border).  The FP64 quotient is off by at most one; the remainder fixes it.
HD void wave_decode(const WaveGeom& w, uint32_t cp, int& X, int& Y, int& Z) {
    uint32_t xp = (uint32_t)((double)cp * w.inv1);
    int32_t r = (int32_t)(cp - xp * w.d1);
    if (r < 0) { --xp; r += (int32_t)w.d1; } else if ((uint32_t)r >= w.d1) { ++xp; r -= (int32_t)w.d1; }
    uint32_t yp = (uint32_t)((double)r * w.inv2);
    int32_t z = r - (int32_t)(yp * w.pz);
    if (z < 0) { --yp; z += (int32_t)w.pz; } else if ((uint32_t)z >= w.pz) { ++yp; z -= (int32_t)w.pz; }
    X = (int)xp - 1; Y = (int)yp - 1; Z = (int)z - 1;
}

// bit of padded voxel cp: one bit per voxel, staged in shared memory when it fits
HD bool wave_bit(const uint32_t* occ, bool occ_smem, uint32_t cp) {
    const uint32_t word = occ_smem ? occ[cp >> 5] : hare_ldg(occ + (cp >> 5));
    return (word >> (cp & 31)) & 1u;
}

// ---- SF, part 1: the Shoot in slot s is over -> write its event; a chain reflects and goes on, or ends.
// Afterwards the slot either holds a ray that needs a set-up, or carries WF_NORAY.
template <bool CHAIN, bool COUNT, int SLOTS>
HD void wave_finish(const PolyRec* __restrict__ polys, const WavePool<SLOTS>& p, int s, int order, const WalkOut& out,
                    unsigned int& shots, CntT<COUNT>& c) {
    uint32_t fl = p.U(U_FLAGS, s);
    const uint32_t fin = (fl & WF_FIN_MASK) >> WF_FIN_SHIFT;
    if (fin == FIN_RUN) return;
    const double tmin = p.D(D_TMIN, s);
    Ray3 R = { p.D(D_OX, s), p.D(D_OY, s), p.D(D_OZ, s), p.D(D_DX, s), p.D(D_DY, s), p.D(D_DZ, s) };
    const int pid = (int)p.U(U_PID, s);
    const long long ray = (long long)p.U(U_RAY, s);
    const double ev_t = (fin == FIN_HIT) ? tmin + p.D(D_TSTART, s) : 0.0;
    const int ev_p = (fin == FIN_HIT) ? pid : (fin == FIN_FAULT ? -2 : -1);
    // X_Point = R + d*t of the winning test (Hare_Geometry_Polygons.cs:802)
    const double bx = R.x + R.dx * tmin, by = R.y + R.dy * tmin, bz = R.z + R.dz * tmin;
    if (fin == FIN_HIT) c.hit();
    fl &= ~(WF_FIN_MASK | WF_BMASK_MASK | WF_HAVE);
    if (CHAIN) {
        uint32_t bounce = fl >> WF_BOUNCE_SHIFT;
        ++shots;
        if (out.ev_pid) out.ev_pid[ray * order + bounce] = ev_p;
        if (out.ev_t) out.ev_t[ray * order + bounce] = ev_t;
        chain_row_xyz(out, ray, order, bounce, fin == FIN_HIT, bx, by, bz);
        chain_row_uv(out, ray, order, bounce, 0.0, 0.0);                       // Voxel_Grid events carry u = v = 0 (Voxel_Grid.cs:487-488)
        ++bounce;
        bool go_on = false;
        if (fin == FIN_HIT) {
            const double* P = polys[pid].v;
            const double nx = hare_ldg(P + 12), ny = hare_ldg(P + 13), nz = hare_ldg(P + 14);
            const double k = 2 * ((R.dx * nx) + (R.dy * ny) + (R.dz * nz));
            R.dx = R.dx - k * nx; R.dy = R.dy - k * ny; R.dz = R.dz - k * nz;
            R.x = bx; R.y = by; R.z = bz;
            p.D(D_OX, s) = R.x; p.D(D_OY, s) = R.y; p.D(D_OZ, s) = R.z;
            p.D(D_DX, s) = R.dx; p.D(D_DY, s) = R.dy; p.D(D_DZ, s) = R.dz;
            p.U(U_OR1, s) = (uint32_t)pid;
            go_on = (int)bounce < order;
        }
        fl = (fl & 0xffffu) | (bounce << WF_BOUNCE_SHIFT);
        if (!go_on) {
            for (int q = (int)bounce; q < order; ++q) {
                if (out.ev_pid) out.ev_pid[ray * order + q] = -3;
                if (out.ev_t) out.ev_t[ray * order + q] = 0;
            }
            chain_rows_clear(out, ray, order, (int)bounce);
            if (out.fin_o) { out.fin_o[3 * ray] = R.x; out.fin_o[3 * ray + 1] = R.y; out.fin_o[3 * ray + 2] = R.z; }
            if (out.fin_d) { out.fin_d[3 * ray] = R.dx; out.fin_d[3 * ray + 1] = R.dy; out.fin_d[3 * ray + 2] = R.dz; }
            if (out.nshots) out.nshots[ray] = (int32_t)bounce;
            fl |= WF_NORAY;
        }
    } else {
        const bool h = fin == FIN_HIT;
        out.pid[ray] = ev_p;
        if (out.t) out.t[ray] = ev_t;
        if (out.xyz) { out.xyz[3 * ray] = h ? bx : 0.0; out.xyz[3 * ray + 1] = h ? by : 0.0; out.xyz[3 * ray + 2] = h ? bz : 0.0; }
        if (out.uv) { out.uv[2 * ray] = 0.0; out.uv[2 * ray + 1] = 0.0; }
        if (out.omoved) { out.omoved[3 * ray] = R.x; out.omoved[3 * ray + 1] = R.y; out.omoved[3 * ray + 2] = R.z; }
        fl |= WF_NORAY;
    }
    p.U(U_FLAGS, s) = fl;
}

// ---- SF, part 2: put ray number `ray` into slot s
template <int SLOTS>
HD void wave_fetch(const WavePool<SLOTS>& p, int s, long long ray, const double* __restrict__ o, const double* __restrict__ d,
                   const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid) {
    p.D(D_OX, s) = o[3 * ray]; p.D(D_OY, s) = o[3 * ray + 1]; p.D(D_OZ, s) = o[3 * ray + 2];
    p.D(D_DX, s) = d[3 * ray]; p.D(D_DY, s) = d[3 * ray + 1]; p.D(D_DZ, s) = d[3 * ray + 2];
    p.U(U_OR1, s) = (uint32_t)(o1a ? o1a[ray] : -1);
    p.U(U_OR2, s) = (uint32_t)(o2a ? o2a[ray] : -1);
    p.U(U_RAY, s) = (uint32_t)ray;
    p.U(U_FLAGS, s) = (rid && rid[ray] == 0) ? WF_BLIND : 0u;   // bounce 0, running
}

// ---- SF, part 3: DDA set-up of the ray in slot s   Voxel_Grid.cs:357-422.  Returns the slot's new tag.
template <bool COUNT, int SLOTS>
HD uint32_t wave_setup(const VGrid& g, const uint32_t* occ, bool occ_smem, const WavePool<SLOTS>& p, int s, CntT<COUNT>& c) {   // occ = g.occp or its shared-memory copy
    uint32_t fl = p.U(U_FLAGS, s) & (WF_BLIND | (0xffffu << WF_BOUNCE_SHIFT));
    Ray3 R = { p.D(D_OX, s), p.D(D_OY, s), p.D(D_OZ, s), p.D(D_DX, s), p.D(D_DY, s), p.D(D_DZ, s) };
    double t_start = 0;
    uint32_t fin = FIN_RUN, lpos = 0, lend = 0;
    int X = floor_to_int((R.x - g.ominx) / g.vdx);
    int Y = floor_to_int((R.y - g.ominy) / g.vdy);
    int Z = floor_to_int((R.z - g.ominz) / g.vdz);
    if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) {
        if (!obox_enter(g, R, t_start)) fin = FIN_MISS;
        else {
            p.D(D_OX, s) = R.x; p.D(D_OY, s) = R.y; p.D(D_OZ, s) = R.z;   // the caller's Ray is moved (AABB_Main.cs:255-257)
            X = floor_to_int((R.x - g.ominx + R.dx * 1E-6) / g.vdx);
            Y = floor_to_int((R.y - g.ominy + R.dy * 1E-6) / g.vdy);
            Z = floor_to_int((R.z - g.ominz + R.dz * 1E-6) / g.vdz);
            if (X < 0 || X >= g.nx || Y < 0 || Y >= g.ny || Z < 0 || Z >= g.nz) fin = FIN_FAULT;
        }
    }
    p.D(D_TMIN, s) = DBL_MAX; p.D(D_TSTART, s) = t_start;
    p.U(U_PID, s) = 0xffffffffu; p.U(U_LAST, s) = 0xffffffffu;
    if (fin == FIN_RUN) {
        const bool nx_ = R.dx < 0, ny_ = R.dy < 0, nz_ = R.dz < 0;
        fl |= (nx_ ? WF_NEGX : 0u) | (ny_ ? WF_NEGY : 0u) | (nz_ ? WF_NEGZ : 0u);
        p.D(D_TMX, s) = ((nx_ ? vox_min(X, g.vdx, g.ominx) : vox_max(X, g.vdx, g.ominx)) - R.x) / R.dx;
        p.D(D_TMY, s) = ((ny_ ? vox_min(Y, g.vdy, g.ominy) : vox_max(Y, g.vdy, g.ominy)) - R.y) / R.dy;
        p.D(D_TMZ, s) = ((nz_ ? vox_min(Z, g.vdz, g.ominz) : vox_max(Z, g.vdz, g.ominz)) - R.z) / R.dz;
        p.D(D_TDX, s) = g.vdx / R.dx * (nx_ ? -1.0 : 1.0);
        p.D(D_TDY, s) = g.vdy / R.dy * (ny_ ? -1.0 : 1.0);
        p.D(D_TDZ, s) = g.vdz / R.dz * (nz_ ? -1.0 : 1.0);
        const uint32_t cp = wave_cp(g, X, Y, Z);
        p.U(U_XYZ, s) = cp;
        c.cell();
        if (!(fl & WF_BLIND) && wave_bit(occ, occ_smem, cp)) {
            const uint32_t ci = ((uint32_t)X * (uint32_t)g.ny + (uint32_t)Y) * (uint32_t)g.nz + (uint32_t)Z;
            const uint2 h = hare_ldg(g.cells + ci); lpos = h.x; lend = h.x + h.y;
        }
    }
    fl |= fin << WF_FIN_SHIFT;
    p.U(U_FLAGS, s) = fl; p.U(U_LPOS, s) = lpos; p.U(U_LEND, s) = lend;
    return wave_tag(fl, lpos, lend);
}

// Which axis the 3D-DDA steps along (Voxel_Grid.cs:504-550, quirk Q4): X only if tMaxX is STRICTLY below both others, Y only if
// tMaxY < tMaxZ (given that X lost), otherwise Z -- so an exact tie goes to the later axis.  Branch-free for a converged warp.
HD int dda_axis(double tMaxX, double tMaxY, double tMaxZ) {
    const bool xy = tMaxX < tMaxY, xz = tMaxX < tMaxZ, yz = tMaxY < tMaxZ;
    return (xy & xz) ? 0 : (((!xy) & yz) ? 1 : 2);
}

// ---- W: the slot's list is exhausted -> accept the carried candidate or step the 3D-DDA (<= W_MAX voxels)
template <bool COUNT, int SLOTS, int W_MAX>
HD uint32_t wave_walk(const VGrid& g, const WaveGeom& w, const uint32_t* occ, bool occ_smem, const WavePool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(U_FLAGS, s);
    const Ray3 R = { p.D(D_OX, s), p.D(D_OY, s), p.D(D_OZ, s), p.D(D_DX, s), p.D(D_DY, s), p.D(D_DZ, s) };
    double tMaxX = p.D(D_TMX, s), tMaxY = p.D(D_TMY, s), tMaxZ = p.D(D_TMZ, s);
    const double tDeltaX = p.D(D_TDX, s), tDeltaY = p.D(D_TDY, s), tDeltaZ = p.D(D_TDZ, s);
    uint32_t cp = p.U(U_XYZ, s);
    const int sX = (fl & WF_NEGX) ? -(int)w.d1 : (int)w.d1, sY = (fl & WF_NEGY) ? -(int)w.pz : (int)w.pz, sZ = (fl & WF_NEGZ) ? -1 : 1;
    const bool have = (fl & WF_HAVE) != 0, blind = (fl & WF_BLIND) != 0;
    double bx = 0, by = 0, bz = 0;
    if (have) { const double tmin = p.D(D_TMIN, s); bx = R.x + R.dx * tmin; by = R.y + R.dy * tmin; bz = R.z + R.dz * tmin; }
    uint32_t fin = FIN_RUN, lpos = 0, lend = 0;
    bool stop = false;   // the walk ended on a set bit: an occupied voxel or the border
#pragma unroll 1
    for (int guard = 0; guard < W_MAX; ++guard) {
        // list exhausted: Voxels[X,Y,Z].IsPointInBox(candidate)?   Voxel_Grid.cs:496-500
        if (have) {
            int X, Y, Z;
            wave_decode(w, cp, X, Y, Z);
            const bool in = !(bx < vox_min(X, g.vdx, g.ominx)) & !(by < vox_min(Y, g.vdy, g.ominy)) & !(bz < vox_min(Z, g.vdz, g.ominz)) &
                            !(bx > vox_max(X, g.vdx, g.ominx)) & !(by > vox_max(Y, g.vdy, g.ominy)) & !(bz > vox_max(Z, g.vdz, g.ominz));
            if (in) { fin = FIN_HIT; break; }
        }
        // next voxel   Voxel_Grid.cs:504-550: X only if strictly below both, Y only if below Z, else Z
        {
            const int ax = dda_axis(tMaxX, tMaxY, tMaxZ);
            const bool goX = ax == 0, goY = ax == 1, goZ = ax == 2;
            const double nX = tMaxX + tDeltaX, nY = tMaxY + tDeltaY, nZ = tMaxZ + tDeltaZ;
            tMaxX = goX ? nX : tMaxX; tMaxY = goY ? nY : tMaxY; tMaxZ = goZ ? nZ : tMaxZ;
            cp += (uint32_t)(goX ? sX : (goY ? sY : sZ));
        }
        c.cell();
        // a set bit ends the walk: an occupied voxel (its list header is fetched once, after the loop) or the border.  A blind
        // ray (Ray_ID == 0) ignores lists, so only the border stops it.
        if (wave_bit(occ, occ_smem, cp)) {
            if (!blind) { stop = true; break; }
            int X, Y, Z;
            wave_decode(w, cp, X, Y, Z);
            if ((unsigned)X >= (unsigned)g.nx || (unsigned)Y >= (unsigned)g.ny || (unsigned)Z >= (unsigned)g.nz) { stop = true; break; }
        }
    }
    if (stop) {
        int X, Y, Z;
        wave_decode(w, cp, X, Y, Z);
        if ((unsigned)X >= (unsigned)g.nx || (unsigned)Y >= (unsigned)g.ny || (unsigned)Z >= (unsigned)g.nz) {
            fin = FIN_MISS;                       // left the grid: a miss even if a candidate is held (Voxel_Grid.cs:509-513)
            if (COUNT) --c.cells;                 // the reference never enters that voxel
        } else {
            const uint32_t ci = ((uint32_t)X * (uint32_t)g.ny + (uint32_t)Y) * (uint32_t)g.nz + (uint32_t)Z;
            const uint2 h = hare_ldg(g.cells + ci); lpos = h.x; lend = h.x + h.y;
        }
    }
    fl |= fin << WF_FIN_SHIFT;
    if (fin == FIN_RUN) {
        p.D(D_TMX, s) = tMaxX; p.D(D_TMY, s) = tMaxY; p.D(D_TMZ, s) = tMaxZ;
        p.U(U_XYZ, s) = cp;
        p.U(U_LPOS, s) = lpos; p.U(U_LEND, s) = lend;
    } else {
        p.U(U_FLAGS, s) = fl;
    }
    return wave_tag(fl, lpos, lend);
}

// ---- C: cull the next (up to) four list entries; the survivors wait in the slot for T
template <bool COUNT, int SLOTS>
HD uint32_t wave_cull(const VGrid& g, const WavePool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t lpos = p.U(U_LPOS, s);
    const uint32_t lend = p.U(U_LEND, s);
    const uint32_t n = (lend - lpos) < 4u ? (lend - lpos) : 4u;
    if (COUNT) c.entries += n;
    // the cull works in FP32 in a frame local to the voxel: its ray point is where the ray LEAVES the current voxel,
    // min(tMax), at most a voxel diagonal from anything listed here
    const double dx = p.D(D_DX, s), dy = p.D(D_DY, s), dz = p.D(D_DZ, s);
    const double te = fmin(fmin(p.D(D_TMX, s), p.D(D_TMY, s)), p.D(D_TMZ, s));
    const float fpx = (float)fma(dx, te, p.D(D_OX, s)), fpy = (float)fma(dy, te, p.D(D_OY, s)), fpz = (float)fma(dz, te, p.D(D_OZ, s));
    const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
    const int or1 = (int)p.U(U_OR1, s), or2 = (int)p.U(U_OR2, s), pid = (int)p.U(U_PID, s);
    const uint32_t last = p.U(U_LAST, s);
    // poly_origin skip (Voxel_Grid.cs:477); a polygon already tested for this ray cannot change the result
    auto fresh = [&](uint32_t i) { return !((int)i == or1 || (int)i == or2 || i == last || (int)i == pid); };
    // list entries carry their polygon's padded bounding box and its id: one contiguous 32 bytes per entry, no id -> box dependent load
    const float4* e = g.lbox + 2 * (size_t)lpos;
    const float4 l0 = hare_ldg(e), h0 = hare_ldg(e + 1);
    const float4 l1 = (n > 1) ? hare_ldg(e + 2) : l0, h1 = (n > 1) ? hare_ldg(e + 3) : h0;
    const float4 l2 = (n > 2) ? hare_ldg(e + 4) : l0, h2 = (n > 2) ? hare_ldg(e + 5) : h0;
    const float4 l3 = (n > 3) ? hare_ldg(e + 6) : l0, h3 = (n > 3) ? hare_ldg(e + 7) : h0;
    const uint32_t bid0 = hare_f2u(l0.w), bid1 = hare_f2u(l1.w), bid2 = hare_f2u(l2.w), bid3 = hare_f2u(l3.w);
    const float ix = cull_rcp(fdx), iy = cull_rcp(fdy), iz = cull_rcp(fdz);
    const float pxi = fpx * ix, pyi = fpy * iy, pzi = fpz * iz;
    const uint32_t bmask = ((fresh(bid0) && !cull_box(l0, h0, pxi, pyi, pzi, ix, iy, iz)) ? 1u : 0u) |
                           ((n > 1 && fresh(bid1) && !cull_box(l1, h1, pxi, pyi, pzi, ix, iy, iz)) ? 2u : 0u) |
                           ((n > 2 && fresh(bid2) && !cull_box(l2, h2, pxi, pyi, pzi, ix, iy, iz)) ? 4u : 0u) |
                           ((n > 3 && fresh(bid3) && !cull_box(l3, h3, pxi, pyi, pzi, ix, iy, iz)) ? 8u : 0u);
    if (bmask) { p.U(U_FLAGS, s) |= bmask << WF_BMASK_SHIFT; return PH_T; }   // lpos stays on the batch until T has consumed it
    lpos += n;
    p.U(U_LPOS, s) = lpos;
    return lpos < lend ? PH_C : PH_W;
}

// ---- T: one exact FP64 test of the lowest surviving entry
template <bool COUNT, int SLOTS>
HD uint32_t wave_test(const VGrid& g, const PolyRec* __restrict__ polys, const WavePool<SLOTS>& p, int s, CntT<COUNT>& c) {
    uint32_t fl = p.U(U_FLAGS, s);
    const uint32_t bmask = (fl & WF_BMASK_MASK) >> WF_BMASK_SHIFT;
    const int k = (bmask & 1u) ? 0 : ((bmask & 2u) ? 1 : ((bmask & 4u) ? 2 : 3));   // lowest survivor first
    uint32_t lpos = p.U(U_LPOS, s);
    const uint32_t lend = p.U(U_LEND, s);
    const uint32_t pend = hare_ldg(g.cell_poly + lpos + k);
    fl &= ~((1u << k) << WF_BMASK_SHIFT);
    if (!(fl & WF_BMASK_MASK)) { lpos += (lend - lpos) < 4u ? (lend - lpos) : 4u; p.U(U_LPOS, s) = lpos; }
    c.test();
    const Ray3 R = { p.D(D_OX, s), p.D(D_OY, s), p.D(D_OZ, s), p.D(D_DX, s), p.D(D_DY, s), p.D(D_DZ, s) };
    double P[16], t = 0;
    load_poly(polys, pend, P);
    // Polygon.Ray_Side picks the winding (Hare_Geometry_Polygons.cs:601-606, 637-660, 784-823):
    //   side ? (P0,P1,P2) then (P2,P3,P0) : (P2,P1,P0) then (P0,P3,P2)
    const bool side = !(dot3(R.dx, R.dy, R.dz, P[12], P[13], P[14]) < 0);
    const double ax = side ? P[0] : P[6], ay = side ? P[1] : P[7], az = side ? P[2] : P[8];
    const double cx = side ? P[6] : P[0], cy = side ? P[7] : P[1], cz = side ? P[8] : P[2];
    bool hit = ray_x_tri_fast1(R, ax, ay, az, P[3], P[4], P[5], cx, cy, cz, t);
    if (!hit && P[15] == 4.0) hit = ray_x_tri_fast1(R, cx, cy, cz, P[9], P[10], P[11], ax, ay, az, t);
    p.U(U_LAST, s) = pend;
    if (hit && t > 0.0000000001 && t < p.D(D_TMIN, s)) { p.D(D_TMIN, s) = t; p.U(U_PID, s) = pend; fl |= WF_HAVE; }
    p.U(U_FLAGS, s) = fl;
    return wave_tag(fl, lpos, lend);
}

// ---- ray supply of a warp (shared by the three traversal kernels) ------------------------------------------------------------
// The batch is handed out on demand, in blocks of consecutive rays (of the coherence order, ray_bin.cuh) claimed from a device
// counter.  A persistent kernel lasts as long as its slowest warp, and the work behind a share of the batch varies (4 % per warp
// over 12.5 M binned rays, CPU replay): with a fixed interleave of the groups over the warps the Octree ran 728 Mrays/s on C3, with
// this supply 829 (Voxel_Grid 888 -> 982, KDTree 271 -> 302; a mixed scheme -- fixed for the first 7/8 or 3/4 of the batch --
// landed in between: 794 / 800; blocks of 32 / 64 / 128 / 512 rays: 802 / 829 / 844 / 834 on 100 M rays, 713 / 721 / 708 on 12.5 M).
// A warp always holds one block in reserve, claimed by lane 0 when the previous one was opened and read only when that one is used
// up, many trips later: the atomic's round trip is never waited for.  The first block of a warp is its own number (no atomic, and
// a small batch still spreads over all the warps); the counter hands out the blocks after those.
#ifndef HARE_FEED_BLOCK
#define HARE_FEED_BLOCK 64
#endif
struct RayFeedArgs {
    unsigned long long* ctr;        // rays claimed beyond the warps' first blocks; zero before the launch
    long long first;                // first ray handed out through the counter = warps of the launch x block
    int block;                      // rays per block, >= 32: one trip takes at most 32 rays, i.e. touches at most two blocks
};
struct RayFeed {                    // warp-uniform, except b1: lane 0's until it is broadcast at the next SF trip
    long long b0, b1;               // ray number of the first ray of the open block / of the block in reserve
    int used;                       // rays taken from the open block (< block)
};
// rays per block for a batch of N rays on tw warps: HARE_FEED_BLOCK; 32 when that would leave fewer than eight blocks per warp;
// `huge` (a kernel's choice, >= HARE_FEED_BLOCK) when there are 256 or more
HD int feed_block_for(long long N, long long tw, int huge = HARE_FEED_BLOCK) {
    return N >= tw * HARE_FEED_BLOCK * 256 ? huge : (N >= tw * HARE_FEED_BLOCK * 8 ? HARE_FEED_BLOCK : 32);
}
// claim the next block (one lane per warp)
HD long long feed_claim(const RayFeedArgs& A) {
#if defined(__CUDA_ARCH__)
    return A.first + (long long)atomicAdd(A.ctr, (unsigned long long)A.block);
#else
    const unsigned long long old = *A.ctr; *A.ctr = old + (unsigned long long)A.block; return A.first + (long long)old;
#endif
}
// ray number of the rank-th ray taken in this trip; b1 = the reserve block's first ray (broadcast)
HD long long feed_ray(const RayFeed& f, const RayFeedArgs& A, long long b1, int rank) {
    const int off = f.used + rank;
    return off < A.block ? f.b0 + off : b1 + (off - A.block);
}
// `need` rays were taken; true when the open block is used up: the reserve has been opened and a new one must be claimed into f.b1
HD bool feed_advance(RayFeed& f, const RayFeedArgs& A, int need, long long b1) {
    f.used += need;
    if (f.used < A.block) return false;
    f.used -= A.block; f.b0 = b1;
    return true;
}

#if defined(__CUDACC__)

#ifndef HARE_WAVE_WARPS
#define HARE_WAVE_WARPS 20
#endif

template <bool CHAIN, bool COUNT, bool OCC_SMEM, int SLOTS, int W_MAX>
__global__ void __launch_bounds__(HARE_WAVE_WARPS * 32, 1)
vg_wave_kernel(const VGrid g, const PolyRec* __restrict__ polys,
               const double* __restrict__ o, const double* __restrict__ d,
               const int32_t* __restrict__ o1a, const int32_t* __restrict__ o2a, const int32_t* __restrict__ rid,
               long long N, int order, const uint32_t* __restrict__ perm /* ray order of ray_bin.cuh, or null */, const RayFeedArgs feed, const WalkOut out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    uint32_t* s_occ = reinterpret_cast<uint32_t*>(s_raw);
    uint32_t occ_words = 0;
    if (OCC_SMEM) {
        occ_words = (((uint32_t)g.nx + 2u) * ((uint32_t)g.ny + 2u) * ((uint32_t)g.nz + 2u) + 31u) >> 5;
        for (uint32_t w = threadIdx.x; w < occ_words; w += blockDim.x) s_occ[w] = __ldg(g.occp + w);
        occ_words = (occ_words + 3u) & ~3u;
    }
    const uint32_t* occ = OCC_SMEM ? s_occ : g.occp;
    const WaveGeom wg = wave_geom(g);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WavePool<SLOTS> p;
    p.bind(s_raw + (size_t)occ_words * 4 + (size_t)warp * WavePool<SLOTS>::STRIDE);
    constexpr int GROUPS = (SLOTS + 31) / 32;
#pragma unroll
    for (int k = 0; k < GROUPS; ++k) {
        const int s = k * 32 + lane;
        if (s < SLOTS) { p.U(U_FLAGS, s) = WF_NORAY; p.U(U_LPOS, s) = 0; p.U(U_LEND, s) = 0; p.tag[s] = (uint8_t)PH_SF; }
    }
    __syncthreads();

    CntT<COUNT> c;
    unsigned int shots = 0;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    RayFeed f = { gw * feed.block, 0, 0 };   // see RayFeed in vg_wave.cuh
    if (lane == 0) f.b1 = feed_claim(feed);
    const unsigned lt = (1u << lane) - 1u;

    while (true) {
        // 1. count the slots per phase: one REDUX over byte-packed per-lane counts (a byte holds <= SLOTS <= 255)
        uint32_t tg[GROUPS], packed = 0;
#pragma unroll
        for (int k = 0; k < GROUPS; ++k) {
            const int s = k * 32 + lane;
            tg[k] = (s < SLOTS) ? p.tag[s] : (uint32_t)PH_DONE;
            packed += (tg[k] < (uint32_t)PH_COUNT) ? (1u << (8 * tg[k])) : 0u;
        }
        packed = __reduce_add_sync(0xffffffffu, packed);
        const int n[PH_COUNT] = { (int)(packed & 255u), (int)((packed >> 8) & 255u), (int)((packed >> 16) & 255u), (int)(packed >> 24) };
        // 2. pick one
        const int ph = wave_pick(n);
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
        uint32_t nt = PH_DONE;
        if (ph == PH_T) {
            if (act) nt = wave_test<COUNT, SLOTS>(g, polys, p, s, c);
        } else if (ph == PH_C) {
            if (act) nt = wave_cull<COUNT, SLOTS>(g, p, s, c);
        } else if (ph == PH_W) {
            if (act) nt = wave_walk<COUNT, SLOTS, W_MAX>(g, wg, occ, OCC_SMEM, p, s, c);
        } else {
            if (act) wave_finish<CHAIN, COUNT, SLOTS>(polys, p, s, order, out, shots, c);
            const bool noray = act && (p.U(U_FLAGS, s) & WF_NORAY);
            const unsigned want = __ballot_sync(0xffffffffu, noray);
            bool ready = act;
            const long long b1 = __shfl_sync(0xffffffffu, f.b1, 0);
            if (noray) {
                const long long ray = feed_ray(f, feed, b1, __popc(want & lt));
                if (ray < N) wave_fetch<SLOTS>(p, s, perm ? (long long)__ldg(perm + ray) : ray, o, d, o1a, o2a, rid);
                else ready = false;
            }
            if (feed_advance(f, feed, __popc(want), b1) && lane == 0) f.b1 = feed_claim(feed);
            if (ready) nt = wave_setup<COUNT, SLOTS>(g, occ, OCC_SMEM, p, s, c);
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
