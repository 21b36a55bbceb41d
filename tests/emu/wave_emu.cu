// wave_emu.cu -- TEST INFRASTRUCTURE: replays the wavefront scheduler of hare_b200/csrc/vg_wave.cuh on the CPU.
//
// The per-slot phase functions of vg_wave.cuh (wave_finish / wave_fetch / wave_setup / wave_walk / wave_cull /
// wave_test) and its policy (wave_pick, wave_tag, the RayFeed) are `__host__ __device__`; this file compiles
// them for the host and drives them with a sequential copy of the kernel's trip loop (the simulated warps take
// turns trip by trip, 32 "lanes" run one after the other).  It lets the CPU test-suite check the state machine against the
// oracle without a GPU, and reports how many lanes each phase execution would keep busy.
// Nothing here is shipped or measured; it never launches a kernel.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../hare_b200/csrc/kernels.cuh"
#include "../../hare_b200/csrc/vg_wave.cuh"
#include "../../hare_b200/csrc/pack.hpp"

using namespace hare;

namespace {

struct Stats { double exec[4] = { 0, 0, 0, 0 }, lanes[4] = { 0, 0, 0, 0 }, wsteps_warp = 0, wsteps_lane = 0, trips = 0, whave_exec = 0, whave_lane = 0; };

template <bool CHAIN, int SLOTS, int W_MAX>
void run(const VGrid& g, const PolyRec* polys, const double* o, const double* d, const int32_t* o1a, const int32_t* o2a,
         const int32_t* rid, long long N, int order, const WalkOut& out, int tw, int wexit, Stats& st, unsigned long long* counters) {
    // The simulated warps take turns, one trip each: their pools are in flight together and their claims on the launch's counter
    // interleave, as on the device (where the order is arbitrary -- results must not depend on it).
    struct Warp {
        std::vector<unsigned char> mem;
        WavePool<SLOTS> p; RayFeed f; unsigned int shots = 0; bool done = false;
    };
    CntT<true> c;
    const WaveGeom wg = wave_geom(g);
    unsigned long long total = 0;
    unsigned long long feed_ctr = 0;
    const RayFeedArgs feed = { &feed_ctr, tw * feed_block_for(N, tw), feed_block_for(N, tw) };
    std::vector<Warp> warps((size_t)tw);
    for (long long gw = 0; gw < tw; ++gw) {
        Warp& w = warps[(size_t)gw];
        w.mem.resize(WavePool<SLOTS>::STRIDE + 64);
        w.p.bind(w.mem.data());
        for (int s = 0; s < SLOTS; ++s) { w.p.U(U_FLAGS, s) = WF_NORAY; w.p.U(U_LPOS, s) = 0; w.p.U(U_LEND, s) = 0; w.p.tag[s] = (uint8_t)PH_SF; }
        w.f = RayFeed{ gw * feed.block, 0, 0 };
        w.f.b1 = feed_claim(feed);
    }
    for (long long live = tw; live > 0;) {
        for (long long gw = 0; gw < tw; ++gw) {
            Warp& w = warps[(size_t)gw];
            if (w.done) continue;
            WavePool<SLOTS>& p = w.p; RayFeed& f = w.f; unsigned int& shots = w.shots;
            int n[PH_COUNT] = { 0, 0, 0, 0 };
            for (int s = 0; s < SLOTS; ++s) if (p.tag[s] < PH_COUNT) ++n[p.tag[s]];
            const int ph = wave_pick(n);
            if (ph < 0) { w.done = true; --live; total += shots; continue; }
            // the kernel ranks group 0 (slots 0..31) before group 1, lane order inside a group = slot order
            int sel[32], cnt = 0;
            for (int s = 0; s < SLOTS && cnt < 32; ++s) if (p.tag[s] == ph) sel[cnt++] = s;
            st.exec[ph] += 1; st.lanes[ph] += cnt; st.trips += 1;
            uint32_t nt[32];
            if (ph == PH_T) {
                for (int l = 0; l < cnt; ++l) nt[l] = wave_test<true, SLOTS>(g, polys, p, sel[l], c);
            } else if (ph == PH_C) {
                for (int l = 0; l < cnt; ++l) nt[l] = wave_cull<true, SLOTS>(g, p, sel[l], c);
            } else if (ph == PH_W) {
                if (wexit > 0) {
                    // what-if: leave the walk loop as soon as fewer than `wexit` lanes are still stepping (lockstep replay, one
                    // voxel step per lane and round; a lane's state survives in its slot, so W_MAX = 1 calls compose exactly)
                    bool live[32];
                    for (int l = 0; l < cnt; ++l) { live[l] = true; nt[l] = PH_W; }
                    for (int r = 0; r < W_MAX; ++r) {
                        int act = 0;
                        for (int l = 0; l < cnt; ++l) if (live[l]) ++act;
                        if (act == 0 || (r > 0 && act < wexit)) break;
                        st.wsteps_warp += 1; st.wsteps_lane += act;
                        for (int l = 0; l < cnt; ++l) if (live[l]) { nt[l] = wave_walk<true, SLOTS, 1>(g, wg, g.occp, false, p, sel[l], c); live[l] = nt[l] == PH_W; }
                    }
                } else {
                unsigned mx = 0;
                for (int l = 0; l < cnt; ++l) {
                    const unsigned before = c.cells;
                    const bool had = (p.U(U_FLAGS, sel[l]) & WF_HAVE) != 0;
                    nt[l] = wave_walk<true, SLOTS, W_MAX>(g, wg, g.occp, false, p, sel[l], c);
                    unsigned steps = c.cells - before;
                    if (steps == 0 || (had && nt[l] == PH_SF)) steps += 1;   // an accept / exit iteration enters no cell
                    st.wsteps_lane += steps; if (steps > mx) mx = steps;
                }
                st.wsteps_warp += mx;
                }
            } else {
                for (int l = 0; l < cnt; ++l) wave_finish<CHAIN, true, SLOTS>(polys, p, sel[l], order, out, shots, c);
                int rank = 0;
                for (int l = 0; l < cnt; ++l) {
                    bool ready = true;
                    if (p.U(U_FLAGS, sel[l]) & WF_NORAY) {
                        const long long ray = feed_ray(f, feed, f.b1, rank);
                        ++rank;
                        if (ray < N) wave_fetch<SLOTS>(p, sel[l], ray, o, d, o1a, o2a, rid);
                        else ready = false;
                    }
                    nt[l] = ready ? wave_setup<true, SLOTS>(g, g.occp, false, p, sel[l], c) : (uint32_t)PH_DONE;
                }
                if (feed_advance(f, feed, rank, f.b1)) f.b1 = feed_claim(feed);
            }
            for (int l = 0; l < cnt; ++l) p.tag[sel[l]] = (uint8_t)nt[l];
        }
    }
    if (CHAIN && out.total_shots) *out.total_shots = total;
    if (counters) { counters[0] = c.cells; counters[1] = c.entries; counters[2] = c.tests; counters[3] = c.hits; }
}

}  // namespace

// Host arrays in, host arrays out.  (slots, wmax) from the RUN list below; wexit > 0 = what-if replay of an adaptive walk exit.  stats: 14 doubles
// (exec[4], lanes[4] in phase order SF, W, C, T; warp-level W iterations; lane-level W iterations; trips; 3 spare).
extern "C" int wave_emu(const double* verts, const double* normals, const int32_t* vcount, int64_t P,
                        const double obox[6], const int32_t ct[3], const uint32_t* cell_offset, const uint32_t* cell_poly,
                        const double* o, const double* d, const int32_t* o1, const int32_t* o2, const int32_t* rid, int64_t N,
                        int chain, int order,
                        double* t, double* xyz, int32_t* pid, double* uv, double* omoved,
                        int32_t* ev_pid, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots, unsigned long long* total_shots,
                        int slots, int wmax, int n_warps, int wexit, double* stats, unsigned long long* counters,
                        double* ev_xyz, double* ev_uv /* chain: per-bounce X_Point / u, v rows, optional */) {
    std::vector<PolyRec> recs((size_t)P);
    for (int64_t i = 0; i < P; ++i) {
        for (int k = 0; k < 12; ++k) recs[i].v[k] = verts[12 * i + k];
        if (vcount[i] == 3) for (int a = 0; a < 3; ++a) recs[i].v[9 + a] = verts[12 * i + 6 + a];
        for (int a = 0; a < 3; ++a) recs[i].v[12 + a] = normals[3 * i + a];
        recs[i].v[15] = (double)vcount[i];
    }
    const int64_t ncells = (int64_t)ct[0] * ct[1] * ct[2];
    std::vector<uint2> cells((size_t)ncells);
    std::vector<uint32_t> occ((size_t)(ncells + 31) / 32 + 1, 0u);
    for (int64_t i = 0; i < ncells; ++i) {
        cells[i] = make_uint2(cell_offset[i], cell_offset[i + 1] - cell_offset[i]);
        if (cells[i].y) occ[i >> 5] |= 1u << (i & 31);
    }
    VGrid g;
    g.ominx = obox[0]; g.ominy = obox[1]; g.ominz = obox[2]; g.omaxx = obox[3]; g.omaxy = obox[4]; g.omaxz = obox[5];
    g.vdx = (obox[3] - obox[0]) / ct[0]; g.vdy = (obox[4] - obox[1]) / ct[1]; g.vdz = (obox[5] - obox[2]) / ct[2];
    g.nx = ct[0]; g.ny = ct[1]; g.nz = ct[2];
    // border-padded occupancy bitmap, as vg_pad_occupancy makes it
    const int64_t px = ct[0] + 2, py = ct[1] + 2, pz = ct[2] + 2;
    std::vector<uint32_t> occp((size_t)(px * py * pz + 31) / 32 + 1, 0u);
    for (int64_t xp = 0; xp < px; ++xp) for (int64_t yp = 0; yp < py; ++yp) for (int64_t zp = 0; zp < pz; ++zp) {
        const int64_t cp = (xp * py + yp) * pz + zp;
        bool bit = xp == 0 || xp == px - 1 || yp == 0 || yp == py - 1 || zp == 0 || zp == pz - 1;
        if (!bit) { const int64_t ci = ((xp - 1) * ct[1] + (yp - 1)) * ct[2] + (zp - 1); bit = cells[ci].y != 0; }
        if (bit) occp[cp >> 5] |= 1u << (cp & 31);
    }
    g.cells = cells.data(); g.cell_poly = cell_poly; g.occ = occ.data(); g.occp = occp.data();
    // per-entry padded boxes with the id in lo.w, as vg_gather_list_box makes them
    std::vector<float4> lbox(2 * (size_t)cell_offset[ncells] + 2);
    for (uint32_t k = 0; k < cell_offset[ncells]; ++k) {
        const int64_t i = cell_poly[k];
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            double l = verts[12 * i + a], h = l;
            for (int q = 1; q < vcount[i]; ++q) { l = std::fmin(l, verts[12 * i + 3 * q + a]); h = std::fmax(h, verts[12 * i + 3 * q + a]); }
            const double pad = hare_box_pad(l, h);
            lo[a] = (float)(l - pad); while ((double)lo[a] > l - pad) lo[a] = std::nextafter(lo[a], -INFINITY);
            hi[a] = (float)(h + pad); while ((double)hi[a] < h + pad) hi[a] = std::nextafter(hi[a], INFINITY);
        }
        lbox[2 * k] = make_float4(lo[0], lo[1], lo[2], hare_u2f((uint32_t)i)); lbox[2 * k + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    g.lbox = lbox.data();
    WalkOut out = { t, xyz, pid, uv, omoved, ev_pid, ev_t, fin_o, fin_d, nshots, total_shots, nullptr, ev_xyz, ev_uv };
    Stats st;
#define RUN(S, W) if (slots == S && wmax == W) { if (chain) run<true, S, W>(g, recs.data(), o, d, o1, o2, rid, N, order, out, n_warps, wexit, st, counters); \
                                                 else run<false, S, W>(g, recs.data(), o, d, o1, o2, rid, N, order, out, n_warps, wexit, st, counters); ok = 1; }
    int ok = 0;
    RUN(40, 4) RUN(48, 4) RUN(64, 4) RUN(96, 4) RUN(64, 2) RUN(64, 8) RUN(48, 8) RUN(48, 2) RUN(32, 4) RUN(64, 16) RUN(96, 8)
#undef RUN
    if (!ok) return -1;
    if (stats) {
        for (int k = 0; k < 4; ++k) { stats[k] = st.exec[k]; stats[4 + k] = st.lanes[k]; }
        stats[8] = st.wsteps_warp; stats[9] = st.wsteps_lane; stats[10] = st.trips; stats[11] = st.whave_exec; stats[12] = st.whave_lane;
    }
    return 0;
}

// ---- cull_box exactly as the kernels use it, for the property test of its conservativeness (tests/test_wave_emu.py) ----------
// verts: n x 4 x 3 (a triangle repeats vertex 2), o/d: n x 3 rays, t_frame: n ray parameters of the FP32 frame point.
// out[i] = 1 when the polygon's padded box is rejected for ray i.
extern "C" void emu_cull_box(const double* verts, const int32_t* vcount, const double* o, const double* d, const double* t_frame, int64_t n,
                             unsigned char* out) {
    for (int64_t i = 0; i < n; ++i) {
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            double l = verts[12 * i + a], h = l;
            for (int q = 1; q < vcount[i]; ++q) { l = std::fmin(l, verts[12 * i + 3 * q + a]); h = std::fmax(h, verts[12 * i + 3 * q + a]); }
            const double pad = hare_box_pad(l, h);
            lo[a] = (float)(l - pad); while ((double)lo[a] > l - pad) lo[a] = std::nextafter(lo[a], -INFINITY);
            hi[a] = (float)(h + pad); while ((double)hi[a] < h + pad) hi[a] = std::nextafter(hi[a], INFINITY);
        }
        const float fdx = (float)d[3 * i], fdy = (float)d[3 * i + 1], fdz = (float)d[3 * i + 2];
        const float ix = cull_rcp(fdx), iy = cull_rcp(fdy), iz = cull_rcp(fdz);
        const float px = (float)fma(d[3 * i], t_frame[i], o[3 * i]), py = (float)fma(d[3 * i + 1], t_frame[i], o[3 * i + 1]),
                    pz = (float)fma(d[3 * i + 2], t_frame[i], o[3 * i + 2]);
        out[i] = cull_box(make_float4(lo[0], lo[1], lo[2], 0.f), make_float4(hi[0], hi[1], hi[2], 0.f), px * ix, py * iy, pz * iz, ix, iy, iz) ? 1 : 0;
    }
}

// ---- the DDA step selection exactly as the kernel uses it (quirk Q4 unit test) ----------
extern "C" int emu_dda_axis(double tx, double ty, double tz) { return dda_axis(tx, ty, tz); }

// ---- the coherence pre-pass key (ray_bin.cuh), for its range / degenerate-input test ----------
#include "../../hare_b200/csrc/ray_bin.cuh"
extern "C" void emu_ray_bin_keys(const double* o, const double* d, int64_t n, const double* minmax, uint32_t* keys, uint32_t* buckets) {
    RayBinGeom g;
    g.ox = (float)minmax[0]; g.oy = (float)minmax[1]; g.oz = (float)minmax[2];
    g.sx = 4.0f / (float)(minmax[3] - minmax[0]); g.sy = 4.0f / (float)(minmax[4] - minmax[1]); g.sz = 4.0f / (float)(minmax[5] - minmax[2]);
    g.dirbits = ray_bin_dirbits(n);
    for (int64_t i = 0; i < n; ++i) keys[i] = ray_bin_key(o + 3 * i, d + 3 * i, g);
    *buckets = ray_bin_buckets(g.dirbits);
}

// ---- the ray supply (RayFeed, vg_wave.cuh): tw simulated warps take rays in a random interleaving, 1..32 at a time -------------
// counts[r] = how often ray r was handed out; returns the number of rays the claimed blocks cover.  A warp stops once 64 of its requests in a row
// were answered with ray numbers >= N (in the kernels: every slot of its pool has gone to DONE).
extern "C" long long emu_ray_feed(int64_t N, int64_t tw, uint64_t seed, uint32_t* counts) {
    unsigned long long ctr = 0;
    const int fb = feed_block_for(N, tw, 128);   // (the Octree launcher's choice for huge batches)
    const RayFeedArgs feed = { &ctr, tw * fb, fb };
    std::vector<RayFeed> f((size_t)tw);
    std::vector<int> dry((size_t)tw, 0);
    for (int64_t w = 0; w < tw; ++w) { f[w] = RayFeed{ w * feed.block, 0, 0 }; f[w].b1 = feed_claim(feed); }
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    int64_t live = tw;
    while (live > 0) {
        const int64_t w = (int64_t)(rnd() % (uint64_t)tw);
        if (dry[w] >= 64) continue;
        const int need = 1 + (int)(rnd() % 32);
        for (int r = 0; r < need; ++r) {
            const long long ray = feed_ray(f[w], feed, f[w].b1, r);
            if (ray < 0) return -1;
            if (ray < N) { ++counts[ray]; dry[w] = 0; } else ++dry[w];
        }
        if (feed_advance(f[w], feed, need, f[w].b1)) f[w].b1 = feed_claim(feed);
        if (dry[w] >= 64) --live;
    }
    return feed.first + (long long)ctr;
}

// ---- the chunk schedule of hare_shoot_batch (schedule.hpp) -------------------------------------------------------------------
#include "../../hare_b200/csrc/schedule.hpp"
extern "C" int emu_shoot_schedule(int64_t n, int64_t* sizes, int cap) {
    const std::vector<int64_t> v = shoot_schedule(n);
    for (size_t k = 0; k < v.size() && (int)k < cap; ++k) sizes[k] = v[k];
    return (int)v.size();
}
