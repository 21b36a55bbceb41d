// oct_emu.cu -- TEST INFRASTRUCTURE: replays the wavefront scheduler of hare_b200/csrc/oct_wave.cuh on the CPU.
//
// The per-slot phase functions of oct_wave.cuh (octw_finish / octw_fetch / octw_setup / octw_node / octw_group / octw_cull /
// octw_test) and its policy (oct_pick, oct_tag) are `__host__ __device__`; this file compiles them for the host and drives
// them with a sequential copy of the kernel's trip loop (the simulated warps take turns trip by trip, 32 "lanes" one after the other).
// The device arrays are built by the same pack_octree() the library uses (hare_b200/csrc/pack.hpp).  Nothing here is
// shipped or measured; it never launches a kernel.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../hare_b200/csrc/kernels.cuh"
#include "../../hare_b200/csrc/oct_wave.cuh"
#include "../../hare_b200/csrc/pack.hpp"

using namespace hare;

namespace {

struct Stats { double exec[OP_COUNT] = {}, lanes[OP_COUNT] = {}, trips = 0, nsteps_warp = 0, nsteps_lane = 0; };

template <bool CHAIN, int SLOTS, int N_MAX>
void run(const OctDev& T, int depth, const PolyRec* polys, const double* o, const double* d, const int32_t* o1a, const int32_t* o2a,
         long long N, int order, const WalkOut& out, int tw, Stats& st, unsigned long long* counters, uint32_t* ray_steps) {
    // The simulated warps take turns, one trip each: their pools are in flight together and their claims on the launch's counter
    // interleave, as on the device (where the order is arbitrary -- results must not depend on it).
    struct Warp {
        std::vector<unsigned char> mem; std::vector<double2> fab; std::vector<uint2> fcq;
        OctPool<SLOTS> p; OctFrames F; RayFeed f; unsigned int shots = 0; bool done = false;
    };
    CntT<true> c;
    unsigned long long total = 0;
    unsigned long long feed_ctr = 0;
    const RayFeedArgs feed = { &feed_ctr, tw * feed_block_for(N, tw), feed_block_for(N, tw) };
    std::vector<Warp> warps((size_t)tw);
    for (long long gw = 0; gw < tw; ++gw) {
        Warp& w = warps[(size_t)gw];
        w.mem.resize(OctPool<SLOTS>::STRIDE + 64); w.fab.resize((size_t)SLOTS * (depth + 1)); w.fcq.resize((size_t)SLOTS * (depth + 1));
        w.F = OctFrames{ w.fab.data(), w.fcq.data(), depth + 1 };
        w.p.bind(w.mem.data());
        for (int s = 0; s < SLOTS; ++s) { w.p.U(OU_FLAGS, s) = OFL_NORAY; w.p.U(OU_LPOS, s) = 0; w.p.U(OU_LEND, s) = 0; w.p.U(OU_MASKS, s) = 0; w.p.tag[s] = (uint8_t)OP_SF; }
        w.f = RayFeed{ gw * feed.block, 0, 0 };
        w.f.b1 = feed_claim(feed);
    }
    for (long long live = tw; live > 0;) {
        for (long long gw = 0; gw < tw; ++gw) {
            Warp& w = warps[(size_t)gw];
            if (w.done) continue;
            OctPool<SLOTS>& p = w.p; const OctFrames& F = w.F; RayFeed& f = w.f; unsigned int& shots = w.shots;
            int n[OP_COUNT] = {};
            for (int s = 0; s < SLOTS; ++s) if (p.tag[s] < OP_COUNT) ++n[p.tag[s]];
            const int ph = oct_pick(n);
            if (ph < 0) { w.done = true; --live; total += shots; continue; }
            int sel[32], cnt = 0;
            for (int s = 0; s < SLOTS && cnt < 32; ++s) if (p.tag[s] == ph) sel[cnt++] = s;
            st.exec[ph] += 1; st.lanes[ph] += cnt; st.trips += 1;
            if (ray_steps && ph != OP_SF) for (int l = 0; l < cnt; ++l) ++ray_steps[p.U(OU_RAY, sel[l])];   // length of the ray's dependent chain of phase executions
            uint32_t nt[32];
            if (ph == OP_T) {
                for (int l = 0; l < cnt; ++l) nt[l] = octw_test<CHAIN, true, SLOTS>(T, polys, p, sel[l], order, out, c);
            } else if (ph == OP_C) {
                for (int l = 0; l < cnt; ++l) nt[l] = octw_cull<true, SLOTS>(T, p, sel[l], c);
            } else if (ph == OP_G) {
                for (int l = 0; l < cnt; ++l) nt[l] = octw_group<true, SLOTS>(T, p, sel[l], c);
            } else if (ph == OP_N) {
                for (int l = 0; l < cnt; ++l) nt[l] = octw_node<true, SLOTS, N_MAX>(T, F, (size_t)sel[l], p, sel[l], c);
            } else {
                for (int l = 0; l < cnt; ++l) octw_finish<CHAIN, true, SLOTS>(polys, p, sel[l], order, out, shots, c);
                int rank = 0;
                for (int l = 0; l < cnt; ++l) {
                    bool ready = true;
                    if (p.U(OU_FLAGS, sel[l]) & OFL_NORAY) {
                        const long long ray = feed_ray(f, feed, f.b1, rank);
                        ++rank;
                        if (ray < N) octw_fetch<SLOTS>(p, sel[l], ray, o, d, o1a, o2a);
                        else ready = false;
                    }
                    nt[l] = ready ? octw_setup<true, SLOTS>(T, p, sel[l], c) : (uint32_t)OP_DONE;
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

// Host arrays in, host arrays out.  stats: exec[5], lanes[5] in phase order SF, N, G, C, T; trips.
extern "C" int oct_emu(const double* verts, const double* normals, const int32_t* vcount, int64_t P,
                       const double* node_box, const int32_t* first_child, const uint32_t* list_off, const uint32_t* list_cnt, const uint32_t* lists,
                       int64_t n_nodes, int64_t n_list,
                       const double* o, const double* d, const int32_t* o1, const int32_t* o2, int64_t N, int chain, int order,
                       double* t, double* xyz, int32_t* pid, double* uv, double* omoved,
                       int32_t* ev_pid, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots, unsigned long long* total_shots,
                       int slots, int nmax, int n_warps, int regular_ok, double* stats, unsigned long long* counters, uint32_t* ray_steps,
                        double* ev_xyz, double* ev_uv /* chain: per-bounce X_Point / u, v rows, optional */) {
    std::vector<PolyRec> recs((size_t)P);
    std::vector<float> pbox6((size_t)P * 6);
    for (int64_t i = 0; i < P; ++i) {
        for (int k = 0; k < 12; ++k) recs[i].v[k] = verts[12 * i + k];
        if (vcount[i] == 3) for (int a = 0; a < 3; ++a) recs[i].v[9 + a] = verts[12 * i + 6 + a];
        for (int a = 0; a < 3; ++a) recs[i].v[12 + a] = normals[3 * i + a];
        recs[i].v[15] = (double)vcount[i];
        poly_pad_box(verts + 12 * i, vcount[i], &pbox6[6 * (size_t)i]);
    }
    OctTree tr;
    tr.box.assign(node_box, node_box + 6 * n_nodes); tr.first_child.assign(first_child, first_child + n_nodes);
    tr.list_off.assign(list_off, list_off + n_nodes); tr.list_cnt.assign(list_cnt, list_cnt + n_nodes);
    tr.polys.assign(lists, lists + n_list);
    const int depth = oct_depth_of(tr);
    PackedOct pk;
    pack_octree(tr, pbox6.data(), pk);
    std::vector<float4> pbox((size_t)P * 2);
    for (int64_t i = 0; i < P; ++i) {
        pbox[2 * i] = make_float4(pbox6[6 * i], pbox6[6 * i + 1], pbox6[6 * i + 2], 0.f);
        pbox[2 * i + 1] = make_float4(pbox6[6 * i + 3], pbox6[6 * i + 4], pbox6[6 * i + 5], 0.f);
    }
    pk.cbox.resize(pk.cbox.size() + 64, 0.f); pk.gbox.resize(pk.gbox.size() + 16, 0.f);
    OctDev T = {};
    T.nodes = pk.nodes.data(); T.lists = tr.polys.data();
    T.cbox = reinterpret_cast<const float4*>(pk.cbox.data()); T.gbox = reinterpret_cast<const float4*>(pk.gbox.data());
    T.pbox = pbox.data(); T.nbox = reinterpret_cast<const float4*>(pk.nbox.data());
    T.depth = depth; T.regular = (pk.regular && regular_ok) ? 1 : 0;
    WalkOut out = { t, xyz, pid, uv, omoved, ev_pid, ev_t, fin_o, fin_d, nshots, total_shots, nullptr, ev_xyz, ev_uv };
    Stats st;
#define RUN(S, M) if (slots == S && nmax == M) { if (chain) run<true, S, M>(T, depth, recs.data(), o, d, o1, o2, N, order, out, n_warps, st, counters, ray_steps); \
                                                 else run<false, S, M>(T, depth, recs.data(), o, d, o1, o2, N, order, out, n_warps, st, counters, ray_steps); ok = 1; }
    int ok = 0;
    RUN(64, 4) RUN(64, 1) RUN(64, 2) RUN(64, 8) RUN(48, 4) RUN(32, 4) RUN(40, 2) RUN(96, 4)
#undef RUN
    if (!ok) return -1;
    if (stats) {
        for (int k = 0; k < OP_COUNT; ++k) { stats[k] = st.exec[k]; stats[OP_COUNT + k] = st.lanes[k]; }
        stats[2 * OP_COUNT] = st.trips;
    }
    return 0;
}
