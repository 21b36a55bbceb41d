// kd_emu.cu -- TEST INFRASTRUCTURE: replays the wavefront scheduler of hare_b200/csrc/kd_wave.cuh on the CPU (see oct_emu.cu).
// The device arrays are built by the same pack_kdtree() the library uses.  Nothing here is shipped or measured.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../hare_b200/csrc/kernels.cuh"
#include "../../hare_b200/csrc/kd_wave.cuh"
#include "../../hare_b200/csrc/pack.hpp"

using namespace hare;

namespace {

struct Stats { double exec[KP_COUNT] = {}, lanes[KP_COUNT] = {}, trips = 0; };

template <bool CHAIN, int SLOTS, int N_MAX>
void run(const KdDev& T, const PolyRec* polys, const double* o, const double* d, const int32_t* o1a, const int32_t* o2a, const int32_t* rid,
         long long N, int order, const WalkOut& out, int tw, Stats& st, unsigned long long* counters) {
    // The simulated warps take turns, one trip each: their pools are in flight together and their claims on the launch's counter
    // interleave, as on the device (where the order is arbitrary -- results must not depend on it).
    struct Warp {
        std::vector<unsigned char> mem; std::vector<uint4> stk;
        KdPool<SLOTS> p; KdStacks S; RayFeed f; unsigned int shots = 0; bool done = false;
    };
    const int sdepth = 3 * (T.depth / 2 + 2) + 4;
    CntT<true> c;
    unsigned long long total = 0;
    unsigned long long feed_ctr = 0;
    const RayFeedArgs feed = { &feed_ctr, tw * feed_block_for(N, tw), feed_block_for(N, tw) };
    std::vector<Warp> warps((size_t)tw);
    for (long long gw = 0; gw < tw; ++gw) {
        Warp& w = warps[(size_t)gw];
        w.mem.resize(KdPool<SLOTS>::STRIDE + 64); w.stk.resize((size_t)SLOTS * sdepth);
        w.S = KdStacks{ w.stk.data(), sdepth };
        w.p.bind(w.mem.data());
        for (int s = 0; s < SLOTS; ++s) { w.p.U(KU_FLAGS, s) = KFL_NORAY; w.p.U(KU_LPOS, s) = 0; w.p.U(KU_LEND, s) = 0; w.p.tag[s] = (uint8_t)KP_SF; }
        w.f = RayFeed{ gw * feed.block, 0, 0 };
        w.f.b1 = feed_claim(feed);
    }
    for (long long live = tw; live > 0;) {
        for (long long gw = 0; gw < tw; ++gw) {
            Warp& w = warps[(size_t)gw];
            if (w.done) continue;
            KdPool<SLOTS>& p = w.p; const KdStacks& S = w.S; RayFeed& f = w.f; unsigned int& shots = w.shots;
            int n[KP_COUNT] = {};
            for (int s = 0; s < SLOTS; ++s) if (p.tag[s] < KP_COUNT) ++n[p.tag[s]];
            const int ph = kd_pick(n);
            if (ph < 0) { w.done = true; --live; total += shots; continue; }
            int sel[32], cnt = 0;
            for (int s = 0; s < SLOTS && cnt < 32; ++s) if (p.tag[s] == ph) sel[cnt++] = s;
            st.exec[ph] += 1; st.lanes[ph] += cnt; st.trips += 1;
            uint32_t nt[32];
            if (ph == KP_T) {
                for (int l = 0; l < cnt; ++l) nt[l] = kdw_test<true, SLOTS>(T, polys, p, sel[l], c);
            } else if (ph == KP_C) {
                for (int l = 0; l < cnt; ++l) nt[l] = kdw_cull<true, SLOTS>(T, p, sel[l], c);
            } else if (ph == KP_N) {
                for (int l = 0; l < cnt; ++l) nt[l] = kdw_node<true, SLOTS, N_MAX>(T, S, (size_t)sel[l], p, sel[l], c);
            } else {
                for (int l = 0; l < cnt; ++l) kdw_finish<CHAIN, true, SLOTS>(polys, p, sel[l], order, out, shots, c);
                int rank = 0;
                for (int l = 0; l < cnt; ++l) {
                    bool ready = true;
                    if (p.U(KU_FLAGS, sel[l]) & KFL_NORAY) {
                        const long long ray = feed_ray(f, feed, f.b1, rank);
                        ++rank;
                        if (ray < N) kdw_fetch<SLOTS>(p, sel[l], ray, o, d, o1a, o2a, rid);
                        else ready = false;
                    }
                    nt[l] = ready ? kdw_setup<true, SLOTS>(T, p, sel[l], c) : (uint32_t)KP_DONE;
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

extern "C" int kd_emu(const double* verts, const double* normals, const int32_t* vcount, int64_t P,
                      const double* node_box, const double* split, const int32_t* axis, const int32_t* left,
                      const uint32_t* list_off, const uint32_t* list_cnt, const uint32_t* lists, int64_t n_nodes, int64_t n_list,
                      const double* o, const double* d, const int32_t* o1, const int32_t* o2, const int32_t* rid, int64_t N, int chain, int order,
                      double* t, double* xyz, int32_t* pid, double* uv, double* omoved,
                      int32_t* ev_pid, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots, unsigned long long* total_shots,
                      int slots, int nmax, int n_warps, int tie_rule_on_tight_boxes, double* stats, unsigned long long* counters,
                        double* ev_xyz, double* ev_uv /* chain: per-bounce X_Point / u, v rows, optional */) {
    std::vector<PolyRec> recs((size_t)P);
    HostTopo M;
    M.P = P; M.verts.assign(verts, verts + 12 * P); M.vcount.assign(vcount, vcount + P);
    for (int64_t i = 0; i < P; ++i) {
        for (int k = 0; k < 12; ++k) recs[i].v[k] = verts[12 * i + k];
        if (vcount[i] == 3) for (int a = 0; a < 3; ++a) recs[i].v[9 + a] = verts[12 * i + 6 + a];
        for (int a = 0; a < 3; ++a) recs[i].v[12 + a] = normals[3 * i + a];
        recs[i].v[15] = (double)vcount[i];
    }
    KdTree tr;
    tr.box.assign(node_box, node_box + 6 * n_nodes); tr.split.assign(split, split + n_nodes); tr.axis.assign(axis, axis + n_nodes);
    tr.left.assign(left, left + n_nodes); tr.list_off.assign(list_off, list_off + n_nodes); tr.list_cnt.assign(list_cnt, list_cnt + n_nodes);
    tr.polys.assign(lists, lists + n_list);
    std::vector<KdNode> nodes; std::vector<KdNodeC> hot;
    pack_kdtree(tr, M, nodes);
    pack_kdtree_hot(nodes, hot);
    std::vector<KdWide> wide;
    pack_kdtree_wide(nodes, hot, wide);
    std::vector<float4> lbox(2 * (size_t)n_list + 16);
    for (int64_t k = 0; k < n_list; ++k) {
        float b[6];
        poly_pad_box(verts + 12 * (size_t)lists[k], vcount[lists[k]], b);
        lbox[2 * k] = make_float4(b[0], b[1], b[2], hare_u2f(lists[k])); lbox[2 * k + 1] = make_float4(b[3], b[4], b[5], 0.f);
    }
    KdDev T = {};
    T.wide = wide.data(); T.hot = hot.data(); T.nodes = nodes.data(); T.lists = tr.polys.data(); T.lbox = lbox.data(); T.depth = kd_depth_of(tr); T.ref_box = tr.box.data();
    // what-if for the tie test's teeth: evaluate the reference's first/second rule on the content-tightened device boxes (the round-1 bug)
    std::vector<double> tightbox;
    if (tie_rule_on_tight_boxes) {
        for (const KdNode& n : nodes) { const double b[6] = { n.mnx, n.mny, n.mnz, n.mxx, n.mxy, n.mxz }; tightbox.insert(tightbox.end(), b, b + 6); }
        T.ref_box = tightbox.data();
    }
    WalkOut out = { t, xyz, pid, uv, omoved, ev_pid, ev_t, fin_o, fin_d, nshots, total_shots, nullptr, ev_xyz, ev_uv };
    Stats st;
#define RUN(S, M) if (slots == S && nmax == M) { if (chain) run<true, S, M>(T, recs.data(), o, d, o1, o2, rid, N, order, out, n_warps, st, counters); \
                                                 else run<false, S, M>(T, recs.data(), o, d, o1, o2, rid, N, order, out, n_warps, st, counters); ok = 1; }
    int ok = 0;
    RUN(64, 4) RUN(64, 1) RUN(64, 2) RUN(64, 8) RUN(48, 4) RUN(32, 4) RUN(40, 2) RUN(96, 4)
#undef RUN
    if (!ok) return -1;
    if (stats) {
        for (int k = 0; k < KP_COUNT; ++k) { stats[k] = st.exec[k]; stats[KP_COUNT + k] = st.lanes[k]; }
        stats[2 * KP_COUNT] = st.trips;
    }
    return 0;
}
