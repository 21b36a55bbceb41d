// pack.hpp -- host-side packing of a built partition into the arrays the traversal kernels read (device layout, DESIGN.md
// section 4).  Header-only so that tests/emu can build exactly the same arrays for the CPU replay of the kernels.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "host_build.hpp"
#include "shoot.cuh"

namespace hare {

// padded FP32 bounding box of one polygon (cull_box): exact box -/+ hare_box_pad, rounded outwards.  out = lo xyz, hi xyz
inline void poly_pad_box(const double* verts12, int vcount, float out[6]) {
    for (int a = 0; a < 3; ++a) {
        double lo = verts12[a], hi = lo;
        for (int k = 1; k < vcount; ++k) { lo = std::min(lo, verts12[3 * k + a]); hi = std::max(hi, verts12[3 * k + a]); }
        const double pad = hare_box_pad(lo, hi);
        float bl = (float)(lo - pad), bh = (float)(hi + pad);
        while ((double)bl > lo - pad) bl = std::nextafter(bl, -INFINITY);
        while ((double)bh < hi + pad) bh = std::nextafter(bh, INFINITY);
        out[a] = bl; out[3 + a] = bh;
    }
}

struct PackedOct {
    bool regular = true;          // every child box equals BuildOctree's function of its parent's box ("Octree - alt.cs":99-114), bit for bit
    std::vector<OctNode> nodes;   // box, first_child, list range; pad = content mask of the 8 children (internal) / first chunk index (leaf)
    std::vector<float> cbox;      // per run of HARE_OCT_CHUNK leaf-list entries: union of the members' padded boxes (lo.xyz, 0, hi.xyz, 0)
    std::vector<float> gbox;      // per group of 8 runs (a leaf's first run is a multiple of 8)
    std::vector<float> nbox;      // per node: union of the padded boxes of every polygon listed below it
};

// pbox6: P x 6 padded polygon boxes (poly_pad_box).  Children have larger indices than their parents (checked at upload).
inline void pack_octree(const OctTree& t, const float* pbox6, PackedOct& out) {
    const size_t N = t.first_child.size();
    out.nodes.resize(N);
    for (size_t i = 0; i < N; ++i) {
        OctNode& n = out.nodes[i];
        n.mnx = t.box[6 * i]; n.mny = t.box[6 * i + 1]; n.mnz = t.box[6 * i + 2]; n.mxx = t.box[6 * i + 3]; n.mxy = t.box[6 * i + 4]; n.mxz = t.box[6 * i + 5];
        n.first_child = t.first_child[i]; n.list_off = t.list_off[i]; n.list_cnt = t.list_cnt[i]; n.pad = 0;
    }
    out.regular = true;
    for (size_t i = 0; i < N && out.regular; ++i) {
        if (t.first_child[i] < 0) continue;
        for (int c = 0; c < 8 && out.regular; ++c) {
            const double* cb = &t.box[6 * ((size_t)t.first_child[i] + c)];
            for (int a = 0; a < 3; ++a) {
                const double mn = t.box[6 * i + a], mx = t.box[6 * i + 3 + a], mid = (mx + mn) / 2;
                const bool upper = (c & (4 >> a)) != 0;
                const double cmn = (upper ? mid : mn) - 0.1, cmx = (upper ? mx : mid) + 0.1;
                if (std::memcmp(&cmn, &cb[a], 8) != 0 || std::memcmp(&cmx, &cb[3 + a], 8) != 0) out.regular = false;
            }
        }
    }
    // internal node: pad = bit c set when the subtree of child c holds at least one polygon: the kernel never enters the others
    // (entering a polygon-free subtree has no effect on the result)
    {
        std::vector<uint8_t> has(N, 0);
        for (size_t i = N; i-- > 0;) {
            if (t.first_child[i] < 0) has[i] = t.list_cnt[i] > 0;
            else {
                uint32_t m = 0;
                for (int c = 0; c < 8; ++c) if (has[(size_t)t.first_child[i] + c]) m |= 1u << c;
                out.nodes[i].pad = m; has[i] = m != 0;
            }
        }
    }
    // chunk boxes; every leaf starts a multiple of 8 chunks (leaf.pad = its first chunk), gbox[g] encloses chunks 8g .. 8g+7
    const float kEmptyBox[8] = { INFINITY, INFINITY, INFINITY, 0.f, -INFINITY, -INFINITY, -INFINITY, 0.f };
    out.cbox.clear(); out.gbox.clear();
    for (size_t i = 0; i < N; ++i) {
        if (t.first_child[i] >= 0) continue;
        while ((out.cbox.size() / 8) % 8) out.cbox.insert(out.cbox.end(), kEmptyBox, kEmptyBox + 8);
        out.nodes[i].pad = (uint32_t)(out.cbox.size() / 8);
        for (uint32_t b = 0; b < t.list_cnt[i]; b += HARE_OCT_CHUNK) {
            const uint32_t e = std::min<uint32_t>(b + HARE_OCT_CHUNK, t.list_cnt[i]);
            float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
            for (uint32_t k = b; k < e; ++k) {
                const float* q = pbox6 + 6 * (size_t)t.polys[t.list_off[i] + k];
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], q[a]); hi[a] = std::max(hi[a], q[3 + a]); }
            }
            const float rec[8] = { lo[0], lo[1], lo[2], 0.f, hi[0], hi[1], hi[2], 0.f };
            out.cbox.insert(out.cbox.end(), rec, rec + 8);
        }
    }
    for (size_t g = 0; g * 64 < out.cbox.size(); ++g) {
        float rec[8] = { INFINITY, INFINITY, INFINITY, 0.f, -INFINITY, -INFINITY, -INFINITY, 0.f };
        for (size_t c = 8 * g; c < 8 * g + 8 && c * 8 < out.cbox.size(); ++c)
            for (int a = 0; a < 3; ++a) { rec[a] = std::min(rec[a], out.cbox[8 * c + a]); rec[4 + a] = std::max(rec[4 + a], out.cbox[8 * c + 4 + a]); }
        out.gbox.insert(out.gbox.end(), rec, rec + 8);
    }
    // node content boxes, bottom up
    out.nbox.assign(N * 8, 0.f);
    for (size_t i = N; i-- > 0;) {
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
        if (t.first_child[i] < 0) {
            for (uint32_t k = 0; k < t.list_cnt[i]; ++k) {
                const float* q = pbox6 + 6 * (size_t)t.polys[t.list_off[i] + k];
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], q[a]); hi[a] = std::max(hi[a], q[3 + a]); }
            }
        } else {
            for (int c = 0; c < 8; ++c) {
                const float* q = &out.nbox[8 * ((size_t)t.first_child[i] + c)];
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], q[a]); hi[a] = std::max(hi[a], q[4 + a]); }
            }
        }
        const float rec[8] = { lo[0], lo[1], lo[2], 0.f, hi[0], hi[1], hi[2], 0.f };
        std::memcpy(&out.nbox[8 * i], rec, sizeof rec);
    }
}

// Device node records of a kd-tree.  The node box is the reference's (KDTree.cs:68-83, 107-121) INTERSECTED with the bounding box of
// the polygons listed below the node: the walk only uses a node box to decide that no polygon of the subtree can be hit at
// t <= closest inside it (kd_box_reachable); a hit inside the node box lies on a listed polygon, hence inside that polygon's
// bounding box too, so the intersection prunes as validly and far tighter -- the median splits leave most leaves holding a sliver
// of wall in a box of air.  The reference's own boxes stay in KdTree::box (downloads, and the tie rule's side table).
inline void pack_kdtree(const KdTree& t, const HostTopo& M, std::vector<KdNode>& nodes) {
    const size_t N = t.axis.size();
    std::vector<double> tight(t.box), content(6 * N);
    std::vector<std::pair<int, int>> st; st.push_back({ 0, 0 });   // (node, 0 = enter / 1 = leave)
    while (!st.empty()) {
        auto [n, leave] = st.back(); st.pop_back();
        double* cb = &content[6 * (size_t)n];
        if (t.left[n] < 0) {
            for (int a = 0; a < 3; ++a) { cb[a] = INFINITY; cb[3 + a] = -INFINITY; }
            for (uint32_t k = 0; k < t.list_cnt[n]; ++k) {
                const int64_t q = t.polys[t.list_off[n] + k];
                for (int v = 0; v < M.vcount[q]; ++v)
                    for (int a = 0; a < 3; ++a) { const double c = M.verts[12 * q + 3 * v + a]; cb[a] = std::min(cb[a], c); cb[3 + a] = std::max(cb[3 + a], c); }
            }
        } else if (!leave) {
            st.push_back({ n, 1 }); st.push_back({ t.left[n], 0 }); st.push_back({ t.left[n] + 1, 0 });
            continue;
        } else {
            const double *l = &content[6 * (size_t)t.left[n]], *r = l + 6;
            for (int a = 0; a < 3; ++a) { cb[a] = std::min(l[a], r[a]); cb[3 + a] = std::max(l[3 + a], r[3 + a]); }
        }
        for (int a = 0; a < 3; ++a) { tight[6 * (size_t)n + a] = std::max(tight[6 * (size_t)n + a], cb[a]); tight[6 * (size_t)n + 3 + a] = std::min(tight[6 * (size_t)n + 3 + a], cb[3 + a]); }
    }
    nodes.resize(N);
    for (size_t i = 0; i < N; ++i) {
        KdNode& n = nodes[i];
        n.mnx = tight[6 * i]; n.mny = tight[6 * i + 1]; n.mnz = tight[6 * i + 2]; n.mxx = tight[6 * i + 3]; n.mxy = tight[6 * i + 4]; n.mxz = tight[6 * i + 5];
        n.left = t.left[i]; n.axis = t.left[i] >= 0 ? t.axis[i] : 0;
        if (t.left[i] >= 0) n.split = t.split[i];
        else { uint64_t bits = (uint64_t)t.list_off[i] | ((uint64_t)t.list_cnt[i] << 32); std::memcpy(&n.split, &bits, 8); }
    }
}

// the walk's 32-byte FP32 records (KdNodeC) from the packed FP64 ones: box rounded outwards and padded like a polygon box
inline void pack_kdtree_hot(const std::vector<KdNode>& nodes, std::vector<KdNodeC>& hot) {
    hot.resize(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const KdNode& n = nodes[i];
        const double mn[3] = { n.mnx, n.mny, n.mnz }, mx[3] = { n.mxx, n.mxy, n.mxz };
        float lo[3], hi[3];
        for (int a = 0; a < 3; ++a) {
            if (!(mn[a] <= mx[a])) { lo[a] = INFINITY; hi[a] = -INFINITY; continue; }     // empty content: never reachable
            const double pad = hare_box_pad(mn[a], mx[a]);
            lo[a] = (float)(mn[a] - pad); while ((double)lo[a] > mn[a] - pad) lo[a] = std::nextafter(lo[a], -INFINITY);
            hi[a] = (float)(mx[a] + pad); while ((double)hi[a] < mx[a] + pad) hi[a] = std::nextafter(hi[a], INFINITY);
        }
        KdNodeC& h = hot[i];
        h.mnx = lo[0]; h.mny = lo[1]; h.mnz = lo[2]; h.mxx = hi[0]; h.mxy = hi[1]; h.mxz = hi[2];
        if (n.left >= 0) { const float sp = (float)n.split; std::memcpy(&h.a, &sp, 4); h.b = ((uint32_t)n.left << 2) | (uint32_t)(n.axis & 3); }
        else { uint64_t bits; std::memcpy(&bits, &n.split, 8); h.a = (uint32_t)(bits & 0xffffffffull); h.b = ((uint32_t)(bits >> 32) << 2) | 3u; }
    }
}

// KdWide records: for every internal node its grandchildren (a child that is a leaf stands for itself)
inline void pack_kdtree_wide(const std::vector<KdNode>& nodes, const std::vector<KdNodeC>& hot, std::vector<KdWide>& wide) {
    const size_t N = nodes.size();
    wide.resize(N);
    KdNodeC none = {};
    none.mnx = none.mny = none.mnz = INFINITY; none.mxx = none.mxy = none.mxz = -INFINITY; none.a = 0; none.b = 2u;
    for (size_t i = 0; i < N; ++i) {
        for (int k = 0; k < 4; ++k) wide[i].e[k] = none;
        if (nodes[i].left < 0) continue;
        int n = 0;
        for (int c = 0; c < 2; ++c) {
            const size_t ch = (size_t)nodes[i].left + c;
            const size_t sub[2] = { ch, ch };
            size_t list[2]; int m = 1; list[0] = ch;
            if (nodes[ch].left >= 0) { list[0] = (size_t)nodes[ch].left; list[1] = list[0] + 1; m = 2; }
            (void)sub;
            for (int q = 0; q < m; ++q) {
                KdNodeC e = hot[list[q]];
                if (nodes[list[q]].left >= 0) { e.a = 0; e.b = ((uint32_t)list[q] << 2); }     // internal: link to its own record
                wide[i].e[n++] = e;                                                             // leaf: hot[] already holds (offset, count << 2 | 3)
            }
        }
    }
}

inline int kd_depth_of(const KdTree& t) {
    const size_t N = t.axis.size();
    std::vector<int> dep(N, 0);
    int best = 0;
    for (size_t i = 0; i < N; ++i) {
        best = std::max(best, dep[i]);
        if (t.left[i] >= 0) { dep[(size_t)t.left[i]] = dep[i] + 1; dep[(size_t)t.left[i] + 1] = dep[i] + 1; }
    }
    return best;
}

// deepest level of an octree / kd-tree whose children have larger indices than their parents (root = 0)
inline int oct_depth_of(const OctTree& t) {
    const size_t N = t.first_child.size();
    std::vector<int> dep(N, 0);
    int best = 0;
    for (size_t i = 0; i < N; ++i) {
        best = std::max(best, dep[i]);
        if (t.first_child[i] >= 0) for (int c = 0; c < 8; ++c) dep[(size_t)t.first_child[i] + c] = dep[i] + 1;
    }
    return best;
}

}  // namespace hare
