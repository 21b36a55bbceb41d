// host_build.cpp -- see host_build.hpp.  Compile with -ffp-contract=off.
#include "host_build.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <numeric>
#include <unordered_map>

#include "sat.cuh"

namespace hare {

// ---------------------------------------------------------------------------------------
// Topology ingest
// ---------------------------------------------------------------------------------------
namespace {

// Math.Round(value, 15): scale by 1e15, round half to even, scale back (only below 1e16 in magnitude).
inline double round_15_digits(double value) {
    if (std::fabs(value) < 1e16) {
        value *= 1e15;
        value = std::nearbyint(value);
        value /= 1e15;
    }
    return value;
}

struct WeldKey {
    uint64_t bucket, pos;
    bool operator==(const WeldKey& o) const { return bucket == o.bucket && pos == o.pos; }
};
struct WeldHash {
    size_t operator()(const WeldKey& k) const { return (size_t)(k.bucket * 0x9E3779B97F4A7C15ull ^ (k.pos + (k.bucket << 7))); }
};

struct Vec3 { double x, y, z; };

}  // namespace

int topology_ingest(const double* raw, const int32_t* vcount, int64_t P, const double minpt[3], const double maxpt[3],
                    double* verts_out, double* normals_out, double minmax_out[6], int64_t* vertex_count_out) {
    // Topology(Point Minpt, Point Maxpt): padded bounds define the weld lattice ("Modspace").
    // Hare_Geometry_Topology.cs:85-91, MS_AABB :677-697.
    const double ms_min[3] = { minpt[0] - 0.000000000001, minpt[1] - 0.000000000001, minpt[2] - 0.000000000001 };
    const double ms_max[3] = { maxpt[0] + 0.000000000001, maxpt[1] + 0.000000000001, maxpt[2] + 0.000000000001 };
    int dim = 0;
    for (int a = 0; a < 3; ++a) dim = std::max(dim, (int)std::ceil(ms_max[a] - ms_min[a]));
    const uint64_t ydim = (uint64_t)(int64_t)dim, xytot = (uint64_t)((int64_t)dim * (int64_t)dim);

    std::unordered_map<WeldKey, Vec3, WeldHash> lattice;
    lattice.reserve((size_t)P * 2);
    double lo[3] = { DBL_MAX, DBL_MAX, DBL_MAX }, hi[3] = { -DBL_MAX, -DBL_MAX, -DBL_MAX };
    int64_t nverts = 0;

    for (int64_t p = 0; p < P; ++p) {
        const int n = vcount[p];
        if (n != 3 && n != 4) return -3;
        Vec3 V[4];
        for (int k = 0; k < n; ++k) {
            // AddGetIndex :342-377 -- round, then look the 1 mm cell up; the first vertex seen in a cell wins.
            Vec3 q = { round_15_digits(raw[12 * p + 3 * k]), round_15_digits(raw[12 * p + 3 * k + 1]), round_15_digits(raw[12 * p + 3 * k + 2]) };
            const double off[3] = { q.x - ms_min[0], q.y - ms_min[1], q.z - ms_min[2] };   // Point.Hash2  Primitives.cs:237-250
            uint64_t cell[3], sub[3];
            for (int a = 0; a < 3; ++a) {
                cell[a] = (uint64_t)std::floor(off[a]);
                sub[a] = (uint64_t)((off[a] - (double)cell[a]) * 1000);
            }
            WeldKey key = { xytot * cell[2] + ydim * cell[0] + cell[1], 1000000ull * sub[2] + 1000ull * sub[0] + sub[1] };
            auto ins = lattice.emplace(key, q);
            if (ins.second) {
                ++nverts;
                lo[0] = std::min(lo[0], q.x); lo[1] = std::min(lo[1], q.y); lo[2] = std::min(lo[2], q.z);
                hi[0] = std::max(hi[0], q.x); hi[1] = std::max(hi[1], q.y); hi[2] = std::max(hi[2], q.z);
            }
            V[k] = ins.first->second;
        }
        if (n == 3) V[3] = V[2];
        // Polygon ctor normal  Hare_Geometry_Polygons.cs:159-171 with Hare_math.Cross(Vector,Vector) and Vector.Normalize
        Vec3 N = { 0, 0, 0 };
        const Vec3 a = { V[1].x - V[0].x, V[1].y - V[0].y, V[1].z - V[0].z };
        for (int j = 2; j < n; ++j) {
            const Vec3 b = { V[j].x - V[0].x, V[j].y - V[0].y, V[j].z - V[0].z };
            N.x = a.y * b.z - a.z * b.y;
            N.y = -(a.x * b.z - a.z * b.x);
            N.z = a.x * b.y - a.y * b.x;
            const double len2 = N.x * N.x + N.y * N.y + N.z * N.z;
            if (!(len2 < 4.9406564584124654e-324)) break;   // IsZeroVector uses double.Epsilon
        }
        double f = N.x * N.x + N.y * N.y + N.z * N.z;
        if (f != 0) { f = std::sqrt(f); N.x /= f; N.y /= f; N.z /= f; }
        for (int k = 0; k < 4; ++k) { verts_out[12 * p + 3 * k] = V[k].x; verts_out[12 * p + 3 * k + 1] = V[k].y; verts_out[12 * p + 3 * k + 2] = V[k].z; }
        normals_out[3 * p] = N.x; normals_out[3 * p + 1] = N.y; normals_out[3 * p + 2] = N.z;
    }
    // Finish_Topology :148-167
    for (int a = 0; a < 3; ++a) { minmax_out[a] = lo[a] - 0.000000000001; minmax_out[3 + a] = hi[a] + 0.000000000001; }
    if (vertex_count_out) *vertex_count_out = nverts;
    return 0;
}

// ---------------------------------------------------------------------------------------
// Octree  ("Octree - alt.cs":45-138), built breadth first.
// ---------------------------------------------------------------------------------------
void build_octree(const HostTopo& T, int maxDepth, int maxPolys, OctTree& out) {
    out = OctTree();
    // root cube :65-85.  "center = max + min / 2" is the reference's expression (precedence as written).
    const double ext = net_max(T.vmax[0] - T.vmin[0], net_max(T.vmax[1] - T.vmin[1], T.vmax[2] - T.vmin[2]));
    double c[3];
    for (int a = 0; a < 3; ++a) c[a] = T.vmax[a] + T.vmin[a] / 2;
    auto add_node = [&](const double mn[3], const double mx[3]) {
        for (int a = 0; a < 3; ++a) out.box.push_back(mn[a]);
        for (int a = 0; a < 3; ++a) out.box.push_back(mx[a]);
        out.first_child.push_back(-1); out.list_off.push_back(0); out.list_cnt.push_back(0);
        return (int)out.first_child.size() - 1;
    };
    {
        double mn[3], mx[3];
        for (int a = 0; a < 3; ++a) { mn[a] = c[a] - ext - 1e-1; mx[a] = c[a] + ext + 1e-1; }
        add_node(mn, mx);
    }
    struct Work { int node; std::vector<uint32_t> list; };
    std::vector<Work> level(1);
    level[0].node = 0;
    level[0].list.resize((size_t)T.P);
    std::iota(level[0].list.begin(), level[0].list.end(), 0u);

    for (int depth = 0; !level.empty(); ++depth) {
        out.depth = depth;
        std::vector<Work> next;
        for (Work& w : level) {
            if (depth >= maxDepth || (int64_t)w.list.size() <= (int64_t)maxPolys) {   // :93
                out.list_off[w.node] = (uint32_t)out.polys.size();
                out.list_cnt[w.node] = (uint32_t)w.list.size();
                out.polys.insert(out.polys.end(), w.list.begin(), w.list.end());
                continue;
            }
            double mn[3], mx[3], mid[3];
            for (int a = 0; a < 3; ++a) { mn[a] = out.box[6 * w.node + a]; mx[a] = out.box[6 * w.node + 3 + a]; mid[a] = (mx[a] + mn[a]) / 2; }
            Box3 cb[8];
            const int fc = (int)out.first_child.size();
            out.first_child[w.node] = fc;
            for (int i = 0; i < 8; ++i) {   // octant bit 4 -> x, 2 -> y, 1 -> z; each side padded by 0.1  :99-114
                double cmn[3], cmx[3];
                for (int a = 0; a < 3; ++a) {
                    const bool upper = (i & (4 >> a)) != 0;
                    cmn[a] = (upper ? mid[a] : mn[a]) - 0.1;
                    cmx[a] = (upper ? mx[a] : mid[a]) + 0.1;
                }
                add_node(cmn, cmx);
                cb[i] = make_box(cmn[0], cmn[1], cmn[2], cmx[0], cmx[1], cmx[2]);
            }
            const size_t base = next.size();
            next.resize(base + 8);
            for (int i = 0; i < 8; ++i) next[base + i].node = fc + i;
            for (uint32_t p : w.list) {   // parent order is kept; a polygon may enter several children or none (:118-130)
                bool stored = false;
                for (int i = 0; i < 8; ++i)
                    if (poly_box_overlap(cb[i], &T.verts[12 * (size_t)p], T.vcount[p])) { next[base + i].list.push_back(p); stored = true; }
                if (!stored) ++out.lost;
            }
            std::vector<uint32_t>().swap(w.list);
        }
        level.swap(next);
    }
}

// ---------------------------------------------------------------------------------------
// KDTree  (KDTree.cs:51-139), built breadth first.
// ---------------------------------------------------------------------------------------
void build_kdtree(const HostTopo& T, int maxDepth, int maxPolys, KdTree& out) {
    out = KdTree();
    // Polygon_Centroid  Hare_Geometry_Topology.cs:566-575: running sum from a zero Point, then / VertexCount
    std::vector<double> cen((size_t)T.P * 3);
    for (int64_t p = 0; p < T.P; ++p) {
        double s[3] = { 0, 0, 0 };
        const int n = T.vcount[p];
        for (int k = 0; k < n; ++k) for (int a = 0; a < 3; ++a) s[a] = s[a] + T.verts[12 * p + 3 * k + a];
        for (int a = 0; a < 3; ++a) cen[3 * p + a] = s[a] / n;
    }
    auto add_node = [&](const double mn[3], const double mx[3]) {
        for (int a = 0; a < 3; ++a) out.box.push_back(mn[a]);
        for (int a = 0; a < 3; ++a) out.box.push_back(mx[a]);
        out.split.push_back(0); out.axis.push_back(-1); out.left.push_back(-1); out.list_off.push_back(0); out.list_cnt.push_back(0);
        return (int)out.axis.size() - 1;
    };
    add_node(T.vmin, T.vmax);   // root box = exact vertex bounds :68-83
    struct Work { int node; std::vector<uint32_t> list; };
    std::vector<Work> level(1);
    level[0].node = 0;
    level[0].list.resize((size_t)T.P);
    std::iota(level[0].list.begin(), level[0].list.end(), 0u);

    for (int depth = 0; !level.empty(); ++depth) {
        out.depth = depth;
        const int axis = depth % 3;   // :95
        std::vector<Work> next;
        for (Work& w : level) {
            if (depth >= maxDepth || (int64_t)w.list.size() <= (int64_t)maxPolys) {   // :92
                out.list_off[w.node] = (uint32_t)out.polys.size();
                out.list_cnt[w.node] = (uint32_t)w.list.size();
                out.polys.insert(out.polys.end(), w.list.begin(), w.list.end());
                continue;
            }
            // OrderBy(centroid[axis]) is a stable sort  :98-101
            std::stable_sort(w.list.begin(), w.list.end(), [&](uint32_t x, uint32_t y) { return cen[3 * (size_t)x + axis] < cen[3 * (size_t)y + axis]; });
            const double sv = cen[3 * (size_t)w.list[w.list.size() / 2] + axis];   // :104-105
            double mn[3], mx[3], lmx[3], rmn[3];
            for (int a = 0; a < 3; ++a) { mn[a] = out.box[6 * w.node + a]; mx[a] = out.box[6 * w.node + 3 + a]; lmx[a] = mx[a]; rmn[a] = mn[a]; }
            lmx[axis] = sv; rmn[axis] = sv;
            const int li = add_node(mn, lmx);
            add_node(rmn, mx);
            out.axis[w.node] = axis; out.split[w.node] = sv; out.left[w.node] = li;
            const size_t base = next.size();
            next.resize(base + 2);
            next[base].node = li; next[base + 1].node = li + 1;
            for (uint32_t p : w.list) {   // sorted order :123-133
                bool le = false, gt = false;
                for (int k = 0; k < T.vcount[p]; ++k) {
                    const double v = T.verts[12 * (size_t)p + 3 * k + axis];
                    if (v <= sv) le = true;
                    if (v > sv) gt = true;
                }
                if (le) next[base].list.push_back(p);
                if (gt) next[base + 1].list.push_back(p);
            }
            std::vector<uint32_t>().swap(w.list);
        }
        level.swap(next);
    }
}

}  // namespace hare
