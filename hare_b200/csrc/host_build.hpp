// host_build.hpp -- host-side (CPU, C++) construction steps of the product library:
// Topology ingest and the Octree / KDTree builders.  These run once per model; the
// per-ray work is all on the GPU.  (Building the trees on the device is SURVEY.md 8(f)
// rank 1.)  Independent of oracle/: the oracle is test infrastructure and is never linked here.
#pragma once
#include <cstdint>
#include <vector>

namespace hare {

struct HostTopo {
    int64_t P = 0;
    std::vector<double> verts;     // P x 12
    std::vector<double> normals;   // P x 3
    std::vector<int32_t> vcount;   // P
    double minmax[6] = { 0, 0, 0, 0, 0, 0 };   // Topology.Min, Topology.Max
    // exact vertex bounds (what Octree / KDTree ctors compute by scanning Vertices_List)
    double vmin[3] = { 0, 0, 0 }, vmax[3] = { 0, 0, 0 };
};

// Topology(min,max) + Add_Polygon x P + Finish_Topology().  Returns 0, or -3 for a polygon that is
// not a triangle or quadrilateral.
int topology_ingest(const double* raw_verts, const int32_t* vcount, int64_t P, const double minpt[3], const double maxpt[3],
                    double* verts_out, double* normals_out, double minmax_out[6], int64_t* vertex_count_out);

struct OctTree {
    std::vector<double> box;            // N x 6
    std::vector<int32_t> first_child;   // N
    std::vector<uint32_t> list_off, list_cnt;
    std::vector<uint32_t> polys;
    int64_t lost = 0;
    int depth = 0;
};
void build_octree(const HostTopo& T, int maxDepth, int maxPolys, OctTree& out);

struct KdTree {
    std::vector<double> box;            // N x 6
    std::vector<double> split;
    std::vector<int32_t> axis, left;    // leaf: axis = -1, left = -1; Right = left + 1
    std::vector<uint32_t> list_off, list_cnt;
    std::vector<uint32_t> polys;
    int depth = 0;
};
void build_kdtree(const HostTopo& T, int maxDepth, int maxPolys, KdTree& out);

}  // namespace hare
