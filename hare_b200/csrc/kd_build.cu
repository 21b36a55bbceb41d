// kd_build.cu -- KDTree construction on the GPU: new KDTree(Model, maxDepth, maxPolygonsPerNode),
// KDTree.cs:51-139, level-synchronous.  Produces the same KdTree as host_build.cpp's build_kdtree(): same breadth-first
// node numbering, node boxes, split values and list order (tested identical).
//
// The reference, per node: axis = depth % 3; polygons ordered by centroid[axis] with a STABLE sort (LINQ OrderBy, :98-101);
// split = centroid of the median polygon (:104-105); left child <= every polygon with a vertex <= split, right child <=
// every polygon with a vertex > split, both in the sorted order (:123-133).  A node is a leaf when depth >= maxDepth or it
// holds <= maxPolygonsPerNode polygons (:92); a leaf keeps the order its parent's partition gave it.
//
// Here every level is one flat array of (segment, polygon) entries, one segment per node of the level:
//   1. centroids are static, so each axis gets a DENSE RANK per polygon once (sort of the P centroids, equal values share a
//      rank): the reference's comparator only orders by value, and ties must keep the current list order;
//   2. per level ONE stable LSD radix sort (cub::DeviceRadixSort) of the keys (segment << rank_bits | rank[axis]) is the
//      stable per-node OrderBy of all nodes of the level at once;
//   3. median -> split per segment; le / gt flags from the per-polygon vertex min / max on the axis; two exclusive scans;
//      an order-preserving scatter writes the children's lists (the next level's array);
//   4. leaf segments are copied to the final list buffer first and leave the array.
// The host only keeps the node records (boxes, split, axis, child index) from the per-level counts and split values.
// cub::DeviceRadixSort / DeviceScan are used as plumbing (sort and prefix sum); the SAT-free arithmetic here is compares only.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "hare_math.cuh"
#include "host_build.hpp"

namespace hare {

namespace {

#define KCK(call)                                                                                             \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) { err = std::string("kd build: ") + cudaGetErrorString(e_); return -2; }      \
    } while (0)

// per polygon: Polygon_Centroid (Hare_Geometry_Topology.cs:566-575: running sum from zero, then / VertexCount) and the
// vertex min / max on every axis
__global__ void __launch_bounds__(256)
kd_poly_stats(const PolyRec* __restrict__ polys, long long P, double* __restrict__ cen, double* __restrict__ vmin, double* __restrict__ vmax) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double* v = polys[p].v;
    const int n = (int)v[15];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double s = 0, lo = v[a], hi = v[a];
        for (int k = 0; k < n; ++k) { const double c = v[3 * k + a]; s = s + c; lo = c < lo ? c : lo; hi = c > hi ? c : hi; }
        cen[a * P + p] = s / n;
        vmin[a * P + p] = lo; vmax[a * P + p] = hi;
    }
}

// IEEE double -> unsigned key with the same order; -0.0 and +0.0 compare equal in the reference's comparator
__device__ __forceinline__ unsigned long long orderable(double x) {
    if (x == 0.0) x = 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256)
kd_axis_keys(const double* __restrict__ cen, long long P, unsigned long long* __restrict__ keys, uint32_t* __restrict__ ids) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) { keys[p] = orderable(cen[p]); ids[p] = (uint32_t)p; }
}

__global__ void __launch_bounds__(256)
kd_rank_flags(const unsigned long long* __restrict__ sorted_keys, long long P, uint32_t* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) flag[i] = (i > 0 && sorted_keys[i] != sorted_keys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
kd_rank_scatter(const uint32_t* __restrict__ sorted_ids, const uint32_t* __restrict__ dense /* inclusive scan of flag */, long long P, uint32_t* __restrict__ rank) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) rank[sorted_ids[i]] = dense[i];
}

// per entry of the level: leaf segments copy their entry to the final lists; split segments emit their sort key
// seg_info[s] >= 0: leaf, value = offset of its list in the final buffer; < 0: split segment number -(v + 1)
__global__ void __launch_bounds__(256)
kd_level_keys(const uint32_t* __restrict__ list, const uint32_t* __restrict__ seg, const uint32_t* __restrict__ seg_start,
              const long long* __restrict__ seg_info, const uint32_t* __restrict__ rank, int rank_bits, long long n,
              uint32_t* __restrict__ final_list, unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = seg[i], p = list[i];
    const long long info = seg_info[s];
    if (info >= 0) final_list[info + (i - seg_start[s])] = p;
    keys[i] = ((unsigned long long)s << rank_bits) | rank[p];
}

// split value of every split segment: centroid of the median polygon of the sorted list (KDTree.cs:104-105)
__global__ void __launch_bounds__(256)
kd_level_split(const uint32_t* __restrict__ sorted_list, const uint32_t* __restrict__ seg_start, const long long* __restrict__ seg_info,
               const double* __restrict__ cen, int S, double* __restrict__ split) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    if (seg_info[s] >= 0) { split[s] = 0.0; return; }
    const uint32_t a = seg_start[s], cnt = seg_start[s + 1] - a;
    split[s] = cen[sorted_list[a + cnt / 2]];
}

// left <= any vertex <= split, right <= any vertex > split (KDTree.cs:123-133); leaf segments pass nothing on
__global__ void __launch_bounds__(256)
kd_level_flags(const uint32_t* __restrict__ sorted_list, const uint32_t* __restrict__ sorted_seg, const long long* __restrict__ seg_info,
               const double* __restrict__ split, const double* __restrict__ vmin, const double* __restrict__ vmax, long long n,
               uint32_t* __restrict__ le, uint32_t* __restrict__ gt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = sorted_seg[i], p = sorted_list[i];
    const bool live = seg_info[s] < 0;
    const double sv = split[s];
    le[i] = (live && vmin[p] <= sv) ? 1u : 0u;
    gt[i] = (live && vmax[p] > sv) ? 1u : 0u;
}

// children sizes of every split segment, in child order (L0, R0, L1, R1, ...)
__global__ void __launch_bounds__(256)
kd_level_counts(const uint32_t* __restrict__ seg_start, const long long* __restrict__ seg_info, const uint32_t* __restrict__ sl,
                const uint32_t* __restrict__ sg, int S, uint32_t* __restrict__ child_cnt) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const long long info = seg_info[s];
    if (info >= 0) return;
    const long long k = -(info + 1);
    child_cnt[2 * k] = sl[seg_start[s + 1]] - sl[seg_start[s]];
    child_cnt[2 * k + 1] = sg[seg_start[s + 1]] - sg[seg_start[s]];
}

// order-preserving scatter into the next level's array
__global__ void __launch_bounds__(256)
kd_level_scatter(const uint32_t* __restrict__ sorted_list, const uint32_t* __restrict__ sorted_seg, const uint32_t* __restrict__ seg_start,
                 const long long* __restrict__ seg_info, const uint32_t* __restrict__ le, const uint32_t* __restrict__ gt,
                 const uint32_t* __restrict__ sl, const uint32_t* __restrict__ sg, const uint32_t* __restrict__ child_start, long long n,
                 uint32_t* __restrict__ next_list, uint32_t* __restrict__ next_seg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = sorted_seg[i];
    const long long info = seg_info[s];
    if (info >= 0) return;
    const long long k = -(info + 1);
    const uint32_t a = seg_start[s], p = sorted_list[i];
    if (le[i]) { const uint32_t q = child_start[2 * k] + (sl[i] - sl[a]); next_list[q] = p; next_seg[q] = (uint32_t)(2 * k); }
    if (gt[i]) { const uint32_t q = child_start[2 * k + 1] + (sg[i] - sg[a]); next_list[q] = p; next_seg[q] = (uint32_t)(2 * k + 1); }
}

__global__ void __launch_bounds__(256)
kd_iota(uint32_t* __restrict__ list, uint32_t* __restrict__ seg, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { list[i] = (uint32_t)i; seg[i] = 0; }
}

template <class T>
struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n, bool keep = false, cudaStream_t st = nullptr) {
        if (n <= cap) return cudaSuccess;
        const size_t ncap = std::max<size_t>(n + n / 4, 1024);
        T* q = nullptr;
        cudaError_t e = cudaMalloc((void**)&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        if (keep && p && cap) { e = cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st); if (e == cudaSuccess) e = cudaStreamSynchronize(st); }
        cudaFree(p);
        p = q; cap = ncap;
        return e;
    }
    ~DevBuf() { cudaFree(p); }
};

inline unsigned grid_for(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

unsigned long long g_kd_build_launches = 0;   // added to hare_launch_count() by hare_abi.cu

int build_kdtree_gpu(const HostTopo& M, const PolyRec* d_polys, int dev, cudaStream_t st, int maxDepth, int maxPolys, KdTree& out, std::string& err) {
    out = KdTree();
    const long long P = M.P;
    KCK(cudaSetDevice(dev));
    int rank_bits = 1;
    while ((1ll << rank_bits) < P) ++rank_bits;

    DevBuf<double> cen, vmin, vmax, split;
    DevBuf<uint32_t> rank, list[2], seg[2], vals_tmp, le, gt, sl, sg, final_list, seg_start, child_cnt, child_start;
    DevBuf<unsigned long long> keys[2];
    DevBuf<long long> seg_info;
    DevBuf<unsigned char> tmp;
    KCK(cen.ensure(3 * (size_t)P)); KCK(vmin.ensure(3 * (size_t)P)); KCK(vmax.ensure(3 * (size_t)P)); KCK(rank.ensure(3 * (size_t)P));
    auto ensure_level = [&](size_t n) -> cudaError_t {
        cudaError_t e;
        for (int k = 0; k < 2; ++k) {
            if ((e = list[k].ensure(n, true, st)) != cudaSuccess) return e;
            if ((e = seg[k].ensure(n, true, st)) != cudaSuccess) return e;
            if ((e = keys[k].ensure(n)) != cudaSuccess) return e;
        }
        if ((e = vals_tmp.ensure(n)) != cudaSuccess) return e;
        if ((e = le.ensure(n + 1)) != cudaSuccess) return e;
        if ((e = gt.ensure(n + 1)) != cudaSuccess) return e;
        if ((e = sl.ensure(n + 1)) != cudaSuccess) return e;
        if ((e = sg.ensure(n + 1)) != cudaSuccess) return e;
        // temp storage of the widest cub call on n items
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, keys[0].p, keys[1].p, list[0].p, list[1].p, (long long)n, 0, 64, st);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, le.p, sl.p, (long long)n + 1, st);
        cub::DeviceScan::InclusiveSum(nullptr, b3, le.p, sl.p, (long long)n + 1, st);
        return tmp.ensure(std::max(b1, std::max(b2, b3)) + 256);
    };
    KCK(ensure_level((size_t)P));

    kd_poly_stats<<<grid_for(P), 256, 0, st>>>(d_polys, P, cen.p, vmin.p, vmax.p);
    ++g_kd_build_launches;
    // dense rank of every polygon's centroid, per axis
    for (int a = 0; a < 3; ++a) {
        kd_axis_keys<<<grid_for(P), 256, 0, st>>>(cen.p + a * P, P, keys[0].p, list[0].p);
        size_t bytes = tmp.cap;
        KCK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys[0].p, keys[1].p, list[0].p, list[1].p, P, 0, 64, st));
        kd_rank_flags<<<grid_for(P), 256, 0, st>>>(keys[1].p, P, le.p);
        bytes = tmp.cap;
        KCK(cub::DeviceScan::InclusiveSum(tmp.p, bytes, le.p, sl.p, P, st));
        kd_rank_scatter<<<grid_for(P), 256, 0, st>>>(list[1].p, sl.p, P, rank.p + a * P);
        g_kd_build_launches += 5;
    }
    KCK(cudaGetLastError());

    // host-side node records, exactly as build_kdtree() keeps them
    auto add_node = [&](const double mn[3], const double mx[3]) {
        for (int a = 0; a < 3; ++a) out.box.push_back(mn[a]);
        for (int a = 0; a < 3; ++a) out.box.push_back(mx[a]);
        out.split.push_back(0); out.axis.push_back(-1); out.left.push_back(-1); out.list_off.push_back(0); out.list_cnt.push_back(0);
        return (int)out.axis.size() - 1;
    };
    add_node(M.vmin, M.vmax);   // root box = exact vertex bounds :68-83
    std::vector<int> seg_node(1, 0);               // node of every segment of the level
    std::vector<uint32_t> h_start(2); h_start[0] = 0; h_start[1] = (uint32_t)P;
    long long n = P;
    size_t final_n = 0;
    int cur = 0;
    kd_iota<<<grid_for(P), 256, 0, st>>>(list[0].p, seg[0].p, P);
    ++g_kd_build_launches;

    for (int depth = 0; n > 0 || !seg_node.empty(); ++depth) {
        const int S = (int)seg_node.size();
        if (S == 0) break;
        out.depth = depth;
        const int axis = depth % 3;
        // leaf or split?  (:92)
        std::vector<long long> h_info(S);
        int n_split = 0;
        size_t leaf_total = 0;
        for (int s = 0; s < S; ++s) {
            const uint32_t cnt = h_start[s + 1] - h_start[s];
            if (depth >= maxDepth || (long long)cnt <= (long long)maxPolys) {
                out.list_off[seg_node[s]] = (uint32_t)(final_n + leaf_total);
                out.list_cnt[seg_node[s]] = cnt;
                h_info[s] = (long long)(final_n + leaf_total);
                leaf_total += cnt;
            } else {
                h_info[s] = -((long long)n_split + 1);
                ++n_split;
            }
        }
        if (n == 0) break;
        KCK(final_list.ensure(final_n + leaf_total, true, st));
        KCK(seg_start.ensure((size_t)S + 1)); KCK(seg_info.ensure((size_t)S)); KCK(split.ensure((size_t)S));
        KCK(child_cnt.ensure(2 * (size_t)n_split + 2)); KCK(child_start.ensure(2 * (size_t)n_split + 2));
        KCK(cudaMemcpyAsync(seg_start.p, h_start.data(), ((size_t)S + 1) * 4, cudaMemcpyHostToDevice, st));
        KCK(cudaMemcpyAsync(seg_info.p, h_info.data(), (size_t)S * 8, cudaMemcpyHostToDevice, st));
        int seg_bits = 1;
        while ((1ll << seg_bits) < S) ++seg_bits;
        kd_level_keys<<<grid_for(n), 256, 0, st>>>(list[cur].p, seg[cur].p, seg_start.p, seg_info.p, rank.p + (size_t)axis * P, rank_bits, n,
                                                   final_list.p, keys[0].p);
        ++g_kd_build_launches;
        final_n += leaf_total;
        if (n_split == 0) { KCK(cudaStreamSynchronize(st)); break; }
        // the stable per-node OrderBy of the whole level: one LSD radix sort over (segment, rank) keys, twice (list and seg ride along)
        size_t bytes = tmp.cap;
        KCK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys[0].p, keys[1].p, list[cur].p, list[cur ^ 1].p, n, 0, rank_bits + seg_bits, st));
        bytes = tmp.cap;
        KCK(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys[0].p, keys[1].p, seg[cur].p, seg[cur ^ 1].p, n, 0, rank_bits + seg_bits, st));
        const uint32_t* s_list = list[cur ^ 1].p; const uint32_t* s_seg = seg[cur ^ 1].p;
        kd_level_split<<<grid_for(S), 256, 0, st>>>(s_list, seg_start.p, seg_info.p, cen.p + (size_t)axis * P, S, split.p);
        kd_level_flags<<<grid_for(n), 256, 0, st>>>(s_list, s_seg, seg_info.p, split.p, vmin.p + (size_t)axis * P, vmax.p + (size_t)axis * P, n, le.p, gt.p);
        bytes = tmp.cap;
        KCK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, le.p, sl.p, n + 1, st));
        bytes = tmp.cap;
        KCK(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, gt.p, sg.p, n + 1, st));
        kd_level_counts<<<grid_for(S), 256, 0, st>>>(seg_start.p, seg_info.p, sl.p, sg.p, S, child_cnt.p);
        g_kd_build_launches += 7;
        std::vector<double> h_split(S);
        std::vector<uint32_t> h_cnt(2 * (size_t)n_split);
        KCK(cudaMemcpyAsync(h_split.data(), split.p, (size_t)S * 8, cudaMemcpyDeviceToHost, st));
        KCK(cudaMemcpyAsync(h_cnt.data(), child_cnt.p, h_cnt.size() * 4, cudaMemcpyDeviceToHost, st));
        KCK(cudaStreamSynchronize(st));
        // node records of the children (:107-121), next level's segment table
        std::vector<int> next_node(2 * (size_t)n_split);
        std::vector<uint32_t> next_start(2 * (size_t)n_split + 1);
        next_start[0] = 0;
        for (int s = 0, k = 0; s < S; ++s) {
            if (h_info[s] >= 0) continue;
            const int node = seg_node[s];
            const double sv = h_split[s];
            double mn[3], mx[3], lmx[3], rmn[3];
            for (int a = 0; a < 3; ++a) { mn[a] = out.box[6 * (size_t)node + a]; mx[a] = out.box[6 * (size_t)node + 3 + a]; lmx[a] = mx[a]; rmn[a] = mn[a]; }
            lmx[axis] = sv; rmn[axis] = sv;
            const int li = add_node(mn, lmx);
            add_node(rmn, mx);
            out.axis[node] = axis; out.split[node] = sv; out.left[node] = li;
            next_node[2 * (size_t)k] = li; next_node[2 * (size_t)k + 1] = li + 1;
            next_start[2 * (size_t)k + 1] = next_start[2 * (size_t)k] + h_cnt[2 * (size_t)k];
            next_start[2 * (size_t)k + 2] = next_start[2 * (size_t)k + 1] + h_cnt[2 * (size_t)k + 1];
            ++k;
        }
        const long long n_next = next_start.back();
        // the sorted arrays live in [cur ^ 1]; the next level is written into [cur] (its old contents are dead) -- grow first
        KCK(list[cur].ensure((size_t)n_next)); KCK(seg[cur].ensure((size_t)n_next));
        KCK(cudaMemcpyAsync(child_start.p, next_start.data(), next_start.size() * 4, cudaMemcpyHostToDevice, st));
        kd_level_scatter<<<grid_for(n), 256, 0, st>>>(list[cur ^ 1].p, seg[cur ^ 1].p, seg_start.p, seg_info.p, le.p, gt.p, sl.p, sg.p, child_start.p, n,
                                                      list[cur].p, seg[cur].p);
        ++g_kd_build_launches;
        KCK(cudaGetLastError());
        KCK(cudaStreamSynchronize(st));   // h_start / h_info / next_start are reused below
        // make every per-level buffer large enough for the next level (list[cur] already holds it)
        n = n_next;
        KCK(ensure_level((size_t)n));
        seg_node.swap(next_node);
        h_start.swap(next_start);
    }
    out.polys.resize(final_n);
    if (final_n) KCK(cudaMemcpyAsync(out.polys.data(), final_list.p, final_n * 4, cudaMemcpyDeviceToHost, st));
    KCK(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace hare
