// ingest.cu -- Topology ingest on the GPU: Topology(min, max) + Add_Polygon x P + Finish_Topology()
// (Hare_Geometry_Topology.cs:85-91, 225-254, 342-377, 148-179; Polygon ctor normal Hare_Geometry_Polygons.cs:159-171).
// Same outputs, bit for bit, as host_build.cpp's topology_ingest() (tested identical).
//
// The reference welds vertices sequentially: a vertex is rounded to 15 digits, hashed to a 1 mm lattice cell and the FIRST
// vertex ever seen in a cell wins (AddGetIndex :342-377).  "First seen" is the smallest occurrence number 4*polygon + corner,
// so the sequential dictionary becomes: key every occurrence by its cell, stable-sort the (cell, occurrence) pairs
// (cub::DeviceRadixSort -- plumbing), and let every run of equal cells take the coordinates of its first member
// (head flags + inclusive max-scan of the head positions).  Normals and the padded bounds are per-polygon / per-head work.
//
// A vertex outside the declared bounds, a non-finite coordinate, or a lattice too large for a 64-bit cell key makes this
// path step aside (return 1): the caller then runs the host ingest, whose behaviour for such input is the reference's.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <string>

#include "host_build.hpp"

namespace hare {

namespace {

#define ICK(call)                                                                                            \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) { err = std::string("device ingest: ") + cudaGetErrorString(e_); rc = -2; goto done; } \
    } while (0)

// Math.Round(value, 15): scale by 1e15, round half to even, scale back (only below 1e16 in magnitude)
__device__ __forceinline__ double round_15_digits(double value) {
    if (fabs(value) < 1e16) {
        value *= 1e15;
        value = rint(value);
        value /= 1e15;
    }
    return value;
}

struct Lattice { double ms_min[3]; unsigned long long ydim, xytot; };

// one thread per vertex occurrence i = 4 * polygon + corner: its lattice-cell key (Point.Hash2, Primitives.cs:237-250)
__global__ void __launch_bounds__(256)
ingest_keys(const double* __restrict__ raw, const int32_t* __restrict__ vcount, long long P, Lattice L,
            unsigned long long* __restrict__ keys, uint32_t* __restrict__ occ, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * P) return;
    const long long p = i >> 2; const int k = (int)(i & 3);
    occ[i] = (uint32_t)i;
    const int n = vcount[p];
    if (n != 3 && n != 4) { *bad = 3; keys[i] = ~0ull; return; }
    if (k >= n) { keys[i] = ~0ull; return; }   // unused corner of a triangle: sorts to the end
    unsigned long long cell[3], sub[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double q = round_15_digits(raw[12 * p + 3 * k + a]);
        const double off = q - L.ms_min[a];
        if (!(off >= 0.0) || !(off < 4.0e6)) { *bad = 1; keys[i] = ~0ull; return; }
        const double fl = floor(off);
        cell[a] = (unsigned long long)fl;
        sub[a] = (unsigned long long)((off - (double)cell[a]) * 1000);
    }
    const unsigned long long bucket = L.xytot * cell[2] + L.ydim * cell[0] + cell[1];
    const unsigned long long pos = 1000000ull * sub[2] + 1000ull * sub[0] + sub[1];
    keys[i] = bucket * 1000000000ull + pos;   // pos < 1e9; the caller checked that bucket * 1e9 fits
}

__global__ void __launch_bounds__(256)
ingest_heads(const unsigned long long* __restrict__ skeys, long long n, uint32_t* __restrict__ head_pos) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) head_pos[j] = (j == 0 || skeys[j] != skeys[j - 1]) ? (uint32_t)j : 0u;
}

struct MaxOp { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } };

// IEEE double <-> unsigned key with the same order (for atomicMin / atomicMax on doubles)
__device__ __forceinline__ unsigned long long orderable(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// per sorted position j: the occurrence sidx[j] takes the coordinates of the first occurrence of its run; heads also
// feed Vertex_Count and the bounds of the welded vertex set (Finish_Topology :148-167)
__global__ void __launch_bounds__(256)
ingest_winners(const unsigned long long* __restrict__ skeys, const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ head_of, long long n,
               const double* __restrict__ raw, uint32_t* __restrict__ win, unsigned long long* __restrict__ stats /* nverts, lo[3], hi[3] */) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool head = false;
    double q[3] = { 0, 0, 0 };
    if (j < n && skeys[j] != ~0ull) {
        const uint32_t h = head_of[j];
        win[sidx[j]] = sidx[h];
        head = (h == (uint32_t)j);
        if (head) {
            const uint32_t i = sidx[j];
#pragma unroll
            for (int a = 0; a < 3; ++a) q[a] = round_15_digits(raw[12 * (long long)(i >> 2) + 3 * (i & 3) + a]);
        }
    }
    // warp-aggregate the head statistics
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (m == 0) return;
    unsigned long long lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = head ? orderable(q[a]) : ~0ull; hi[a] = head ? orderable(q[a]) : 0ull; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned long long l2 = __shfl_xor_sync(0xffffffffu, lo[a], off), h2 = __shfl_xor_sync(0xffffffffu, hi[a], off);
            lo[a] = l2 < lo[a] ? l2 : lo[a]; hi[a] = h2 > hi[a] ? h2 : hi[a];
        }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats, (unsigned long long)__popc(m));
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(stats + 1 + a, lo[a]); atomicMax(stats + 4 + a, hi[a]); }
    }
}

// per polygon: welded vertices (a triangle repeats vertex 2 in slot 3) and the Polygon ctor's unit normal
// (Hare_Geometry_Polygons.cs:159-171 with Hare_math.Cross(Vector,Vector) and Vector.Normalize)
__global__ void __launch_bounds__(256)
ingest_polygons(const double* __restrict__ raw, const int32_t* __restrict__ vcount, const uint32_t* __restrict__ win, long long P,
                double* __restrict__ verts_out, double* __restrict__ normals_out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int n = vcount[p];
    double V[4][3];
    for (int k = 0; k < n; ++k) {
        const uint32_t w = win[4 * p + k];
#pragma unroll
        for (int a = 0; a < 3; ++a) V[k][a] = round_15_digits(raw[12 * (long long)(w >> 2) + 3 * (w & 3) + a]);
    }
    if (n == 3) { V[3][0] = V[2][0]; V[3][1] = V[2][1]; V[3][2] = V[2][2]; }
    double N[3] = { 0, 0, 0 };
    const double ax = V[1][0] - V[0][0], ay = V[1][1] - V[0][1], az = V[1][2] - V[0][2];
    for (int j = 2; j < n; ++j) {
        const double bx = V[j][0] - V[0][0], by = V[j][1] - V[0][1], bz = V[j][2] - V[0][2];
        N[0] = ay * bz - az * by;
        N[1] = -(ax * bz - az * bx);
        N[2] = ax * by - ay * bx;
        const double len2 = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
        if (!(len2 < 4.9406564584124654e-324)) break;   // IsZeroVector uses double.Epsilon
    }
    double f = N[0] * N[0] + N[1] * N[1] + N[2] * N[2];
    if (f != 0) { f = sqrt(f); N[0] /= f; N[1] /= f; N[2] /= f; }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) verts_out[12 * p + 3 * k + a] = V[k][a];
    normals_out[3 * p] = N[0]; normals_out[3 * p + 1] = N[1]; normals_out[3 * p + 2] = N[2];
}

inline unsigned grid_for(long long n) { return (unsigned)((n + 255) / 256); }

inline double unorderable(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double x; memcpy(&x, &b, 8); return x;
}

}  // namespace

unsigned long long g_ingest_launches = 0;   // added to hare_launch_count() by hare_abi.cu

// returns 0 done, 1 "not for this path" (the caller runs the host ingest), -3 polygon with more than 4 sides, -2 CUDA error
int topology_ingest_gpu(int dev, const double* raw, const int32_t* vcount, int64_t P, const double minpt[3], const double maxpt[3],
                        double* verts_out, double* normals_out, double minmax_out[6], int64_t* vertex_count_out, std::string& err) {
    if (P <= 0 || 4 * P >= (1ll << 32)) return 1;
    Lattice L;
    int dim = 0;
    for (int a = 0; a < 3; ++a) {
        L.ms_min[a] = minpt[a] - 0.000000000001;
        const double mx = maxpt[a] + 0.000000000001;
        if (!(mx - L.ms_min[a] < 2.0e6)) return 1;
        dim = std::max(dim, (int)std::ceil(mx - L.ms_min[a]));
    }
    L.ydim = (unsigned long long)(long long)dim; L.xytot = (unsigned long long)((long long)dim * (long long)dim);
    // largest bucket the guarded kernel can form: off < 4e6 per axis
    const long double max_bucket = (long double)L.xytot * 4.0e6L + (long double)L.ydim * 4.0e6L + 4.0e6L;
    if (max_bucket * 1.0e9L >= 1.8e19L) return 1;

    int rc = 0;
    const long long n = 4 * P;
    double *d_raw = nullptr, *d_verts = nullptr, *d_normals = nullptr;
    int32_t* d_cnt = nullptr; int* d_bad = nullptr;
    unsigned long long *d_keys = nullptr, *d_skeys = nullptr, *d_stats = nullptr;
    uint32_t *d_occ = nullptr, *d_sidx = nullptr, *d_head = nullptr, *d_headof = nullptr, *d_win = nullptr;
    void* d_tmp = nullptr;
    size_t b1 = 0, b2 = 0;
    cudaStream_t st = nullptr;
    int h_bad = 0;
    unsigned long long h_stats[7];
    ICK(cudaSetDevice(dev));
    ICK(cudaStreamCreate(&st));
    ICK(cudaMalloc((void**)&d_raw, (size_t)P * 96)); ICK(cudaMalloc((void**)&d_verts, (size_t)P * 96)); ICK(cudaMalloc((void**)&d_normals, (size_t)P * 24));
    ICK(cudaMalloc((void**)&d_cnt, (size_t)P * 4)); ICK(cudaMalloc((void**)&d_bad, 4)); ICK(cudaMalloc((void**)&d_stats, 7 * 8));
    ICK(cudaMalloc((void**)&d_keys, (size_t)n * 8)); ICK(cudaMalloc((void**)&d_skeys, (size_t)n * 8));
    ICK(cudaMalloc((void**)&d_occ, (size_t)n * 4)); ICK(cudaMalloc((void**)&d_sidx, (size_t)n * 4)); ICK(cudaMalloc((void**)&d_head, (size_t)n * 4));
    ICK(cudaMalloc((void**)&d_headof, (size_t)n * 4)); ICK(cudaMalloc((void**)&d_win, (size_t)n * 4));
    cub::DeviceRadixSort::SortPairs(nullptr, b1, d_keys, d_skeys, d_occ, d_sidx, n, 0, 64, st);
    cub::DeviceScan::InclusiveScan(nullptr, b2, d_head, d_headof, MaxOp(), n, st);
    ICK(cudaMalloc(&d_tmp, std::max(b1, b2) + 256));
    ICK(cudaMemcpyAsync(d_raw, raw, (size_t)P * 96, cudaMemcpyHostToDevice, st));
    ICK(cudaMemcpyAsync(d_cnt, vcount, (size_t)P * 4, cudaMemcpyHostToDevice, st));
    ICK(cudaMemsetAsync(d_bad, 0, 4, st));
    {
        const unsigned long long init[7] = { 0ull, ~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull };
        ICK(cudaMemcpyAsync(d_stats, init, sizeof init, cudaMemcpyHostToDevice, st));
    }
    ingest_keys<<<grid_for(n), 256, 0, st>>>(d_raw, d_cnt, P, L, d_keys, d_occ, d_bad);
    {
        size_t bytes = std::max(b1, b2) + 256;
        ICK(cub::DeviceRadixSort::SortPairs(d_tmp, bytes, d_keys, d_skeys, d_occ, d_sidx, n, 0, 64, st));   // stable: runs keep ascending occurrence order
        ingest_heads<<<grid_for(n), 256, 0, st>>>(d_skeys, n, d_head);
        bytes = std::max(b1, b2) + 256;
        ICK(cub::DeviceScan::InclusiveScan(d_tmp, bytes, d_head, d_headof, MaxOp(), n, st));
    }
    ingest_winners<<<grid_for(n), 256, 0, st>>>(d_skeys, d_sidx, d_headof, n, d_raw, d_win, d_stats);
    ingest_polygons<<<grid_for(P), 256, 0, st>>>(d_raw, d_cnt, d_win, P, d_verts, d_normals);
    g_ingest_launches += 6;
    ICK(cudaGetLastError());
    ICK(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    ICK(cudaMemcpyAsync(h_stats, d_stats, sizeof h_stats, cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));
    if (h_bad == 3) { rc = -3; goto done; }
    if (h_bad) { rc = 1; goto done; }
    ICK(cudaMemcpyAsync(verts_out, d_verts, (size_t)P * 96, cudaMemcpyDeviceToHost, st));
    ICK(cudaMemcpyAsync(normals_out, d_normals, (size_t)P * 24, cudaMemcpyDeviceToHost, st));
    ICK(cudaStreamSynchronize(st));
    for (int a = 0; a < 3; ++a) {   // Finish_Topology :148-167
        minmax_out[a] = unorderable(h_stats[1 + a]) - 0.000000000001;
        minmax_out[3 + a] = unorderable(h_stats[4 + a]) + 0.000000000001;
    }
    if (vertex_count_out) *vertex_count_out = (int64_t)h_stats[0];
done:
    cudaFree(d_raw); cudaFree(d_verts); cudaFree(d_normals); cudaFree(d_cnt); cudaFree(d_bad); cudaFree(d_stats); cudaFree(d_keys); cudaFree(d_skeys);
    cudaFree(d_occ); cudaFree(d_sidx); cudaFree(d_head); cudaFree(d_headof); cudaFree(d_win); cudaFree(d_tmp);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // namespace hare
