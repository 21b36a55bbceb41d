// ray_bin.cuh -- coherence pre-pass of a Shoot batch: rays are handed to the traversal kernels grouped by (origin cell, direction
// cell) instead of in the caller's order.
//
// A Shoot's result does not depend on which other rays are in flight, and the kernels write events by ray number, so the order is
// free.  Rays that start in the same part of the model and point the same way walk the same nodes / voxels and test the same
// polygons: with 64 such rays in a warp's pool the node records, boxes and polygon records come out of L1 instead of L2 / HBM, and
// the slots move through the phases together (more lanes per instruction).  Callers hand over rays in whatever order their source
// loop produces (the BASELINE batches: sources round-robin, isotropic random directions -- no coherence at all).
//
// Counting sort on an 18..25-bit key, three small kernels: key + histogram, exclusive scan (scan_u32), scatter through per-bucket cursors.
// The order inside a bucket is arbitrary (atomics) -- it changes no result.  FP32 arithmetic: the key only steers scheduling.
#pragma once
#include <cstdint>
#include "hare_math.cuh"

namespace hare {

struct RayBinGeom { float ox, oy, oz, sx, sy, sz; int dirbits; };   // origin cell = clamp((o - (ox,oy,oz)) * (sx,sy,sz), 0, 3): 4 x 4 x 4 cells over the model's bounds

// dirbits = bits per cube-face coordinate (6 x 4^dirbits direction cells), chosen per batch so that there are about as many buckets as
// rays: 4 (65 k rays) .. 8 (>= 20 M rays; measured on C3, 20 M rays: 5 bits 582, 6 615, 7 645, 8 657 Mrays/s, unsorted 560)
HD uint32_t ray_bin_buckets(int dirbits) { return 64u * 6u * (1u << (2 * dirbits)); }
HD int ray_bin_dirbits(long long n) { int b = 4; while (b < 8 && (long long)ray_bin_buckets(b) < n) ++b; return b; }

HD uint32_t ray_bin_key(const double* __restrict__ o, const double* __restrict__ d, const RayBinGeom& g) {
    const float dx = (float)d[0], dy = (float)d[1], dz = (float)d[2];
    const float ax = fabsf(dx), ay = fabsf(dy), az = fabsf(dz);
    int f = 0; float m = ax;
    if (ay > m) { f = 1; m = ay; }
    if (az > m) { f = 2; m = az; }
    if (!(m > 0.0f) || !(m < 3.0e38f)) return 0u;                     // zero / NaN / Inf direction: any bucket will do
    const float w = f == 0 ? dx : (f == 1 ? dy : dz);
    const float u = (f == 0 ? dy : (f == 1 ? dz : dx)) / m, v = (f == 0 ? dz : (f == 1 ? dx : dy)) / m;   // in [-1, 1]
    const int nd = 1 << g.dirbits;
    int ui = (int)((u + 1.0f) * (0.5f * nd)), vi = (int)((v + 1.0f) * (0.5f * nd));
    ui = ui < 0 ? 0 : (ui > nd - 1 ? nd - 1 : ui); vi = vi < 0 ? 0 : (vi > nd - 1 ? nd - 1 : vi);
    uint32_t mort = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) mort |= (((uint32_t)ui >> b) & 1u) << (2 * b) | (((uint32_t)vi >> b) & 1u) << (2 * b + 1);
    float cx = ((float)o[0] - g.ox) * g.sx, cy = ((float)o[1] - g.oy) * g.sy, cz = ((float)o[2] - g.oz) * g.sz;
    cx = cx >= 0.0f ? cx : 0.0f; cy = cy >= 0.0f ? cy : 0.0f; cz = cz >= 0.0f ? cz : 0.0f;     // (NaN -> 0)
    const uint32_t ix = cx < 3.0f ? (uint32_t)cx : 3u, iy = cy < 3.0f ? (uint32_t)cy : 3u, iz = cz < 3.0f ? (uint32_t)cz : 3u;
    const uint32_t cell = (ix << 4) | (iy << 2) | iz;
    return ((cell * 6u + (uint32_t)f * 2u + (w < 0.0f ? 1u : 0u)) << (2 * g.dirbits)) | mort;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256)
ray_bin_count(const double* __restrict__ o, const double* __restrict__ d, long long N, const RayBinGeom g, uint32_t* __restrict__ keys, uint32_t* __restrict__ counts) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t k = ray_bin_key(o + 3 * i, d + 3 * i, g);
        keys[i] = k;
        atomicAdd(counts + k, 1u);
    }
}

__global__ void __launch_bounds__(256)
ray_bin_scatter(const uint32_t* __restrict__ keys, long long N, uint32_t* __restrict__ cursor /* exclusive offsets, consumed */, uint32_t* __restrict__ perm) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        perm[atomicAdd(cursor + keys[i], 1u)] = (uint32_t)i;
}
#endif

}  // namespace hare
