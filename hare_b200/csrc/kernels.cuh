// kernels.cuh -- __global__ entry points of the build path (sm_100a); the traversal kernels live in vg_wave.cuh / oct_wave.cuh /
// kd_wave.cuh.
//   vg_* kernels          K2: Voxel_Grid cell-list build (count -> scan -> scatter -> per-cell sort)
//   oct_* kernels         GPU Octree build (level-synchronous SAT masks, scan, order-preserving scatter)
#pragma once
#include <cstdint>
#include "hare_math.cuh"
#include "sat.cuh"
#include "shoot.cuh"

namespace hare {

// ---------------------------------------------------------------------------------------
// K2: Voxel_Grid build.  Voxel_Grid(Model, Domain) tests every voxel against every polygon
// (Voxel_Grid.cs:273-304).  Here each polygon enumerates only the voxels whose inflated box
// can touch the polygon's bounding box (a superset of the voxels the SAT's box-axis tests
// accept) and runs the same predicate on each; lists end up ascending like the reference's.
// ---------------------------------------------------------------------------------------
struct VGBuild {
    double ominx, ominy, ominz, vdx, vdy, vdz;
    int nx, ny, nz;
};

__device__ __forceinline__ void candidate_range(double mn, double mx, double omin, double vd, int n, int& lo, int& hi) {
    // voxel X spans [X*vd - eps + omin, (X+1)*vd + eps + omin]; 1e-9 of a voxel of slack keeps this a superset
    double a = floor((mn - HARE_EPS - omin) / vd - 1e-9) ;
    double b = floor((mx + HARE_EPS - omin) / vd + 1e-9);
    a = fmax(a, 0.0); b = fmin(b, (double)(n - 1));
    lo = (int)a; hi = (int)b;
    if (!(a <= b)) { lo = 0; hi = -1; }
}

// MODE 0: count per cell; MODE 1: scatter into cell_poly via per-cell cursors.
// Eight lanes per polygon (four polygons per warp): a polygon of a tessellated hall touches 2-12 voxels, so
// a whole warp per polygon would leave three quarters of the lanes idle; the eight lanes stride over the
// polygon's candidate voxels and read its 128-byte record as one broadcast line.
template <int MODE>
__global__ void __launch_bounds__(256)
vg_bin_kernel(const VGBuild g, const PolyRec* __restrict__ polys, long long P,
              uint32_t* __restrict__ cell_count, const uint32_t* __restrict__ cell_offset,
              uint32_t* __restrict__ cursor, uint32_t* __restrict__ cell_poly) {
    const int lane = threadIdx.x & 7;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 3;
    for (long long p = warp; p < P; p += nwarps) {
        double V[16];
        load_poly(polys, (uint32_t)p, V);
        const int n = (V[15] == 4.0) ? 4 : 3;
        int lo[3], hi[3];
        {
            double mnx = fmin(fmin(V[0], V[3]), V[6]), mxx = fmax(fmax(V[0], V[3]), V[6]);
            double mny = fmin(fmin(V[1], V[4]), V[7]), mxy = fmax(fmax(V[1], V[4]), V[7]);
            double mnz = fmin(fmin(V[2], V[5]), V[8]), mxz = fmax(fmax(V[2], V[5]), V[8]);
            if (n == 4) {
                mnx = fmin(mnx, V[9]); mxx = fmax(mxx, V[9]); mny = fmin(mny, V[10]); mxy = fmax(mxy, V[10]);
                mnz = fmin(mnz, V[11]); mxz = fmax(mxz, V[11]);
            }
            candidate_range(mnx, mxx, g.ominx, g.vdx, g.nx, lo[0], hi[0]);
            candidate_range(mny, mxy, g.ominy, g.vdy, g.ny, lo[1], hi[1]);
            candidate_range(mnz, mxz, g.ominz, g.vdz, g.nz, lo[2], hi[2]);
        }
        const int ex = hi[0] - lo[0] + 1, ey = hi[1] - lo[1] + 1, ez = hi[2] - lo[2] + 1;
        if (ex <= 0 || ey <= 0 || ez <= 0) continue;
        const long long total = (long long)ex * ey * ez;
        for (long long q = lane; q < total; q += 8) {
            const int z = lo[2] + (int)(q % ez);
            const int y = lo[1] + (int)((q / ez) % ey);
            const int x = lo[0] + (int)(q / ((long long)ez * ey));
            // voxel box exactly as Fill_Voxels builds it (Voxel_Grid.cs:283-285)
            const Box3 B = make_box(((double)x * g.vdx - HARE_EPS) + g.ominx, ((double)y * g.vdy - HARE_EPS) + g.ominy,
                                    ((double)z * g.vdz - HARE_EPS) + g.ominz,
                                    ((double)(x + 1) * g.vdx + HARE_EPS) + g.ominx, ((double)(y + 1) * g.vdy + HARE_EPS) + g.ominy,
                                    ((double)(z + 1) * g.vdz + HARE_EPS) + g.ominz);
            if (poly_box_overlap(B, V, n)) {
                const uint32_t ci = ((uint32_t)x * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nz + (uint32_t)z;
                if (MODE == 0) atomicAdd(cell_count + ci, 1u);
                else cell_poly[cell_offset[ci] + atomicAdd(cursor + ci, 1u)] = (uint32_t)p;
            }
        }
    }
}

// Level step of the hierarchical constructor Voxel_Grid(Model, MaxDomain, Avg_polys) (Voxel_Grid.cs:161-253): a child voxel tests only
// the polygons of its parent's list (:207-215), so every (parent voxel, polygon) pair of the previous level runs PolyBoxOverlap on
// the parent's eight children -- one thread per (pair, child), the eight lanes of a pair share the parent look-up and the record.
//   PASS 0: SAT -> one mask byte per pair + the child voxels' list lengths.   PASS 1: scatter from the stored masks.
// g = the CHILD grid (counts 2x the parent's); lists come out unordered and are sorted by vg_finish_cells like the flat build's.
template <int PASS>
__global__ void __launch_bounds__(256)
vg_refine_kernel(const VGBuild g, const PolyRec* __restrict__ polys, const uint32_t* __restrict__ parent_offset,
                 const uint32_t* __restrict__ parent_poly, long long npairs, long long ncells_parent, uint8_t* __restrict__ mask,
                 uint32_t* __restrict__ cell_count, const uint32_t* __restrict__ cell_offset, uint32_t* __restrict__ cursor,
                 uint32_t* __restrict__ cell_poly) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pair = tid >> 3;
    const int ch = (int)(tid & 7), lane = threadIdx.x & 31, lead = lane & ~7;
    const bool valid = pair < npairs;
    long long cell = 0;
    if (valid && lane == lead) {       // parent voxel of this pair: the last c with parent_offset[c] <= pair
        long long lo = 0, hi = ncells_parent;          // invariant: parent_offset[lo] <= pair < parent_offset[hi]
        while (hi - lo > 1) { const long long mid = (lo + hi) >> 1; if ((long long)parent_offset[mid] <= pair) lo = mid; else hi = mid; }
        cell = lo;
    }
    cell = __shfl_sync(0xffffffffu, cell, lead);
    const int pnz = g.nz >> 1, pny = g.ny >> 1;
    const int Z = (int)(cell % pnz), Y = (int)((cell / pnz) % pny), X = (int)(cell / ((long long)pnz * pny));
    const int x = 2 * X + ((ch >> 2) & 1), y = 2 * Y + ((ch >> 1) & 1), z = 2 * Z + (ch & 1);
    const uint32_t ci = ((uint32_t)x * (uint32_t)g.ny + (uint32_t)y) * (uint32_t)g.nz + (uint32_t)z;
    const uint32_t poly = valid ? parent_poly[pair] : 0u;
    if (PASS == 0) {
        bool hit = false;
        if (valid) {
            double V[16];
            load_poly(polys, poly, V);
            const int n = (V[15] == 4.0) ? 4 : 3;
            // voxel box exactly as the constructor builds it (Voxel_Grid.cs:198-201): (x * VoxelDims - Epsilon) + OBox.Min ...
            const Box3 B = make_box(((double)x * g.vdx - HARE_EPS) + g.ominx, ((double)y * g.vdy - HARE_EPS) + g.ominy,
                                    ((double)z * g.vdz - HARE_EPS) + g.ominz,
                                    ((double)(x + 1) * g.vdx + HARE_EPS) + g.ominx, ((double)(y + 1) * g.vdy + HARE_EPS) + g.ominy,
                                    ((double)(z + 1) * g.vdz + HARE_EPS) + g.ominz);
            hit = poly_box_overlap(B, V, n);
            if (hit) atomicAdd(cell_count + ci, 1u);
        }
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (valid && lane == lead) mask[pair] = (uint8_t)((b >> lead) & 0xffu);
    } else {
        if (valid && ((mask[pair] >> ch) & 1u)) cell_poly[cell_offset[ci] + atomicAdd(cursor + ci, 1u)] = poly;
    }
}

// number of non-empty voxels (the stop rule's denominator, Voxel_Grid.cs:251-252)
__global__ void __launch_bounds__(256) vg_count_nonempty(const uint32_t* __restrict__ cell_count, long long ncells, unsigned long long* __restrict__ out) {
    unsigned int s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncells; i += (long long)gridDim.x * blockDim.x) s += cell_count[i] ? 1u : 0u;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(out, (unsigned long long)s);
}

__global__ void __launch_bounds__(256) fill_iota(uint32_t* __restrict__ a, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] = (uint32_t)i;
}

// Exclusive scan of n uint32 in three passes (tile sums, scan of tile sums, tile scan + offset).
#define HARE_SCAN_TILE 4096   /* 1024 threads x 4 items */

__device__ __forceinline__ uint32_t block_exclusive_scan_1024(uint32_t v, uint32_t* warp_sums, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += y; }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = warp_sums[lane], si = s;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, si, off); if (lane >= off) si += y; }
        warp_sums[lane] = si - s;          // exclusive per-warp base
        if (lane == 31) warp_sums[32] = si;  // block total
    }
    __syncthreads();
    block_total = warp_sums[32];
    return inc - v + warp_sums[w];
}

__global__ void __launch_bounds__(1024) scan_tile_sums(const uint32_t* __restrict__ in, long long n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t ws[33];
    const long long base = (long long)blockIdx.x * HARE_SCAN_TILE + threadIdx.x * 4;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (base + k < n) s += in[base + k];
    uint32_t tot;
    block_exclusive_scan_1024(s, ws, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: exclusive scan of m tile sums in place (m arbitrary, processed 1024 at a time)
__global__ void __launch_bounds__(1024) scan_tile_offsets(uint32_t* __restrict__ tile_sums, long long m, uint32_t* __restrict__ grand_total) {
    __shared__ uint32_t ws[33];
    uint32_t carry = 0;
    for (long long b = 0; b < m; b += 1024) {
        const long long i = b + threadIdx.x;
        const uint32_t v = (i < m) ? tile_sums[i] : 0u;
        uint32_t tot;
        const uint32_t ex = block_exclusive_scan_1024(v, ws, tot);
        if (i < m) tile_sums[i] = ex + carry;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

// out[i] = exclusive prefix; out[n] = total is written by the host wrapper from grand_total.
__global__ void __launch_bounds__(1024) scan_tiles(const uint32_t* __restrict__ in, long long n, const uint32_t* __restrict__ tile_offsets,
                                                  uint32_t* __restrict__ out) {
    __shared__ uint32_t ws[33];
    const long long base = (long long)blockIdx.x * HARE_SCAN_TILE + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan_1024(s, ws, tot) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
}

// Sort each cell's list ascending (the scatter order is arbitrary), and emit the packed
// (offset, count) header plus the occupancy bit.  One thread per cell; lists are short.
__global__ void __launch_bounds__(256)
vg_finish_cells(const uint32_t* __restrict__ cell_offset, const uint32_t* __restrict__ cell_count, long long ncells,
                uint32_t* __restrict__ cell_poly, uint2* __restrict__ cells, uint32_t* __restrict__ occ) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (c < ncells) {
        const uint32_t off = cell_offset[c];
        cnt = cell_count[c];
        uint32_t* L = cell_poly + off;
        if (cnt <= 48) {
            for (uint32_t i = 1; i < cnt; ++i) {
                const uint32_t key = L[i]; uint32_t j = i;
                while (j > 0 && L[j - 1] > key) { L[j] = L[j - 1]; --j; }
                L[j] = key;
            }
        } else {   // heap sort in place
            auto sift = [&](uint32_t start, uint32_t end) {
                uint32_t root = start;
                while (2 * root + 1 < end) {
                    uint32_t ch = 2 * root + 1;
                    if (ch + 1 < end && L[ch] < L[ch + 1]) ++ch;
                    if (L[root] < L[ch]) { uint32_t tmp = L[root]; L[root] = L[ch]; L[ch] = tmp; root = ch; } else break;
                }
            };
            for (int s = (int)(cnt / 2) - 1; s >= 0; --s) sift((uint32_t)s, cnt);
            for (uint32_t e = cnt - 1; e > 0; --e) { uint32_t tmp = L[0]; L[0] = L[e]; L[e] = tmp; sift(0, e); }
        }
        cells[c] = make_uint2(off, cnt);
    }
    const unsigned int bits = __ballot_sync(0xffffffffu, cnt > 0);
    if ((threadIdx.x & 31) == 0 && c < ncells) occ[c >> 5] = bits;
}

// ---------------------------------------------------------------------------------------
// GPU Octree build ("Octree - alt.cs":91-138), one level at a time.  The host keeps the node table and decides
// leaf / split (:93); the device does the O(entries x 8) SAT work and the order-preserving distribution:
//   oct_fill_slot    every entry of a node that splits learns its slot in this level's split table
//   oct_mask_kernel  entry x child PolyBoxOverlap against the child boxes the host computed (:99-114); flags are
//                    stored child-major (flag[c*E + e]) so that ONE exclusive scan ranks every child's entries in
//                    parent-list order (:118-130)
//   oct_child_counts per (split node, child): number of entries = difference of two scan values
//   oct_scatter      entry -> position in the next level's entry array (children laid out in (node, child) order)
//   oct_copy_lists   leaf lists -> final list array
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
oct_fill_slot(const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_cnt, int nslots, uint32_t* __restrict__ entry_slot) {
    for (int s = blockIdx.x; s < nslots; s += gridDim.x) {
        const uint32_t b = seg_start[s], n = seg_cnt[s];
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) entry_slot[b + k] = (uint32_t)s;
    }
}

__global__ void __launch_bounds__(256)
oct_mask_kernel(const PolyRec* __restrict__ polys, const uint32_t* __restrict__ entry_poly, const uint32_t* __restrict__ entry_slot,
                const double* __restrict__ child_box /* nslots x 8 x 6 */, long long E, uint32_t* __restrict__ flags /* 8 x E */,
                unsigned long long* __restrict__ lost) {
    // 8 threads per entry: thread c tests child c
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gid >> 3; const int c = (int)(gid & 7);
    bool ov = false;
    if (e < E) {
        const uint32_t s = entry_slot[e];
        if (s != 0xffffffffu) {
            double V[16];
            load_poly(polys, entry_poly[e], V);
            const double* b = child_box + ((size_t)s * 8 + c) * 6;
            const Box3 B = make_box(b[0], b[1], b[2], b[3], b[4], b[5]);
            ov = poly_box_overlap(B, V, (V[15] == 4.0) ? 4 : 3);
            flags[(long long)c * E + e] = ov ? 1u : 0u;
        } else {
            flags[(long long)c * E + e] = 0u;
        }
    }
    // polygons overlapping no child are dropped (:116,129): count them
    const unsigned m = __ballot_sync(0xffffffffu, ov);
    if (e < E && c == 0 && entry_slot[e] != 0xffffffffu) {
        const int sh = (threadIdx.x & 31) & ~7;
        if (((m >> sh) & 0xffu) == 0) atomicAdd(lost, 1ull);
    }
}

__global__ void __launch_bounds__(256)
oct_child_counts(const uint32_t* __restrict__ scan /* 8E + 1 */, const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ seg_cnt,
                 int nslots, long long E, uint32_t* __restrict__ counts /* nslots x 8 */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslots * 8) return;
    const int s = i >> 3, c = i & 7;
    const long long b = (long long)c * E + seg_start[s];
    counts[i] = scan[b + seg_cnt[s]] - scan[b];
}

__global__ void __launch_bounds__(256)
oct_scatter(const uint32_t* __restrict__ entry_poly, const uint32_t* __restrict__ entry_slot, const uint32_t* __restrict__ flags,
            const uint32_t* __restrict__ scan, const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ child_base /* nslots x 8 */,
            long long E, uint32_t* __restrict__ next_poly) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = gid >> 3; const int c = (int)(gid & 7);
    if (e >= E) return;
    const uint32_t s = entry_slot[e];
    if (s == 0xffffffffu) return;
    const long long idx = (long long)c * E + e;
    if (flags[idx]) next_poly[child_base[(size_t)s * 8 + c] + (scan[idx] - scan[(long long)c * E + seg_start[s]])] = entry_poly[e];
}

__global__ void __launch_bounds__(256)
oct_copy_lists(const uint32_t* __restrict__ src, const uint32_t* __restrict__ src_start, const uint32_t* __restrict__ dst_start,
               const uint32_t* __restrict__ cnt, int nleaves, uint32_t* __restrict__ dst) {
    for (int l = blockIdx.x; l < nleaves; l += gridDim.x) {
        const uint32_t a = src_start[l], b = dst_start[l], n = cnt[l];
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) dst[b + k] = src[a + k];
    }
}

// per list entry (cell_poly order): the polygon's padded FP32 bounding box with its id riding in lo.w (VGrid::lbox, cull_box)
__global__ void __launch_bounds__(256)
vg_gather_list_box(const uint32_t* __restrict__ cell_poly, const PolyRec* __restrict__ polys, uint32_t total, float4* __restrict__ lbox) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const uint32_t i = cell_poly[k];
    double P[16];
    load_poly(polys, i, P);   // a triangle repeats vertex 2 in slot 3: min/max over the four slots is its box
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double l = fmin(fmin(P[a], P[3 + a]), fmin(P[6 + a], P[9 + a])), h = fmax(fmax(P[a], P[3 + a]), fmax(P[6 + a], P[9 + a]));
        const double pad = hare_box_pad(l, h);
        lo[a] = __double2float_rd(l - pad); hi[a] = __double2float_ru(h + pad);
    }
    lbox[2 * (size_t)k] = make_float4(lo[0], lo[1], lo[2], __uint_as_float(i));
    lbox[2 * (size_t)k + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// per polygon: padded FP32 bounding box (lo, hi), indexed by polygon id (tree kernels, HARE_*_ENTRY_PBOX)
__global__ void __launch_bounds__(256)
poly_box_table(const PolyRec* __restrict__ polys, uint32_t P, float4* __restrict__ pbox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    double V[16];
    load_poly(polys, i, V);
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double l = fmin(fmin(V[a], V[3 + a]), fmin(V[6 + a], V[9 + a])), h = fmax(fmax(V[a], V[3 + a]), fmax(V[6 + a], V[9 + a]));
        const double pad = hare_box_pad(l, h);
        lo[a] = __double2float_rd(l - pad); hi[a] = __double2float_ru(h + pad);
    }
    pbox[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
    pbox[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// occupancy bitmap of the grid padded by one voxel on every side: bit = list non-empty, or border voxel (vg_wave.cuh)
__global__ void __launch_bounds__(256)
vg_pad_occupancy(const uint32_t* __restrict__ occ, int nx, int ny, int nz, uint32_t* __restrict__ occp) {
    const uint32_t py = (uint32_t)ny + 2u, pz = (uint32_t)nz + 2u;
    const uint32_t total = ((uint32_t)nx + 2u) * py * pz;
    const uint32_t cp = blockIdx.x * blockDim.x + threadIdx.x;
    bool bit = false;
    if (cp < total) {
        const uint32_t xp = cp / (py * pz), r = cp - xp * py * pz, yp = r / pz, zp = r - yp * pz;
        if (xp == 0 || xp == (uint32_t)nx + 1u || yp == 0 || yp == (uint32_t)ny + 1u || zp == 0 || zp == (uint32_t)nz + 1u) bit = true;
        else {
            const uint32_t ci = ((xp - 1u) * (uint32_t)ny + (yp - 1u)) * (uint32_t)nz + (zp - 1u);
            bit = (occ[ci >> 5] >> (ci & 31)) & 1u;
        }
    }
    const unsigned int bits = __ballot_sync(0xffffffffu, bit);
    if ((threadIdx.x & 31) == 0 && cp < total) occp[cp >> 5] = bits;
}

// host-uploaded CSR -> packed headers + occupancy
__global__ void __launch_bounds__(256)
vg_pack_cells(const uint32_t* __restrict__ cell_offset, long long ncells, uint2* __restrict__ cells, uint32_t* __restrict__ occ) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (c < ncells) {
        const uint32_t off = cell_offset[c];
        cnt = cell_offset[c + 1] - off;
        cells[c] = make_uint2(off, cnt);
    }
    const unsigned int bits = __ballot_sync(0xffffffffu, cnt > 0);
    if ((threadIdx.x & 31) == 0 && c < ncells) occ[c >> 5] = bits;
}

}  // namespace hare
