/* hare_b200.h -- C ABI of libhare_b200.so, the B200 (sm_100a) implementation of
 * Hare's batched closest-hit Spatial_Partition.Shoot.
 *
 * This is the drop-in boundary.  The reference
 * (PachydermAcoustic/Hare, C#) has no FFI of its own: these entry points are
 * what a [DllImport("hare_b200")] block in the Hare_NC build binds (see
 * INTEGRATION.md and hare_b200/csharp/Hare_B200.cs).  Each function cites the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions: plain pointers and sizes only; all pointer arguments are HOST
 * memory unless the name ends in _device; the caller owns every host buffer,
 * the library owns all device memory behind the opaque handles; return 0 on
 * success, a negative hare_status otherwise (text via hare_last_error()).
 * Calls on one handle are serialised on that handle's stream; different
 * handles may be used from different threads.  There is no CPU fallback:
 * every compute entry point fails with HARE_ERR_CUDA when no sm_100 device is
 * usable.
 */
#ifndef HARE_B200_H
#define HARE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hare_topo_s* hare_topo_t;   /* flattened Topology (device) */
typedef struct hare_part_s* hare_part_t;   /* flattened Spatial_Partition (device) */

enum hare_status {
    HARE_OK = 0,
    HARE_ERR_INVALID = -1,        /* bad argument */
    HARE_ERR_CUDA = -2,           /* CUDA runtime error / no device */
    HARE_ERR_UNSUPPORTED = -3,    /* e.g. polygon with more than 4 sides: Hare_Geometry_Topology.cs:245-248 */
    HARE_ERR_NOMEM = -4
};

/* poly_id values written by the shoot entry points */
#define HARE_MISS (-1)            /* X_Event(): Hit=false, Poly_id=-1  (Hare_Geometry_Primitives.cs:454-462) */
#define HARE_RAY_FAULT (-2)       /* the reference throws IndexOutOfRangeException for this ray (Voxel_Grid.cs:374-383) */
#define HARE_NOT_SHOT (-3)        /* reflect_chain event slot after the chain ended */

enum hare_part_kind { HARE_VOXEL_GRID = 1, HARE_OCTREE = 2, HARE_KDTREE = 3 };

/* counters[] layout (uint64 x 4), filled when the pointer is non-NULL */
enum { HARE_CNT_CELLS = 0, HARE_CNT_ENTRIES = 1, HARE_CNT_TESTS = 2, HARE_CNT_HITS = 3, HARE_CNT_N = 4 };

/* ---- library ----------------------------------------------------------- */
const char* hare_version(void);
const char* hare_last_error(void);                 /* thread-local text of the last failure */
int hare_device_count(void);
/* Select the CUDA devices this process uses (device_ids == NULL: device 0 ... n-1;
 * n_devices == 0: the current device only).  Geometry created afterwards is
 * replicated on every selected device and ray batches are block-sharded over them.
 * n_devices == -1 selects host-only handles for build-time tooling on a machine without a
 * GPU: ingest, the tree builders, *_info and *_download work; hare_voxelgrid_build and every
 * Shoot / reflect call fail with HARE_ERR_CUDA. */
int hare_init(const int* device_ids, int n_devices);

/* ---- Topology ---------------------------------------------------------- */
/* Host-side restatement of Topology(Point min, Point max) + Add_Polygon xP + Finish_Topology()
 * (Hare_Geometry_Topology.cs:85-91, 225-254, 342-377, 148-179; Polygon ctor normal
 * Hare_Geometry_Polygons.cs:159-171) for hosts that do not run Hare's own C# Topology:
 * rounds to 15 digits, welds vertices sharing a 1 mm cell, computes unit normals and the
 * padded bounds.  raw_verts is P x 4 x 3 (a triangle ignores slot 3), vcount[i] in {3,4}.
 * Outputs: verts_out P x 4 x 3 (triangles repeat vertex 2 in slot 3), normals_out P x 3,
 * minmax_out = Topology.Min xyz, Topology.Max xyz, *vertex_count_out = Vertex_Count. */
int hare_topology_ingest(const double* raw_verts, const int32_t* vcount, int64_t P,
                         const double minpt[3], const double maxpt[3],
                         double* verts_out, double* normals_out, double minmax_out[6],
                         int64_t* vertex_count_out);

/* Flatten an already-built Topology: verts = Polys[i].Points (P x 4 x 3), normals = Polys[i].Normal,
 * vcount = Polys[i].VertextCT, minmax = Topology.Min / .Max (Hare_Geometry_Topology.cs:469-477). */
int hare_topology_create(const double* verts, const double* normals, const int32_t* vcount, int64_t P,
                         const double minmax[6], hare_topo_t* out);
int64_t hare_topology_polygon_count(hare_topo_t topo);
/* Lifetime: a partition reads its Topology's device records for as long as it lives.  Destroy partitions first;
 * hare_topology_destroy fails with HARE_ERR_INVALID (and frees nothing) while any partition built on the handle is alive. */
int hare_topology_destroy(hare_topo_t topo);

/* ---- Voxel_Grid -------------------------------------------------------- */
/* new Voxel_Grid(Model, Domain)  (Voxel_Grid.cs:48-121 + Fill_Voxels :273-304), single topology.
 * Cell lists are built on the GPU (count / scan / scatter / per-cell sort) and are
 * identical to the reference's ascending lists. */
int hare_voxelgrid_build(hare_topo_t topo, int domain, hare_part_t* out);
/* new Voxel_Grid(Model, MaxDomain, Avg_polys)  (Voxel_Grid.cs:128-254): the grid is refined 2x per axis per
 * level, up to 2^MaxDomain voxels per axis, stopping after level k > 1 once the mean list length of the
 * non-empty voxels drops below Avg_polys (:252).  Built the reference's way on the GPU: a child voxel tests
 * only the polygons of its parent's list (:207-215), level by level from the single voxel that lists every
 * polygon (CSR equality with the oracle's hierarchical build is tested up to 256^3 / 2M polygons). */
int hare_voxelgrid_build_adaptive(hare_topo_t topo, int max_domain_log2, int avg_polys, hare_part_t* out);
/* Upload a grid built by the host (e.g. Hare's hierarchical ctor, Voxel_Grid.cs:128-254):
 * obox = OBox.Min xyz, OBox.Max xyz; ct = VoxelCtX/Y/Z; CSR lists, cell index ((x*Ny+y)*Nz+z).
 * Checked: cell_offset[0] == 0, non-decreasing offsets, polygon indices < P (HARE_ERR_INVALID otherwise). */
int hare_voxelgrid_upload(hare_topo_t topo, const double obox[6], const int32_t ct[3],
                          const uint32_t* cell_offset, const uint32_t* cell_poly, hare_part_t* out);
int hare_voxelgrid_info(hare_part_t part, double obox[6], double voxeldims[3], int32_t ct[3], int64_t* npairs);
int hare_voxelgrid_download(hare_part_t part, uint32_t* cell_offset, uint32_t* cell_poly);

/* ---- Octree ("Octree - alt.cs") ----------------------------------------- */
/* new Octree(Model, maxDepth, maxPolygonsPerNode)  (:45-138). */
int hare_octree_build(hare_topo_t topo, int maxDepth, int maxPolygonsPerNode, hare_part_t* out);
/* Upload a host-built tree: node_box N x 6 (Min xyz, Max xyz); first_child[i] = index of child 0
 * (children are 8 consecutive nodes) or -1 for a leaf; leaf lists are polys[list_off, +list_cnt).
 * Checked (also for files read by hare_part_load): children follow their parent (i < first_child[i]), every node but the root has
 * exactly one parent (no cycles, no shared subtrees, no orphans), depth < 20, list ranges and polygon indices in bounds. */
int hare_octree_upload(hare_topo_t topo, const double* node_box, const int32_t* first_child,
                       const uint32_t* list_off, const uint32_t* list_cnt, const uint32_t* polys,
                       int64_t n_nodes, int64_t n_list, hare_part_t* out);
int hare_octree_info(hare_part_t part, int64_t* n_nodes, int64_t* n_list, int64_t* lost_polys, int32_t* depth);
int hare_octree_download(hare_part_t part, double* node_box, int32_t* first_child,
                         uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys);

/* ---- KDTree (KDTree.cs) -------------------------------------------------- */
/* new KDTree(Model, maxDepth, maxPolygonsPerNode)  (:51-139). */
int hare_kdtree_build(hare_topo_t topo, int maxDepth, int maxPolygonsPerNode, hare_part_t* out);
/* Upload: node_box N x 6; split[i], axis[i] (-1 leaf); left[i] (right child = left[i] + 1). */
int hare_kdtree_upload(hare_topo_t topo, const double* node_box, const double* split, const int32_t* axis,
                       const int32_t* left, const uint32_t* list_off, const uint32_t* list_cnt,
                       const uint32_t* polys, int64_t n_nodes, int64_t n_list, hare_part_t* out);
int hare_kdtree_info(hare_part_t part, int64_t* n_nodes, int64_t* n_list, int32_t* depth);
int hare_kdtree_download(hare_part_t part, double* node_box, double* split, int32_t* axis, int32_t* left,
                         uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys);

/* On-disk form of a flattened partition (the reference has none): a large hall can skip its rebuild.
 * The file records the Topology it was built for (polygon count + a hash of the vertices); hare_part_load refuses another
 * one and otherwise applies the checks and device set-up of the *_upload entry points. */
int hare_part_save(hare_part_t part, const char* path);
int hare_part_load(hare_topo_t topo, const char* path, hare_part_t* out);

/* Device time of the build's kernels (CUDA events) and host wall time of the constructor call, in ms (Voxel_Grid builds; 0 otherwise). */
int hare_part_build_ms(hare_part_t part, double* kernel_ms, double* wall_ms);
int hare_part_kind(hare_part_t part);
int64_t hare_part_device_bytes(hare_part_t part);
int hare_part_destroy(hare_part_t part);

/* ---- Shoot --------------------------------------------------------------- */
/* Batched Spatial_Partition.Shoot(Ray, top_index = 0, out X_Event, poly_origin1, poly_origin2)
 * (Spatial_Partition.cs:32-33; Voxel_Grid.cs:351-552; "Octree - alt.cs":159-284; KDTree.cs:198-361).
 *   o, d        N x 3 ray origins / directions (Ray.x..z, Ray.dx..dz)
 *   origin1/2   N polygon indices to skip, or NULL (= -1: the 3-argument overload)
 *   ray_id      N Ray.Ray_ID values or NULL.  Only its zero-ness matters.  The reference asks for
 *               unique non-zero ids; its mailboxes (Voxel_Grid.cs:54-62, 478-480) start at zero, so
 *               a ray with Ray_ID == 0 skips every polygon whose mailbox entry was never written
 *               -- history-dependent in the reference.  Here Ray_ID == 0 means the fresh-mailbox
 *               case: the ray tests nothing and misses (Voxel_Grid, KDTree; the Octree mailbox is
 *               commented out, "Octree - alt.cs":221-222).  NULL = all non-zero.
 * Outputs (any may be NULL except poly_id):
 *   t           X_Event.t (0 on miss; includes t_start for rays entering a Voxel_Grid from outside)
 *   xyz         N x 3 X_Event.X_Point (0 on miss)
 *   poly_id     X_Event.Poly_id, HARE_MISS or HARE_RAY_FAULT; Hit == (poly_id >= 0)
 *   uv          N x 2 X_Event.u, v (always 0 for Voxel_Grid: Voxel_Grid.cs:487-488)
 *   o_moved     N x 3 ray origin after the call: Voxel_Grid moves a ray that starts outside the
 *               grid to its entry point (AABB_Main.cs:255-257; Ray is a class, the caller sees it)
 *   counters    HARE_CNT_N totals over the batch (cells or nodes visited, list entries scanned,
 *               polygon tests, hits); costs a little time, pass NULL when not needed.
 * Rays are independent: inside the call a batch of >= 65 536 rays is traversed grouped by origin / direction cell (ray_bin.cuh), the
 * warps claim their rays on demand, and the batch is cut into pipelined chunks on three streams (1 M rays growing to 16 M and shrinking
 * again; device staging of up to 3 x 2^24 rays x 52..136 bytes is allocated on first use and kept with the handle); events are written
 * by ray number, so the caller sees its own order and identical results whatever the batch size or the rays around a ray. */
int hare_shoot_batch(hare_part_t part, const double* o, const double* d,
                     const int32_t* origin1, const int32_t* origin2, const int32_t* ray_id, int64_t N,
                     double* t, double* xyz, int32_t* poly_id, double* uv, double* o_moved,
                     uint64_t* counters);

/* Same, with every array already resident on the partition's device (single-device handles);
 * launched on cuda_stream (a cudaStream_t; NULL = the handle's own stream; pass cudaStreamLegacy,
 * (void*)0x1, to name the legacy default stream) without synchronising.
 * counters_device, when non-NULL, points to HARE_CNT_N device uint64 that are ADDED to. */
int hare_shoot_batch_device(hare_part_t part, const double* o, const double* d,
                            const int32_t* origin1, const int32_t* origin2, const int32_t* ray_id, int64_t N,
                            double* t, double* xyz, int32_t* poly_id, double* uv, double* o_moved,
                            uint64_t* counters_device, void* cuda_stream);

/* Specular reflection chains kept on the device (BASELINE config 2; the caller-side loop around
 * Shoot -- Hare itself has no reflection).  Per bounce: Shoot with poly_origin1 = last hit polygon;
 * on a hit n = Polys[p].Normal, k = 2*((dx*nx)+(dy*ny)+(dz*nz)), d' = d - k*n, o' = X_Point; a chain
 * ends on a miss or after `order` Shoots.
 *   ev_poly_id, ev_t   N x order per-bounce events or NULL (HARE_NOT_SHOT / 0 after the chain ended)
 *   fin_o, fin_d       N x 3 ray state after the last Shoot, or NULL
 *   nshots             N number of Shoots performed per chain, or NULL
 *   total_shots        sum of nshots (always written) */
int hare_reflect_chain(hare_part_t part, const double* o, const double* d, int64_t N, int order,
                       int32_t* ev_poly_id, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots,
                       uint64_t* total_shots, uint64_t* counters);
int hare_reflect_chain_device(hare_part_t part, const double* o, const double* d, int64_t N, int order,
                              int32_t* ev_poly_id, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots,
                              uint64_t* total_shots_device, uint64_t* counters_device, void* cuda_stream);
/* The same with the rest of every bounce's X_Event as streams: ev_xyz N x order x 3 = X_Point
 * (Primitives.cs:435-481; the next segment starts there), ev_uv N x order x 2 = u, v (0 for Voxel_Grid, Voxel_Grid.cs:487-488).
 * Either may be NULL; rows of a miss and of Shoots that never happened are 0.  24 + 16 bytes per Shoot on top of the 12 of
 * ev_poly_id / ev_t: with host buffers these streams are what the call's time goes into (PCIe). */
int hare_reflect_chain_events(hare_part_t part, const double* o, const double* d, int64_t N, int order,
                              int32_t* ev_poly_id, double* ev_t, double* ev_xyz, double* ev_uv, double* fin_o, double* fin_d,
                              int32_t* nshots, uint64_t* total_shots, uint64_t* counters);
int hare_reflect_chain_events_device(hare_part_t part, const double* o, const double* d, int64_t N, int order,
                                     int32_t* ev_poly_id, double* ev_t, double* ev_xyz, double* ev_uv, double* fin_o, double* fin_d,
                                     int32_t* nshots, uint64_t* total_shots_device, uint64_t* counters_device, void* cuda_stream);

/* ---- host and device buffers for the batched calls --------------------------------- */
/* Page-locked host memory.  The reference's callers hold managed arrays; C# `fixed` / GCHandle only pins an array for the
 * garbage collector -- it does NOT page-lock it for CUDA, and cudaMemcpyAsync from pageable memory is staged and
 * synchronous, so the three-stream copy/compute pipeline of hare_shoot_batch / hare_reflect_chain degenerates to serial
 * copies.  Either allocate the ray / event arrays here (hare_host_alloc; wrap the pointer in a Span<T> / Memory<T>), or
 * page-lock an existing pinned array for the lifetime of the batches (hare_host_register, after GCHandle.Alloc(Pinned)).
 * hare_host_is_pinned reports how the library sees a pointer (1 page-locked, 0 pageable). */
int hare_host_alloc(size_t bytes, void** out);
int hare_host_free(void* p);
int hare_host_register(void* p, size_t bytes);
int hare_host_unregister(void* p);
int hare_host_is_pinned(const void* p);

/* Raw device buffers for the *_device entry points (cudaMalloc on `device`), and their export to the other ranks of a
 * one-process-per-GPU job: a rank that opened rank 0's result buffers passes those pointers (offset by its first ray) as the
 * outputs of hare_shoot_batch_device, and the traversal kernel's finish phase stores its X_Event rows straight into rank 0's
 * memory over NVLink -- the path's only cross-GPU step, fused into the kernel instead of a gather after it.
 * handle = the 64 bytes of a cudaIpcMemHandle_t. */
int hare_device_alloc(int device, size_t bytes, void** out);
int hare_device_free(int device, void* p);
int hare_device_memcpy(void* dst, const void* src, size_t bytes, int kind /* 1 H2D, 2 D2H, 3 D2D */, int device);
int hare_ipc_export(int device, void* dev_ptr, unsigned char handle[64]);
int hare_ipc_open(int device, const unsigned char handle[64], void** out);
int hare_ipc_close(int device, void* p);

/* Kernel launches issued by this library since load (bench.py's gpu_launches). */
uint64_t hare_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HARE_B200_H */
