// hare_abi.cu -- implementation of include/hare_b200.h: handles, device buffers, streams,
// kernel launches.  No CPU fallback: every compute entry point needs a usable CUDA device.
#include "../../include/hare_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host_build.hpp"
#include "kernels.cuh"
#include "oct_wave.cuh"
#include "kd_wave.cuh"
#include "ray_bin.cuh"
#include "pack.hpp"
#include "schedule.hpp"

using namespace hare;

// ---------------------------------------------------------------------------------------
// errors / globals
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{ 0 };
static std::mutex g_mu;
static std::vector<int> g_devices;   // empty = not initialised
static bool g_host_only = false;     // hare_init(NULL, -1): build-time tooling without a device

static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// host-side loops over polygons / nodes (record packing, bounding volumes): contiguous blocks on up to 16 threads
template <class F>
static void parallel_for(int64_t n, F&& body /* (begin, end, thread index) */) {
    const int64_t kMinBlock = 1 << 15;
    int nt = (int)std::min<int64_t>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u), (n + kMinBlock - 1) / kMinBlock);
    if (nt <= 1) { body((int64_t)0, n, 0); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < nt; ++k) th.emplace_back([&, k] { body(n * k / nt, n * (k + 1) / nt, k); });
    for (auto& x : th) x.join();
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(HARE_ERR_CUDA, b_);                                                        \
        }                                                                                          \
    } while (0)

static int ensure_init() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_host_only || !g_devices.empty()) return HARE_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(HARE_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (hare_b200 has no CPU fallback)");
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    g_devices.push_back(cur);
    return HARE_OK;
}

struct DevInfo { int sms = 0; };
static DevInfo dev_info(int dev) {
    DevInfo d; cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev); return d;
}

template <class T>
static cudaError_t dmalloc(T** p, size_t n) { *p = nullptr; return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }

// Build-path memory comes from the device's stream-ordered pool (cudaMallocAsync): a rebuild reuses the blocks the previous one
// returned instead of paying ~10 cudaMalloc / cudaFree round trips (which dominated and randomised the build's wall time).  The pool
// keeps up to 8 GB of freed blocks.  Persistent arrays allocated this way are released with plain cudaFree in free_partdev.
static void pool_keep(int dev) {
    static std::mutex mu; static std::vector<int> done;
    std::lock_guard<std::mutex> lk(mu);
    if (std::find(done.begin(), done.end(), dev) != done.end()) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) { uint64_t keep = 8ull << 30; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
    done.push_back(dev);
}
template <class T>
static cudaError_t pmalloc(T** p, size_t n, cudaStream_t st) { *p = nullptr; return cudaMallocAsync((void**)p, std::max<size_t>(n, 1) * sizeof(T), st); }
// temporaries of one build step: returned to the pool (stream-ordered) when the guard leaves scope, on every path
struct StreamTemps {
    cudaStream_t st; std::vector<void*> v;
    explicit StreamTemps(cudaStream_t s) : st(s) {}
    template <class T> cudaError_t get(T** p, size_t n) { cudaError_t e = pmalloc(p, n, st); if (e == cudaSuccess) v.push_back(*p); return e; }
    void keep(void* p) { v.erase(std::remove(v.begin(), v.end(), p), v.end()); }   // ownership moves to the caller
    void drop(void* p) { if (!p) return; keep(p); cudaFreeAsync(p, st); }
    ~StreamTemps() { for (void* p : v) cudaFreeAsync(p, st); }
};

// ---------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------
struct hare_topo_s {
    HostTopo host;
    std::atomic<int> parts{ 0 };       // partitions built on this Topology that are still alive (they read its device records)
    std::vector<int> devs;
    std::vector<PolyRec*> d_polys;   // one replica per device
    std::vector<float> pbox;         // host copy of the padded FP32 bounding boxes (P x 6: lo xyz, hi xyz), see cull_box()
};

// streams (and staging sets) per device of the host-buffer entry points: with three, the H2D copy of chunk k+1, the kernel of chunk k and
// the D2H copy of chunk k-1 overlap (two leave the copy engines idle while the second kernel runs)
static const int kStreams = 3;

struct PartDev {
    int dev = 0, sms = 0;
    cudaStream_t stream[kStreams] = {};
    const PolyRec* polys = nullptr;
    // voxel grid
    uint2* cells = nullptr; uint32_t* cell_poly = nullptr; uint32_t* occ = nullptr; uint32_t* occp = nullptr; uint32_t* cell_offset = nullptr;
    float4* list_box = nullptr;   // per list entry: padded FP32 bounding box + polygon id (VGrid::lbox; vg_wave.cuh's cull)
    // trees
    void* nodes = nullptr; uint32_t* lists = nullptr; KdNodeC* kd_hot = nullptr; KdWide* kd_wide = nullptr;
    double* ref_box = nullptr; float4* cbox = nullptr; float4* gbox = nullptr; float4* tbox = nullptr; float4* nbox = nullptr; float4* pbox = nullptr;   // octree chunk boxes; per tree-list entry boxes (+ polygon id); octree node content boxes
    // staging (per stream), sized for `cap` rays
    int64_t cap = 0; unsigned have = 0;   // capacity in rays; optional arrays present (ST_*)
    double *s_o[kStreams] = {}, *s_d[kStreams] = {}, *s_t[kStreams] = {}, *s_xyz[kStreams] = {}, *s_uv[kStreams] = {}, *s_om[kStreams] = {};
    int32_t *s_o1[kStreams] = {}, *s_o2[kStreams] = {}, *s_rid[kStreams] = {}, *s_pid[kStreams] = {};
    // chain staging
    int64_t ccap = 0, celems = 0;
    int32_t* c_evpid[kStreams] = {}; double* c_evt[kStreams] = {}; int32_t* c_ns[kStreams] = {};
    double* c_evxyz[kStreams] = {}; double* c_evuv[kStreams] = {}; int64_t cxyz = 0, cuv = 0;   // per-bounce X_Point / u, v rows (elements of capacity)
    unsigned long long* counters = nullptr;   // 4 counters + total_shots
    size_t bytes = 0;
};

struct hare_part_s {
    int kind = 0;
    hare_topo_s* topo = nullptr;
    std::vector<PartDev> dev;
    std::mutex mu;
    // voxel grid
    double obox[6] = {}, vd[3] = {}; int ct[3] = {}; int64_t npairs = 0;
    double build_kernel_ms = 0, build_wall_ms = 0;   // device time of the build kernels (CUDA events) / host wall time of the constructor
    // trees (host copies, for *_download and *_info)
    OctTree oct; KdTree kd; bool oct_regular = false;
};

static void free_partdev(PartDev& d) {
    cudaSetDevice(d.dev);
    for (int s = 0; s < kStreams; ++s) {
        cudaFree(d.s_o[s]); cudaFree(d.s_d[s]); cudaFree(d.s_t[s]); cudaFree(d.s_xyz[s]); cudaFree(d.s_uv[s]); cudaFree(d.s_om[s]);
        cudaFree(d.s_o1[s]); cudaFree(d.s_o2[s]); cudaFree(d.s_rid[s]); cudaFree(d.s_pid[s]);
        cudaFree(d.c_evpid[s]); cudaFree(d.c_evt[s]); cudaFree(d.c_ns[s]); cudaFree(d.c_evxyz[s]); cudaFree(d.c_evuv[s]);
        if (d.stream[s]) cudaStreamDestroy(d.stream[s]);
    }
    cudaFree(d.cells); cudaFree(d.cell_poly); cudaFree(d.list_box); cudaFree(d.occp); cudaFree(d.occ); cudaFree(d.cell_offset); cudaFree(d.nodes); cudaFree(d.kd_hot); cudaFree(d.kd_wide); cudaFree(d.lists); cudaFree(d.cbox); cudaFree(d.gbox); cudaFree(d.tbox); cudaFree(d.nbox); cudaFree(d.pbox); cudaFree(d.ref_box); cudaFree(d.counters);
}

static int init_partdev(PartDev& d, int dev, const PolyRec* polys) {
    d.dev = dev; d.polys = polys; d.sms = dev_info(dev).sms;
    CK(cudaSetDevice(dev));
    pool_keep(dev);
    for (int s = 0; s < kStreams; ++s) CK(cudaStreamCreateWithFlags(&d.stream[s], cudaStreamNonBlocking));
    CK(dmalloc(&d.counters, 8));
    CK(cudaMemset(d.counters, 0, 8 * sizeof(unsigned long long)));
    return HARE_OK;
}

// optional staging arrays (o, d and poly_id are always there)
enum : unsigned { ST_O1 = 1u, ST_O2 = 2u, ST_RID = 4u, ST_T = 8u, ST_XYZ = 16u, ST_UV = 32u, ST_OM = 64u };

// Staging sets for chunks of up to n rays carrying the arrays in `need`.  Grows only (capacity and set of arrays); called at the start
// of a host-buffer call, when the previous one has drained its streams.
static int ensure_staging(PartDev& d, int64_t n, unsigned need) {
    if (n <= d.cap && (need & ~d.have) == 0) return HARE_OK;
    n = std::max(n, d.cap); need |= d.have;
    CK(cudaSetDevice(d.dev));
    for (int s = 0; s < kStreams; ++s) {
        cudaFree(d.s_o[s]); cudaFree(d.s_d[s]); cudaFree(d.s_t[s]); cudaFree(d.s_xyz[s]); cudaFree(d.s_uv[s]); cudaFree(d.s_om[s]);
        cudaFree(d.s_o1[s]); cudaFree(d.s_o2[s]); cudaFree(d.s_rid[s]); cudaFree(d.s_pid[s]);
        d.s_o[s] = d.s_d[s] = d.s_t[s] = d.s_xyz[s] = d.s_uv[s] = d.s_om[s] = nullptr;
        d.s_o1[s] = d.s_o2[s] = d.s_rid[s] = d.s_pid[s] = nullptr;
    }
    d.cap = 0; d.have = 0;
    for (int s = 0; s < kStreams; ++s) {
        CK(dmalloc(&d.s_o[s], 3 * n)); CK(dmalloc(&d.s_d[s], 3 * n)); CK(dmalloc(&d.s_pid[s], n));
        if (need & ST_T) CK(dmalloc(&d.s_t[s], n));
        if (need & ST_XYZ) CK(dmalloc(&d.s_xyz[s], 3 * n));
        if (need & ST_UV) CK(dmalloc(&d.s_uv[s], 2 * n));
        if (need & ST_OM) CK(dmalloc(&d.s_om[s], 3 * n));
        if (need & ST_O1) CK(dmalloc(&d.s_o1[s], n));
        if (need & ST_O2) CK(dmalloc(&d.s_o2[s], n));
        if (need & ST_RID) CK(dmalloc(&d.s_rid[s], n));
    }
    d.cap = n; d.have = need;
    return HARE_OK;
}

static int ensure_chain_staging(PartDev& d, int64_t n, int order) {
    if (n * order <= d.celems && n <= d.ccap) return HARE_OK;
    CK(cudaSetDevice(d.dev));
    for (int s = 0; s < kStreams; ++s) {
        cudaFree(d.c_evpid[s]); cudaFree(d.c_evt[s]); cudaFree(d.c_ns[s]);
        CK(dmalloc(&d.c_evpid[s], (size_t)n * order)); CK(dmalloc(&d.c_evt[s], (size_t)n * order)); CK(dmalloc(&d.c_ns[s], n));
    }
    d.ccap = n; d.celems = n * order;
    return HARE_OK;
}

// staging for the optional per-bounce X_Point (3 doubles) / u, v (2 doubles) rows of n chains x order bounces
static int ensure_chain_rows(PartDev& d, int64_t n, int order, bool xyz, bool uv) {
    CK(cudaSetDevice(d.dev));
    const int64_t rows = n * order;
    if (xyz && 3 * rows > d.cxyz) {
        for (int s = 0; s < kStreams; ++s) { cudaFree(d.c_evxyz[s]); d.c_evxyz[s] = nullptr; }
        d.cxyz = 0;
        for (int s = 0; s < kStreams; ++s) CK(dmalloc(&d.c_evxyz[s], (size_t)(3 * rows)));
        d.cxyz = 3 * rows;
    }
    if (uv && 2 * rows > d.cuv) {
        for (int s = 0; s < kStreams; ++s) { cudaFree(d.c_evuv[s]); d.c_evuv[s] = nullptr; }
        d.cuv = 0;
        for (int s = 0; s < kStreams; ++s) CK(dmalloc(&d.c_evuv[s], (size_t)(2 * rows)));
        d.cuv = 2 * rows;
    }
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------
extern "C" const char* hare_version(void) { return "hare_b200 0.1 (sm_100a)"; }
extern "C" const char* hare_last_error(void) { return g_err.c_str(); }
namespace hare {
extern unsigned long long g_kd_build_launches;   // kd_build.cu
extern unsigned long long g_ingest_launches;     // ingest.cu
int topology_ingest_gpu(int dev, const double* raw, const int32_t* vcount, int64_t P, const double minpt[3], const double maxpt[3],
                        double* verts_out, double* normals_out, double minmax_out[6], int64_t* vertex_count_out, std::string& err);
int build_kdtree_gpu(const HostTopo& M, const PolyRec* d_polys, int dev, cudaStream_t st, int maxDepth, int maxPolys, KdTree& out, std::string& err);
}
extern "C" uint64_t hare_launch_count(void) { return g_launches.load() + hare::g_kd_build_launches + hare::g_ingest_launches; }

extern "C" int hare_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int hare_init(const int* device_ids, int n_devices) {
    if (n_devices == -1) {   // host-only handles: builders, info and download work; every compute call fails
        std::lock_guard<std::mutex> lk(g_mu);
        g_devices.clear(); g_host_only = true;
        return HARE_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(HARE_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (hare_b200 has no CPU fallback)");
    std::vector<int> devs;
    if (n_devices <= 0) {
        int cur = 0; CK(cudaGetDevice(&cur)); devs.push_back(cur);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            int id = device_ids ? device_ids[i] : i;
            if (id < 0 || id >= n) return fail(HARE_ERR_INVALID, "hare_init: device id out of range");
            devs.push_back(id);
        }
    }
    for (int id : devs) {
        cudaDeviceProp p;
        CK(cudaGetDeviceProperties(&p, id));
        if (p.major < 10) return fail(HARE_ERR_CUDA, "hare_init: device is not sm_100 or newer; this library ships sm_100a code only");
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_devices = devs; g_host_only = false;
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// Topology
// ---------------------------------------------------------------------------------------
extern "C" int hare_topology_ingest(const double* raw_verts, const int32_t* vcount, int64_t P, const double minpt[3], const double maxpt[3],
                                    double* verts_out, double* normals_out, double minmax_out[6], int64_t* vertex_count_out) {
    if (!raw_verts || !vcount || P < 0 || !minpt || !maxpt || !verts_out || !normals_out || !minmax_out)
        return fail(HARE_ERR_INVALID, "hare_topology_ingest: null argument");
    // The welding, normals and bounds run on the GPU when one is in use (ingest.cu, bit-identical to the host routine; SURVEY.md
    // 8(f) rank 2).  Ingest is build-time tooling that must also work on a machine without a device (hare_init(NULL, -1)), and
    // input the device path declines (vertices outside the declared bounds, ...) keeps the host routine's behaviour: it is the
    // same computation either way, not a fallback of the Shoot path.  HARE_INGEST_HOST=1 forces the host routine (A/B tests).
    int r = 1;
    {
        const char* e = getenv("HARE_INGEST_HOST");
        int dev = -1;
        if (!(e && *e == '1') && P >= 4096 && ensure_init() == HARE_OK) {
            std::lock_guard<std::mutex> lk(g_mu);
            if (!g_host_only && !g_devices.empty()) dev = g_devices[0];
        }
        if (dev >= 0) {
            std::string msg;
            r = topology_ingest_gpu(dev, raw_verts, vcount, P, minpt, maxpt, verts_out, normals_out, minmax_out, vertex_count_out, msg);
            if (r == -2) return fail(HARE_ERR_CUDA, msg);
        }
    }
    if (r == 1) r = topology_ingest(raw_verts, vcount, P, minpt, maxpt, verts_out, normals_out, minmax_out, vertex_count_out);
    if (r == -3) return fail(HARE_ERR_UNSUPPORTED, "Hare Does not yet support polygons of more than 4 sides.");
    return r;
}

extern "C" int hare_topology_create(const double* verts, const double* normals, const int32_t* vcount, int64_t P,
                                    const double minmax[6], hare_topo_t* out) {
    if (!verts || !normals || !vcount || P <= 0 || !minmax || !out) return fail(HARE_ERR_INVALID, "hare_topology_create: bad argument");
    if (P > 0x7fffffffLL) return fail(HARE_ERR_INVALID, "hare_topology_create: more than 2^31-1 polygons");
    int rc = ensure_init();
    if (rc) return rc;
    for (int64_t i = 0; i < P; ++i)
        if (vcount[i] != 3 && vcount[i] != 4) return fail(HARE_ERR_UNSUPPORTED, "Hare Does not yet support polygons of more than 4 sides.");
    hare_topo_s* t = new hare_topo_s();
    t->host.P = P;
    t->host.verts.assign(verts, verts + 12 * P);
    t->host.normals.assign(normals, normals + 3 * P);
    t->host.vcount.assign(vcount, vcount + P);
    std::memcpy(t->host.minmax, minmax, 6 * sizeof(double));
    for (int a = 0; a < 3; ++a) { t->host.vmin[a] = INFINITY; t->host.vmax[a] = -INFINITY; }
    std::vector<PolyRec> recs((size_t)P);
    t->pbox.resize((size_t)P * 6);
    double part_min[16][3], part_max[16][3];
    for (int k = 0; k < 16; ++k) for (int a = 0; a < 3; ++a) { part_min[k][a] = INFINITY; part_max[k][a] = -INFINITY; }
    parallel_for(P, [&](int64_t i0, int64_t i1, int tid) {
        double* vmn = part_min[tid]; double* vmx = part_max[tid];
        for (int64_t i = i0; i < i1; ++i) {
            for (int k = 0; k < 12; ++k) recs[i].v[k] = verts[12 * i + k];
            if (vcount[i] == 3) for (int a = 0; a < 3; ++a) recs[i].v[9 + a] = verts[12 * i + 6 + a];
            for (int a = 0; a < 3; ++a) recs[i].v[12 + a] = normals[3 * i + a];
            recs[i].v[15] = (double)vcount[i];
            double lo[3], hi[3];
            for (int a = 0; a < 3; ++a) { lo[a] = hi[a] = verts[12 * i + a]; }
            for (int k = 1; k < vcount[i]; ++k)
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], verts[12 * i + 3 * k + a]); hi[a] = std::max(hi[a], verts[12 * i + 3 * k + a]); }
            for (int a = 0; a < 3; ++a) { if (lo[a] < vmn[a]) vmn[a] = lo[a]; if (hi[a] > vmx[a]) vmx[a] = hi[a]; }
            // padded FP32 bounding box (the conservative reject cull_box): exact box -/+ hare_box_pad, rounded outwards
            poly_pad_box(verts + 12 * i, vcount[i], &t->pbox[6 * i]);
        }
    });
    for (int k = 0; k < 16; ++k)
        for (int a = 0; a < 3; ++a) { t->host.vmin[a] = std::min(t->host.vmin[a], part_min[k][a]); t->host.vmax[a] = std::max(t->host.vmax[a], part_max[k][a]); }
    { std::lock_guard<std::mutex> lk(g_mu); t->devs = g_devices; }
    for (int dev : t->devs) {
        PolyRec* d = nullptr;
        cudaError_t e = cudaSetDevice(dev);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d, (size_t)P * sizeof(PolyRec));
        if (e == cudaSuccess) e = cudaMemcpy(d, recs.data(), (size_t)P * sizeof(PolyRec), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            for (size_t k = 0; k < t->d_polys.size(); ++k) { cudaSetDevice(t->devs[k]); cudaFree(t->d_polys[k]); }
            cudaFree(d); delete t;
            return fail(HARE_ERR_CUDA, std::string("hare_topology_create: ") + cudaGetErrorString(e));
        }
        t->d_polys.push_back(d);
    }
    *out = t;
    return HARE_OK;
}

extern "C" int64_t hare_topology_polygon_count(hare_topo_t t) { return t ? t->host.P : -1; }

extern "C" int hare_topology_destroy(hare_topo_t t) {
    if (!t) return HARE_OK;
    if (t->parts.load() > 0) return fail(HARE_ERR_INVALID, "hare_topology_destroy: partitions built on this Topology are still alive (destroy them first)");
    for (size_t k = 0; k < t->d_polys.size(); ++k) { cudaSetDevice(t->devs[k]); cudaFree(t->d_polys[k]); }
    delete t;
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// Voxel_Grid
// ---------------------------------------------------------------------------------------
static void vg_bounds(const HostTopo& M, double obox[6]) {
    // Voxel_Grid.cs:50-75 for a single topology: MinPT/MaxPT = Model.Min/Max -/+ Epsilon, OBox = that -/+ 0.1
    const double Epsilon = 0.001;
    for (int a = 0; a < 3; ++a) {
        double MaxPT = -INFINITY, MinPT = INFINITY;
        if ((M.minmax[3 + a] + 0.01) > MaxPT) MaxPT = (M.minmax[3 + a] + Epsilon);
        if ((M.minmax[a] - 0.01) < MinPT) MinPT = (M.minmax[a] - Epsilon);
        obox[a] = MinPT - .1;
        obox[3 + a] = MaxPT + .1;
    }
}

static VGrid make_vgrid(const hare_part_s* p, const PartDev& d) {
    VGrid g;
    g.ominx = p->obox[0]; g.ominy = p->obox[1]; g.ominz = p->obox[2]; g.omaxx = p->obox[3]; g.omaxy = p->obox[4]; g.omaxz = p->obox[5];
    g.vdx = p->vd[0]; g.vdy = p->vd[1]; g.vdz = p->vd[2];
    g.nx = p->ct[0]; g.ny = p->ct[1]; g.nz = p->ct[2];
    g.cells = d.cells; g.cell_poly = d.cell_poly; g.occ = d.occ; g.occp = d.occp;
    g.lbox = d.list_box;
    return g;
}

static void drop_part(hare_part_s* p) { if (p->topo) --p->topo->parts; delete p; }

static int new_part(hare_topo_t topo, int kind, hare_part_s** out) {
    hare_part_s* p = new hare_part_s();
    p->kind = kind; p->topo = topo;
    p->dev.resize(topo->devs.size());
    for (size_t k = 0; k < topo->devs.size(); ++k) {
        int rc = init_partdev(p->dev[k], topo->devs[k], topo->d_polys[k]);
        if (rc) { for (auto& d : p->dev) free_partdev(d); delete p; return rc; }
    }
    ++topo->parts;
    *out = p;
    return HARE_OK;
}

static int vg_set_dims(hare_part_s* p, const double obox[6], const int32_t ct[3]) {
    std::memcpy(p->obox, obox, 6 * sizeof(double));
    for (int a = 0; a < 3; ++a) {
        p->ct[a] = ct[a];
        p->vd[a] = (obox[3 + a] - obox[a]) / ct[a];   // VoxelDims = BoxDims / VoxelCt  Voxel_Grid.cs:85-86
    }
    return HARE_OK;
}

static int vg_make_list_box(PartDev& d, uint32_t total, cudaStream_t st, const int ct[3]) {
    {   // the wavefront kernel's border-padded occupancy bitmap
        const int64_t padded = ((int64_t)ct[0] + 2) * ((int64_t)ct[1] + 2) * ((int64_t)ct[2] + 2);
        if (padded < (1LL << 32)) {
            CK(pmalloc(&d.occp, (size_t)((padded + 31) / 32 + 1), st));
            vg_pad_occupancy<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(d.occ, ct[0], ct[1], ct[2], d.occp);
            ++g_launches;
            CK(cudaGetLastError());
        }
    }
    if (total == 0) return HARE_OK;
    CK(pmalloc(&d.list_box, 2 * (size_t)total, st));
    vg_gather_list_box<<<(unsigned)(((int64_t)total + 255) / 256), 256, 0, st>>>(d.cell_poly, d.polys, total, d.list_box);
    ++g_launches;
    CK(cudaGetLastError());
    return HARE_OK;
}

static int scan_u32(const uint32_t* in, int64_t n, uint32_t* out /* n + 1 */, uint32_t* tile_tmp, cudaStream_t st) {
    const int64_t tiles = (n + HARE_SCAN_TILE - 1) / HARE_SCAN_TILE;
    scan_tile_sums<<<(unsigned)tiles, 1024, 0, st>>>(in, n, tile_tmp);
    scan_tile_offsets<<<1, 1024, 0, st>>>(tile_tmp, tiles, out + n);
    scan_tiles<<<(unsigned)tiles, 1024, 0, st>>>(in, n, tile_tmp, out);
    g_launches += 3;
    CK(cudaGetLastError());
    return HARE_OK;
}

struct BuildTimer {   // device time between the first and the last kernel of a build (CUDA events on its stream)
    cudaEvent_t e0 = nullptr, e1 = nullptr; cudaStream_t st;
    explicit BuildTimer(cudaStream_t s) : st(s) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    double stop() { float ms = 0; cudaEventRecord(e1, st); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); return ms; }
    ~BuildTimer() { cudaEventDestroy(e0); cudaEventDestroy(e1); }
};

// headers, occupancy bits, ascending lists, padded bitmap and per-entry boxes of a grid whose cell_offset / cell_poly / counts exist
static int vg_finalize(hare_part_s* p, PartDev& d, const uint32_t* count, uint32_t total, cudaStream_t st) {
    const int64_t ncells = (int64_t)p->ct[0] * p->ct[1] * p->ct[2];
    CK(pmalloc(&d.cells, ncells, st)); CK(pmalloc(&d.occ, (ncells + 31) / 32 + 1, st));
    vg_finish_cells<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(d.cell_offset, count, ncells, d.cell_poly, d.cells, d.occ);
    ++g_launches;
    CK(cudaGetLastError());
    int r = vg_make_list_box(d, total, st, p->ct);
    if (r) return r;
    d.bytes = (size_t)ncells * 12 + (size_t)total * 36 + (size_t)ncells / 8;
    return HARE_OK;
}

extern "C" int hare_voxelgrid_build(hare_topo_t topo, int domain, hare_part_t* out) {
    if (!topo || !out || domain < 1 || domain > 1290) return fail(HARE_ERR_INVALID, "hare_voxelgrid_build: bad argument (1 <= Domain <= 1290)");
    if (topo->devs.empty()) return fail(HARE_ERR_CUDA, "hare_voxelgrid_build: host-only handle, no CUDA device (hare_b200 has no CPU fallback)");
    const auto w0 = std::chrono::steady_clock::now();
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_VOXEL_GRID, &p);
    if (rc) return rc;
    double obox[6]; vg_bounds(topo->host, obox);
    const int32_t ct[3] = { domain, domain, domain };
    vg_set_dims(p, obox, ct);
    const int64_t ncells = (int64_t)domain * domain * domain;
    const int64_t P = topo->host.P;
    for (PartDev& d : p->dev) {
        auto body = [&]() -> int {
            CK(cudaSetDevice(d.dev));
            cudaStream_t st = d.stream[0];
            StreamTemps tmp(st);
            BuildTimer timer(st);
            uint32_t *count = nullptr, *cursor = nullptr, *tiles = nullptr;
            CK(tmp.get(&count, ncells)); CK(tmp.get(&cursor, ncells));
            CK(tmp.get(&tiles, (ncells + HARE_SCAN_TILE - 1) / HARE_SCAN_TILE + 1));
            CK(pmalloc(&d.cell_offset, ncells + 1, st));
            CK(cudaMemsetAsync(count, 0, ncells * 4, st)); CK(cudaMemsetAsync(cursor, 0, ncells * 4, st));
            VGBuild g = { obox[0], obox[1], obox[2], p->vd[0], p->vd[1], p->vd[2], domain, domain, domain };
            const int blocks = d.sms * 8;
            vg_bin_kernel<0><<<blocks, 256, 0, st>>>(g, d.polys, P, count, nullptr, nullptr, nullptr);
            ++g_launches;
            CK(cudaGetLastError());
            int r = scan_u32(count, ncells, d.cell_offset, tiles, st);
            if (r) return r;
            uint32_t total = 0;
            CK(cudaMemcpyAsync(&total, d.cell_offset + ncells, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            p->npairs = total;
            CK(pmalloc(&d.cell_poly, (size_t)total, st));
            vg_bin_kernel<1><<<blocks, 256, 0, st>>>(g, d.polys, P, nullptr, d.cell_offset, cursor, d.cell_poly);
            ++g_launches;
            CK(cudaGetLastError());
            r = vg_finalize(p, d, count, total, st);
            if (r) return r;
            p->build_kernel_ms = timer.stop();
            return HARE_OK;
        };
        rc = body();
        if (rc) { for (auto& dd : p->dev) { cudaSetDevice(dd.dev); cudaStreamSynchronize(dd.stream[0]); free_partdev(dd); } drop_part(p); return rc; }
    }
    p->build_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    *out = p;
    return HARE_OK;
}

// new Voxel_Grid(Model, MaxDomain, Avg_polys)  (Voxel_Grid.cs:128-254), built the reference's way: level k + 1 tests each child voxel
// only against its parent's list (:207-215), starting from the single voxel that lists every polygon, and stops after level k > 1 once
// the mean list length of the non-empty voxels drops below Avg_polys (:252).  O(pairs) per level instead of a full flat rebuild, and
// exactly the reference's semantics (no reliance on PolyBoxOverlap being monotone under box inclusion in floating point).
extern "C" int hare_voxelgrid_build_adaptive(hare_topo_t topo, int max_domain_log2, int avg_polys, hare_part_t* out) {
    if (!topo || !out || max_domain_log2 < 1 || max_domain_log2 > 10) return fail(HARE_ERR_INVALID, "hare_voxelgrid_build_adaptive: 1 <= MaxDomain <= 10");
    if (topo->devs.empty()) return fail(HARE_ERR_CUDA, "hare_voxelgrid_build_adaptive: host-only handle, no CUDA device (hare_b200 has no CPU fallback)");
    const auto w0 = std::chrono::steady_clock::now();
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_VOXEL_GRID, &p);
    if (rc) return rc;
    double obox[6]; vg_bounds(topo->host, obox);
    const int64_t P = topo->host.P;
    int final_level = -1;
    for (PartDev& d : p->dev) {
        auto body = [&]() -> int {
            CK(cudaSetDevice(d.dev));
            cudaStream_t st = d.stream[0];
            StreamTemps tmp(st);
            BuildTimer timer(st);
            // level "-1": one voxel listing every polygon (:158-166)
            uint32_t *poff = nullptr, *ppoly = nullptr, *count = nullptr;
            unsigned long long* d_nonempty = nullptr;
            CK(tmp.get(&poff, 2)); CK(tmp.get(&ppoly, (size_t)P)); CK(tmp.get(&d_nonempty, 1));
            const uint32_t off0[2] = { 0u, (uint32_t)P };
            CK(cudaMemcpyAsync(poff, off0, 8, cudaMemcpyHostToDevice, st));
            fill_iota<<<d.sms * 4, 256, 0, st>>>(ppoly, P);
            ++g_launches;
            int64_t npairs = P, pcells = 1;
            for (int k = 0; k < max_domain_log2; ++k) {        // :169-253
                if (final_level >= 0 && k > final_level) break;  // further devices replay the level count the first one settled on
                const int n = 1 << (k + 1);
                const int32_t ct[3] = { n, n, n };
                vg_set_dims(p, obox, ct);
                const int64_t ncells = (int64_t)n * n * n;
                VGBuild g = { obox[0], obox[1], obox[2], p->vd[0], p->vd[1], p->vd[2], n, n, n };
                uint32_t *cnt = nullptr, *cursor = nullptr, *tiles = nullptr, *coff = nullptr, *cpoly = nullptr; uint8_t* mask = nullptr;
                CK(tmp.get(&cnt, ncells)); CK(tmp.get(&cursor, ncells)); CK(tmp.get(&tiles, (ncells + HARE_SCAN_TILE - 1) / HARE_SCAN_TILE + 1));
                CK(tmp.get(&coff, ncells + 1)); CK(tmp.get(&mask, (size_t)npairs));
                CK(cudaMemsetAsync(cnt, 0, ncells * 4, st)); CK(cudaMemsetAsync(cursor, 0, ncells * 4, st)); CK(cudaMemsetAsync(d_nonempty, 0, 8, st));
                const unsigned blocks = (unsigned)((npairs * 8 + 255) / 256);
                if (npairs) vg_refine_kernel<0><<<blocks, 256, 0, st>>>(g, d.polys, poff, ppoly, npairs, pcells, mask, cnt, nullptr, nullptr, nullptr);
                int r = scan_u32(cnt, ncells, coff, tiles, st);
                if (r) return r;
                vg_count_nonempty<<<d.sms * 4, 256, 0, st>>>(cnt, ncells, d_nonempty);
                g_launches += 2;
                uint32_t total = 0; unsigned long long nonempty = 0;
                CK(cudaMemcpyAsync(&total, coff + ncells, 4, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(&nonempty, d_nonempty, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                CK(tmp.get(&cpoly, (size_t)total));
                if (npairs) vg_refine_kernel<1><<<blocks, 256, 0, st>>>(g, d.polys, poff, ppoly, npairs, pcells, mask, nullptr, coff, cursor, cpoly);
                ++g_launches;
                CK(cudaGetLastError());
                tmp.drop(poff); tmp.drop(ppoly); tmp.drop(mask); tmp.drop(cursor); tmp.drop(tiles); tmp.drop(count);
                poff = coff; ppoly = cpoly; count = cnt; npairs = total; pcells = ncells;
                const double avg = nonempty ? (double)total / (double)nonempty : 0.0 / 0.0;   // C#: 0/0 -> NaN, and NaN < Avg_polys is false
                const bool stop = (final_level >= 0) ? (k == final_level) : (k > 1 && avg < (double)avg_polys);   // :252
                if (stop || k == max_domain_log2 - 1) { final_level = k; break; }
            }
            // the last level becomes the partition: sort the lists ascending, headers, bitmaps, per-entry boxes
            tmp.keep(poff); tmp.keep(ppoly);
            d.cell_offset = poff; d.cell_poly = ppoly;
            p->npairs = npairs;
            int r = vg_finalize(p, d, count, (uint32_t)npairs, st);
            if (r) return r;
            p->build_kernel_ms = timer.stop();
            return HARE_OK;
        };
        rc = body();
        if (rc) { for (auto& dd : p->dev) { cudaSetDevice(dd.dev); cudaStreamSynchronize(dd.stream[0]); free_partdev(dd); } drop_part(p); return rc; }
    }
    p->build_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    *out = p;
    return HARE_OK;
}

extern "C" int hare_part_build_ms(hare_part_t p, double* kernel_ms, double* wall_ms) {
    if (!p) return fail(HARE_ERR_INVALID, "hare_part_build_ms: null handle");
    if (kernel_ms) *kernel_ms = p->build_kernel_ms;
    if (wall_ms) *wall_ms = p->build_wall_ms;
    return HARE_OK;
}

extern "C" int hare_voxelgrid_upload(hare_topo_t topo, const double obox[6], const int32_t ct[3],
                                     const uint32_t* cell_offset, const uint32_t* cell_poly, hare_part_t* out) {
    if (!topo || !obox || !ct || !cell_offset || !out || ct[0] < 1 || ct[1] < 1 || ct[2] < 1)
        return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: bad argument");
    const int64_t ncells = (int64_t)ct[0] * ct[1] * ct[2];
    if (ncells > 0x7fffffffLL) return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: more than 2^31-1 cells");
    const uint32_t total = cell_offset[ncells];
    if (total && !cell_poly) return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: cell_poly is null");
    if (cell_offset[0] != 0) return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: cell_offset[0] must be 0");
    for (int64_t c = 0; c < ncells; ++c)
        if (cell_offset[c] > cell_offset[c + 1]) return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: cell_offset must be non-decreasing");
    for (uint32_t k = 0; k < total; ++k)
        if ((int64_t)cell_poly[k] >= topo->host.P) return fail(HARE_ERR_INVALID, "hare_voxelgrid_upload: polygon index out of range");
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_VOXEL_GRID, &p);
    if (rc) return rc;
    vg_set_dims(p, obox, ct);
    p->npairs = total;
    for (PartDev& d : p->dev) {
        auto body = [&]() -> int {
            CK(cudaSetDevice(d.dev));
            cudaStream_t st = d.stream[0];
            CK(dmalloc(&d.cell_offset, ncells + 1)); CK(dmalloc(&d.cells, ncells)); CK(dmalloc(&d.occ, (ncells + 31) / 32 + 1));
            CK(dmalloc(&d.cell_poly, (size_t)total));
            CK(cudaMemcpyAsync(d.cell_offset, cell_offset, (ncells + 1) * 4, cudaMemcpyHostToDevice, st));
            if (total) CK(cudaMemcpyAsync(d.cell_poly, cell_poly, (size_t)total * 4, cudaMemcpyHostToDevice, st));
            vg_pack_cells<<<(unsigned)((ncells + 255) / 256), 256, 0, st>>>(d.cell_offset, ncells, d.cells, d.occ);
            ++g_launches;
            CK(cudaGetLastError());
            int r = vg_make_list_box(d, total, st, ct);
            if (r) return r;
            CK(cudaStreamSynchronize(st));
            d.bytes = (size_t)ncells * 12 + (size_t)total * 36 + (size_t)ncells / 8;
            return HARE_OK;
        };
        rc = body();
        if (rc) { for (auto& dd : p->dev) free_partdev(dd); drop_part(p); return rc; }
    }
    *out = p;
    return HARE_OK;
}

extern "C" int hare_voxelgrid_info(hare_part_t p, double obox[6], double voxeldims[3], int32_t ct[3], int64_t* npairs) {
    if (!p || p->kind != HARE_VOXEL_GRID) return fail(HARE_ERR_INVALID, "hare_voxelgrid_info: not a Voxel_Grid");
    if (obox) std::memcpy(obox, p->obox, 6 * sizeof(double));
    if (voxeldims) std::memcpy(voxeldims, p->vd, 3 * sizeof(double));
    if (ct) std::memcpy(ct, p->ct, 3 * sizeof(int32_t));
    if (npairs) *npairs = p->npairs;
    return HARE_OK;
}

extern "C" int hare_voxelgrid_download(hare_part_t p, uint32_t* cell_offset, uint32_t* cell_poly) {
    if (!p || p->kind != HARE_VOXEL_GRID) return fail(HARE_ERR_INVALID, "hare_voxelgrid_download: not a Voxel_Grid");
    if (p->dev.empty()) return fail(HARE_ERR_CUDA, "hare_voxelgrid_download: host-only handle");
    PartDev& d = p->dev[0];
    const int64_t ncells = (int64_t)p->ct[0] * p->ct[1] * p->ct[2];
    CK(cudaSetDevice(d.dev));
    if (cell_offset) CK(cudaMemcpy(cell_offset, d.cell_offset, (ncells + 1) * 4, cudaMemcpyDeviceToHost));
    if (cell_poly && p->npairs) CK(cudaMemcpy(cell_poly, d.cell_poly, (size_t)p->npairs * 4, cudaMemcpyDeviceToHost));
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// Octree
// ---------------------------------------------------------------------------------------
static int oct_depth(const OctTree& t) { return oct_depth_of(t); }   // pack.hpp: index order, children follow their parents

// per tree-list entry: the polygon's padded FP32 box with its id riding in lo.w -- gathered on the device from the records
// (same kernel as the Voxel_Grid cell lists)
static int tree_entry_boxes(PartDev& d, size_t n_list, int64_t P, bool per_polygon) {
    if (per_polygon) {   // one box per polygon, read by id
        CK(dmalloc(&d.pbox, 2 * (size_t)P));
        poly_box_table<<<(unsigned)((P + 255) / 256), 256, 0, d.stream[0]>>>(d.polys, (uint32_t)P, d.pbox);
    } else {             // one (box, id) record per list entry, read in list order
        CK(dmalloc(&d.tbox, 2 * n_list));
        if (n_list == 0) return HARE_OK;
        vg_gather_list_box<<<(unsigned)((n_list + 255) / 256), 256, 0, d.stream[0]>>>(d.lists, d.polys, (uint32_t)n_list, d.tbox);
    }
    ++g_launches;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(d.stream[0]));
    return HARE_OK;
}

static int oct_to_device(hare_part_s* p) {
    const OctTree& t = p->oct;
    const size_t N = t.first_child.size();
    if (oct_depth(t) >= HARE_OCT_MAXLVL) return fail(HARE_ERR_UNSUPPORTED, "octree deeper than HARE_OCT_MAXLVL levels");
    // node records (content masks, chunk indices), chunk / group boxes and node content boxes: pack.hpp (shared with tests/emu)
    PackedOct pk;
    pack_octree(t, p->topo->pbox.data(), pk);
    p->oct_regular = pk.regular;
    for (PartDev& d : p->dev) {
        CK(cudaSetDevice(d.dev));
        OctNode* dn = nullptr;
        CK(dmalloc(&dn, N)); d.nodes = dn;
        CK(dmalloc(&d.lists, t.polys.size() + 8));
        CK(dmalloc(&d.cbox, pk.cbox.size() / 4 + 16)); CK(dmalloc(&d.nbox, pk.nbox.size() / 4));
        CK(cudaMemcpy(d.nbox, pk.nbox.data(), pk.nbox.size() * 4, cudaMemcpyHostToDevice));
        if (!pk.cbox.empty()) CK(cudaMemcpy(d.cbox, pk.cbox.data(), pk.cbox.size() * 4, cudaMemcpyHostToDevice));
        CK(dmalloc(&d.gbox, pk.gbox.size() / 4 + 2));
        if (!pk.gbox.empty()) CK(cudaMemcpy(d.gbox, pk.gbox.data(), pk.gbox.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dn, pk.nodes.data(), N * sizeof(OctNode), cudaMemcpyHostToDevice));
        if (!t.polys.empty()) CK(cudaMemcpy(d.lists, t.polys.data(), t.polys.size() * 4, cudaMemcpyHostToDevice));
        { int r = tree_entry_boxes(d, t.polys.size(), p->topo->host.P, true); if (r) return r; }
        d.bytes = N * sizeof(OctNode) + t.polys.size() * 4 + (size_t)p->topo->host.P * 32 + (pk.cbox.size() + pk.gbox.size() + pk.nbox.size()) * 4;
    }
    return HARE_OK;
}

// Octree build on the GPU (kernels.cuh "GPU Octree build").  Produces the same OctTree as host_build.cpp's
// build_octree(): same breadth-first node numbering, same child boxes (computed here on the host with the same
// expressions), same list order.
static int build_octree_gpu(hare_topo_s* topo, const PolyRec* d_polys, int dev, cudaStream_t st, int maxDepth, int maxPolys, OctTree& out) {
    const HostTopo& T = topo->host;
    out = OctTree();
    CK(cudaSetDevice(dev));
    const double ext = net_max(T.vmax[0] - T.vmin[0], net_max(T.vmax[1] - T.vmin[1], T.vmax[2] - T.vmin[2]));
    auto add_node = [&](const double* mn, const double* mx) {
        for (int a2 = 0; a2 < 3; ++a2) out.box.push_back(mn[a2]);
        for (int a2 = 0; a2 < 3; ++a2) out.box.push_back(mx[a2]);
        out.first_child.push_back(-1); out.list_off.push_back(0); out.list_cnt.push_back(0);
        return (int)out.first_child.size() - 1;
    };
    {
        double mn[3], mx[3];
        for (int a2 = 0; a2 < 3; ++a2) { const double c = T.vmax[a2] + T.vmin[a2] / 2; mn[a2] = c - ext - 1e-1; mx[a2] = c + ext + 1e-1; }   // :79-82
        add_node(mn, mx);
    }
    struct LNode { int node; uint32_t start, cnt; };
    std::vector<LNode> level(1);
    level[0] = { 0, 0u, (uint32_t)T.P };
    uint32_t* cur = nullptr;   // this level's entry array (polygon ids, node-major)
    CK(dmalloc(&cur, (size_t)T.P));
    {
        std::vector<uint32_t> iota((size_t)T.P);
        for (size_t i = 0; i < iota.size(); ++i) iota[i] = (uint32_t)i;
        CK(cudaMemcpyAsync(cur, iota.data(), iota.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
    }
    int64_t E = T.P;
    unsigned long long* d_lost = nullptr;
    CK(dmalloc(&d_lost, 1)); CK(cudaMemsetAsync(d_lost, 0, 8, st));
    std::vector<uint32_t*> leaf_src;                       // device entry arrays holding leaf lists, kept until the final gather
    struct LeafCopy { uint32_t src, dst, cnt; int arr; };
    std::vector<LeafCopy> copies;
    uint32_t list_total = 0;
    int rcode = HARE_OK;
    for (int depth = 0; !level.empty(); ++depth) {
        out.depth = depth;
        std::vector<LNode> split;
        leaf_src.push_back(cur);
        const int arr = (int)leaf_src.size() - 1;
        for (const LNode& n : level) {
            if (depth >= maxDepth || (int64_t)n.cnt <= (int64_t)maxPolys) {                 // :93
                out.list_off[n.node] = list_total; out.list_cnt[n.node] = n.cnt;
                if (n.cnt) copies.push_back({ n.start, list_total, n.cnt, arr });
                list_total += n.cnt;
            } else split.push_back(n);
        }
        if (split.empty()) break;
        const int S = (int)split.size();
        std::vector<double> cbox((size_t)S * 48);
        std::vector<uint32_t> seg_start(S), seg_cnt(S);
        for (int s = 0; s < S; ++s) {
            const int nd = split[s].node;
            double mn[3], mx[3], mid[3];
            for (int a2 = 0; a2 < 3; ++a2) { mn[a2] = out.box[6 * nd + a2]; mx[a2] = out.box[6 * nd + 3 + a2]; mid[a2] = (mx[a2] + mn[a2]) / 2; }
            out.first_child[nd] = (int)out.first_child.size();
            for (int i = 0; i < 8; ++i) {
                double cmn[3], cmx[3];
                for (int a2 = 0; a2 < 3; ++a2) {
                    const bool upper = (i & (4 >> a2)) != 0;
                    cmn[a2] = (upper ? mid[a2] : mn[a2]) - 0.1;
                    cmx[a2] = (upper ? mx[a2] : mid[a2]) + 0.1;
                }
                add_node(cmn, cmx);
                for (int a2 = 0; a2 < 3; ++a2) { cbox[((size_t)s * 8 + i) * 6 + a2] = cmn[a2]; cbox[((size_t)s * 8 + i) * 6 + 3 + a2] = cmx[a2]; }
            }
            seg_start[s] = split[s].start; seg_cnt[s] = split[s].cnt;
        }
        auto body = [&]() -> int {
            double* d_cbox = nullptr; uint32_t *d_ss = nullptr, *d_sc = nullptr, *d_slot = nullptr, *d_flags = nullptr, *d_scan = nullptr, *d_tiles = nullptr, *d_counts = nullptr, *d_base = nullptr;
            CK(dmalloc(&d_cbox, cbox.size())); CK(dmalloc(&d_ss, (size_t)S)); CK(dmalloc(&d_sc, (size_t)S));
            CK(dmalloc(&d_slot, (size_t)E)); CK(dmalloc(&d_flags, (size_t)E * 8)); CK(dmalloc(&d_scan, (size_t)E * 8 + 1));
            CK(dmalloc(&d_tiles, ((size_t)E * 8 + HARE_SCAN_TILE - 1) / HARE_SCAN_TILE + 1)); CK(dmalloc(&d_counts, (size_t)S * 8)); CK(dmalloc(&d_base, (size_t)S * 8));
            CK(cudaMemcpyAsync(d_cbox, cbox.data(), cbox.size() * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_ss, seg_start.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(d_sc, seg_cnt.data(), (size_t)S * 4, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(d_slot, 0xff, (size_t)E * 4, st));
            oct_fill_slot<<<std::min(S, 65535), 256, 0, st>>>(d_ss, d_sc, S, d_slot);
            const unsigned eb = (unsigned)((E * 8 + 255) / 256);
            oct_mask_kernel<<<eb, 256, 0, st>>>(d_polys, cur, d_slot, d_cbox, E, d_flags, d_lost);
            g_launches += 2;
            CK(cudaGetLastError());
            int r = scan_u32(d_flags, E * 8, d_scan, d_tiles, st);
            if (r) return r;
            oct_child_counts<<<(S * 8 + 255) / 256, 256, 0, st>>>(d_scan, d_ss, d_sc, S, E, d_counts);
            ++g_launches;
            std::vector<uint32_t> counts((size_t)S * 8), base((size_t)S * 8);
            CK(cudaMemcpyAsync(counts.data(), d_counts, counts.size() * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            uint64_t tot = 0;
            for (size_t i = 0; i < counts.size(); ++i) { base[i] = (uint32_t)tot; tot += counts[i]; }
            if (tot > 0xfffffff0ull) return fail(HARE_ERR_UNSUPPORTED, "octree level with more than 2^32 list entries");
            uint32_t* nxt = nullptr;
            CK(dmalloc(&nxt, (size_t)tot));
            CK(cudaMemcpyAsync(d_base, base.data(), base.size() * 4, cudaMemcpyHostToDevice, st));
            oct_scatter<<<eb, 256, 0, st>>>(cur, d_slot, d_flags, d_scan, d_ss, d_base, E, nxt);
            ++g_launches;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));
            std::vector<LNode> next;
            next.reserve((size_t)S * 8);
            for (int s = 0; s < S; ++s)
                for (int i = 0; i < 8; ++i) next.push_back({ out.first_child[split[s].node] + i, base[(size_t)s * 8 + i], counts[(size_t)s * 8 + i] });
            level.swap(next);
            cur = nxt; E = (int64_t)tot;
            cudaFree(d_cbox); cudaFree(d_ss); cudaFree(d_sc); cudaFree(d_slot); cudaFree(d_flags); cudaFree(d_scan); cudaFree(d_tiles); cudaFree(d_counts); cudaFree(d_base);
            return HARE_OK;
        };
        rcode = body();
        if (rcode) break;
    }
    if (rcode == HARE_OK) {
        // gather the leaf lists into their final places, then bring them to the host copy of the tree
        uint32_t* d_lists = nullptr;
        cudaError_t e = dmalloc(&d_lists, (size_t)list_total);
        for (int arr = 0; e == cudaSuccess && arr < (int)leaf_src.size(); ++arr) {
            std::vector<uint32_t> s, dd, cc;
            for (const LeafCopy& c2 : copies) if (c2.arr == arr) { s.push_back(c2.src); dd.push_back(c2.dst); cc.push_back(c2.cnt); }
            if (s.empty()) continue;
            uint32_t *ds = nullptr, *ddst = nullptr, *dc = nullptr;
            e = dmalloc(&ds, s.size()); if (e == cudaSuccess) e = dmalloc(&ddst, s.size()); if (e == cudaSuccess) e = dmalloc(&dc, s.size());
            if (e == cudaSuccess) e = cudaMemcpyAsync(ds, s.data(), s.size() * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(ddst, dd.data(), s.size() * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dc, cc.data(), s.size() * 4, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) { oct_copy_lists<<<(unsigned)std::min<size_t>(s.size(), 65535), 256, 0, st>>>(leaf_src[arr], ds, ddst, dc, (int)s.size(), d_lists); ++g_launches; e = cudaGetLastError(); }
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            cudaFree(ds); cudaFree(ddst); cudaFree(dc);
        }
        out.polys.resize(list_total);
        unsigned long long lost = 0;
        if (e == cudaSuccess && list_total) e = cudaMemcpy(out.polys.data(), d_lists, (size_t)list_total * 4, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(&lost, d_lost, 8, cudaMemcpyDeviceToHost);
        out.lost = (int64_t)lost;
        cudaFree(d_lists);
        if (e != cudaSuccess) rcode = fail(HARE_ERR_CUDA, std::string("build_octree_gpu: ") + cudaGetErrorString(e));
    }
    for (uint32_t* p2 : leaf_src) cudaFree(p2);
    if (rcode != HARE_OK && !leaf_src.empty() && cur != leaf_src.back()) cudaFree(cur);
    cudaFree(d_lost);
    return rcode;
}

extern "C" int hare_octree_build(hare_topo_t topo, int maxDepth, int maxPolys, hare_part_t* out) {
    if (!topo || !out || maxDepth < 0 || maxDepth >= HARE_OCT_MAXLVL) return fail(HARE_ERR_INVALID, "hare_octree_build: bad argument");
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_OCTREE, &p);
    if (rc) return rc;
    {
        const char* e = getenv("HARE_OCT_HOST_BUILD");
        if (p->dev.empty() || (e && *e == '1')) build_octree(topo->host, maxDepth, maxPolys, p->oct);   // host-only handles, or forced
        else {
            rc = build_octree_gpu(topo, p->dev[0].polys, p->dev[0].dev, p->dev[0].stream[0], maxDepth, maxPolys, p->oct);
            if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return rc; }
        }
    }
    rc = oct_to_device(p);
    if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return rc; }
    *out = p;
    return HARE_OK;
}

extern "C" int hare_octree_upload(hare_topo_t topo, const double* node_box, const int32_t* first_child,
                                  const uint32_t* list_off, const uint32_t* list_cnt, const uint32_t* polys,
                                  int64_t n_nodes, int64_t n_list, hare_part_t* out) {
    if (!topo || !node_box || !first_child || !list_off || !list_cnt || n_nodes < 1 || n_list < 0 || !out || (n_list && !polys))
        return fail(HARE_ERR_INVALID, "hare_octree_upload: bad argument");
    {   // a tree: children come after their parent (the packing passes and the kernels rely on it), every node but the root has
        // exactly one parent, and the depth stays inside the kernels' frame budget -- cycles and shared subtrees are rejected here
        std::vector<int32_t> level((size_t)n_nodes, -1);
        level[0] = 0;
        int64_t claimed = 1;
        for (int64_t i = 0; i < n_nodes; ++i) {
            if (first_child[i] >= 0) {
                const int64_t fc = first_child[i];
                if (fc <= i || fc + 8 > n_nodes) return fail(HARE_ERR_INVALID, "hare_octree_upload: children must follow their parent (i < first_child[i], first_child[i] + 8 <= n_nodes)");
                if (level[i] < 0) return fail(HARE_ERR_INVALID, "hare_octree_upload: node is not reachable from the root");
                if (level[i] + 1 >= HARE_OCT_MAXLVL) return fail(HARE_ERR_UNSUPPORTED, "hare_octree_upload: octree deeper than HARE_OCT_MAXLVL levels");
                for (int c = 0; c < 8; ++c) {
                    if (level[fc + c] >= 0) return fail(HARE_ERR_INVALID, "hare_octree_upload: a node has two parents");
                    level[fc + c] = level[i] + 1;
                }
                claimed += 8;
            } else if ((int64_t)list_off[i] + list_cnt[i] > n_list) return fail(HARE_ERR_INVALID, "hare_octree_upload: list range out of bounds");
        }
        if (claimed != n_nodes) return fail(HARE_ERR_INVALID, "hare_octree_upload: nodes that are not reachable from the root");
    }
    for (int64_t k = 0; k < n_list; ++k) if ((int64_t)polys[k] >= topo->host.P) return fail(HARE_ERR_INVALID, "hare_octree_upload: polygon index out of range");
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_OCTREE, &p);
    if (rc) return rc;
    p->oct.box.assign(node_box, node_box + 6 * n_nodes);
    p->oct.first_child.assign(first_child, first_child + n_nodes);
    p->oct.list_off.assign(list_off, list_off + n_nodes);
    p->oct.list_cnt.assign(list_cnt, list_cnt + n_nodes);
    if (n_list) p->oct.polys.assign(polys, polys + n_list);
    p->oct.depth = oct_depth(p->oct);
    rc = oct_to_device(p);
    if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return rc; }
    *out = p;
    return HARE_OK;
}

extern "C" int hare_octree_info(hare_part_t p, int64_t* n_nodes, int64_t* n_list, int64_t* lost, int32_t* depth) {
    if (!p || p->kind != HARE_OCTREE) return fail(HARE_ERR_INVALID, "hare_octree_info: not an Octree");
    if (n_nodes) *n_nodes = (int64_t)p->oct.first_child.size();
    if (n_list) *n_list = (int64_t)p->oct.polys.size();
    if (lost) *lost = p->oct.lost;
    if (depth) *depth = p->oct.depth;
    return HARE_OK;
}

extern "C" int hare_octree_download(hare_part_t p, double* node_box, int32_t* first_child, uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys) {
    if (!p || p->kind != HARE_OCTREE) return fail(HARE_ERR_INVALID, "hare_octree_download: not an Octree");
    const OctTree& t = p->oct;
    if (node_box) std::memcpy(node_box, t.box.data(), t.box.size() * 8);
    if (first_child) std::memcpy(first_child, t.first_child.data(), t.first_child.size() * 4);
    if (list_off) std::memcpy(list_off, t.list_off.data(), t.list_off.size() * 4);
    if (list_cnt) std::memcpy(list_cnt, t.list_cnt.data(), t.list_cnt.size() * 4);
    if (polys && !t.polys.empty()) std::memcpy(polys, t.polys.data(), t.polys.size() * 4);
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// KDTree
// ---------------------------------------------------------------------------------------
static int kd_depth(const KdTree& t) { return kd_depth_of(t); }

static int kd_to_device(hare_part_s* p) {
    const KdTree& t = p->kd;
    const size_t N = t.axis.size();
    if (kd_depth(t) + 2 > HARE_KD_MAXSTACK) return fail(HARE_ERR_UNSUPPORTED, "kd-tree deeper than HARE_KD_MAXSTACK allows");
    std::vector<KdNode> nodes; std::vector<KdNodeC> hot;
    pack_kdtree(t, p->topo->host, nodes);   // content-tightened node boxes (pack.hpp)
    pack_kdtree_hot(nodes, hot);            // FP32 padded boxes + links
    std::vector<KdWide> wide;
    pack_kdtree_wide(nodes, hot, wide);     // the walk's 128-byte records: a node's grandchildren
    for (PartDev& d : p->dev) {
        CK(cudaSetDevice(d.dev));
        KdNode* dn = nullptr;
        CK(dmalloc(&dn, N)); d.nodes = dn;
        CK(dmalloc(&d.kd_hot, N)); CK(cudaMemcpy(d.kd_hot, hot.data(), N * sizeof(KdNodeC), cudaMemcpyHostToDevice));
        CK(dmalloc(&d.kd_wide, N)); CK(cudaMemcpy(d.kd_wide, wide.data(), N * sizeof(KdWide), cudaMemcpyHostToDevice));
        CK(dmalloc(&d.lists, t.polys.size() + 8));
        CK(dmalloc(&d.ref_box, 6 * N));
        CK(cudaMemcpy(dn, nodes.data(), N * sizeof(KdNode), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d.ref_box, t.box.data(), 6 * N * sizeof(double), cudaMemcpyHostToDevice));   // the reference's boxes: tie rule only
        if (!t.polys.empty()) CK(cudaMemcpy(d.lists, t.polys.data(), t.polys.size() * 4, cudaMemcpyHostToDevice));
        { int r = tree_entry_boxes(d, t.polys.size(), p->topo->host.P, false); if (r) return r; }
        d.bytes = N * (sizeof(KdNode) + sizeof(KdNodeC) + sizeof(KdWide) + 48) + t.polys.size() * 36;
    }
    return HARE_OK;
}

extern "C" int hare_kdtree_build(hare_topo_t topo, int maxDepth, int maxPolys, hare_part_t* out) {
    if (!topo || !out || maxDepth < 0 || maxDepth + 2 > HARE_KD_MAXSTACK) return fail(HARE_ERR_INVALID, "hare_kdtree_build: bad argument");
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_KDTREE, &p);
    if (rc) return rc;
    {
        const char* e = getenv("HARE_KD_HOST_BUILD");
        if (p->dev.empty() || (e && *e == '1')) build_kdtree(topo->host, maxDepth, maxPolys, p->kd);   // host-only handles, or forced
        else {
            std::string msg;
            rc = build_kdtree_gpu(topo->host, p->dev[0].polys, p->dev[0].dev, p->dev[0].stream[0], maxDepth, maxPolys, p->kd, msg);
            if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return fail(HARE_ERR_CUDA, msg); }
        }
    }
    rc = kd_to_device(p);
    if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return rc; }
    *out = p;
    return HARE_OK;
}

extern "C" int hare_kdtree_upload(hare_topo_t topo, const double* node_box, const double* split, const int32_t* axis, const int32_t* left,
                                  const uint32_t* list_off, const uint32_t* list_cnt, const uint32_t* polys,
                                  int64_t n_nodes, int64_t n_list, hare_part_t* out) {
    if (!topo || !node_box || !split || !axis || !left || !list_off || !list_cnt || n_nodes < 1 || n_list < 0 || !out || (n_list && !polys))
        return fail(HARE_ERR_INVALID, "hare_kdtree_upload: bad argument");
    {   // same tree checks as hare_octree_upload
        std::vector<int32_t> level((size_t)n_nodes, -1);
        level[0] = 0;
        int64_t claimed = 1;
        for (int64_t i = 0; i < n_nodes; ++i) {
            if (left[i] >= 0) {
                const int64_t l = left[i];
                if (l <= i || l + 2 > n_nodes || axis[i] < 0 || axis[i] > 2) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: bad internal node (i < left[i], left[i] + 2 <= n_nodes, axis in 0..2)");
                if (level[i] < 0) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: node is not reachable from the root");
                if (level[i] + 3 > HARE_KD_MAXSTACK) return fail(HARE_ERR_UNSUPPORTED, "hare_kdtree_upload: kd-tree deeper than HARE_KD_MAXSTACK allows");
                if (level[l] >= 0 || level[l + 1] >= 0) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: a node has two parents");
                level[l] = level[l + 1] = level[i] + 1;
                claimed += 2;
            } else if ((int64_t)list_off[i] + list_cnt[i] > n_list) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: list range out of bounds");
        }
        if (claimed != n_nodes) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: nodes that are not reachable from the root");
    }
    for (int64_t k = 0; k < n_list; ++k) if ((int64_t)polys[k] >= topo->host.P) return fail(HARE_ERR_INVALID, "hare_kdtree_upload: polygon index out of range");
    hare_part_s* p = nullptr;
    int rc = new_part(topo, HARE_KDTREE, &p);
    if (rc) return rc;
    p->kd.box.assign(node_box, node_box + 6 * n_nodes);
    p->kd.split.assign(split, split + n_nodes);
    p->kd.axis.assign(axis, axis + n_nodes);
    p->kd.left.assign(left, left + n_nodes);
    p->kd.list_off.assign(list_off, list_off + n_nodes);
    p->kd.list_cnt.assign(list_cnt, list_cnt + n_nodes);
    if (n_list) p->kd.polys.assign(polys, polys + n_list);
    p->kd.depth = kd_depth(p->kd);
    rc = kd_to_device(p);
    if (rc) { for (auto& d : p->dev) free_partdev(d); drop_part(p); return rc; }
    *out = p;
    return HARE_OK;
}

extern "C" int hare_kdtree_info(hare_part_t p, int64_t* n_nodes, int64_t* n_list, int32_t* depth) {
    if (!p || p->kind != HARE_KDTREE) return fail(HARE_ERR_INVALID, "hare_kdtree_info: not a KDTree");
    if (n_nodes) *n_nodes = (int64_t)p->kd.axis.size();
    if (n_list) *n_list = (int64_t)p->kd.polys.size();
    if (depth) *depth = p->kd.depth;
    return HARE_OK;
}

extern "C" int hare_kdtree_download(hare_part_t p, double* node_box, double* split, int32_t* axis, int32_t* left,
                                    uint32_t* list_off, uint32_t* list_cnt, uint32_t* polys) {
    if (!p || p->kind != HARE_KDTREE) return fail(HARE_ERR_INVALID, "hare_kdtree_download: not a KDTree");
    const KdTree& t = p->kd;
    if (node_box) std::memcpy(node_box, t.box.data(), t.box.size() * 8);
    if (split) std::memcpy(split, t.split.data(), t.split.size() * 8);
    if (axis) std::memcpy(axis, t.axis.data(), t.axis.size() * 4);
    if (left) std::memcpy(left, t.left.data(), t.left.size() * 4);
    if (list_off) std::memcpy(list_off, t.list_off.data(), t.list_off.size() * 4);
    if (list_cnt) std::memcpy(list_cnt, t.list_cnt.data(), t.list_cnt.size() * 4);
    if (polys && !t.polys.empty()) std::memcpy(polys, t.polys.data(), t.polys.size() * 4);
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// On-disk form of a flattened partition (the reference has none): lets a large hall skip its
// rebuild.  Little-endian, native types:
//   "HAREB200" | u32 version | i32 kind | i64 P | u64 fnv1a(polygon vertices) | kind-specific arrays, each as (u64 count, data)
// Loading goes through the same checks and device set-up as the *_upload entry points.
// ---------------------------------------------------------------------------------------
namespace {
const char kMagic[8] = { 'H', 'A', 'R', 'E', 'B', '2', '0', '0' };
const uint32_t kFileVersion = 1;

uint64_t topo_fingerprint(const HostTopo& M) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(M.verts.data());
    for (size_t i = 0, n = M.verts.size() * sizeof(double); i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
template <class T> bool put(FILE* f, const T* v, uint64_t n) { return fwrite(&n, 8, 1, f) == 1 && (n == 0 || fwrite(v, sizeof(T), n, f) == n); }
template <class T> bool put(FILE* f, const std::vector<T>& v) { return put(f, v.data(), (uint64_t)v.size()); }
template <class T> bool get(FILE* f, std::vector<T>& v, uint64_t limit) {
    uint64_t n = 0;
    if (fread(&n, 8, 1, f) != 1 || n > limit) return false;
    v.resize((size_t)n);
    return n == 0 || fread(v.data(), sizeof(T), (size_t)n, f) == n;
}
}  // namespace

extern "C" int hare_part_save(hare_part_t p, const char* path) {
    if (!p || !path) return fail(HARE_ERR_INVALID, "hare_part_save: bad argument");
    std::lock_guard<std::mutex> lk(p->mu);
    std::vector<uint32_t> off, pol;
    if (p->kind == HARE_VOXEL_GRID) {
        if (p->dev.empty()) return fail(HARE_ERR_CUDA, "hare_part_save: host-only Voxel_Grid handle");
        const int64_t ncells = (int64_t)p->ct[0] * p->ct[1] * p->ct[2];
        off.resize((size_t)ncells + 1); pol.resize((size_t)p->npairs);
        PartDev& d = p->dev[0];
        CK(cudaSetDevice(d.dev));
        CK(cudaMemcpy(off.data(), d.cell_offset, off.size() * 4, cudaMemcpyDeviceToHost));
        if (p->npairs) CK(cudaMemcpy(pol.data(), d.cell_poly, pol.size() * 4, cudaMemcpyDeviceToHost));
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(HARE_ERR_INVALID, std::string("hare_part_save: cannot open ") + path);
    const int32_t kind = p->kind; const int64_t P = p->topo->host.P; const uint64_t fp = topo_fingerprint(p->topo->host);
    bool ok = fwrite(kMagic, 8, 1, f) == 1 && fwrite(&kFileVersion, 4, 1, f) == 1 && fwrite(&kind, 4, 1, f) == 1 && fwrite(&P, 8, 1, f) == 1 && fwrite(&fp, 8, 1, f) == 1;
    if (kind == HARE_VOXEL_GRID) ok = ok && put(f, p->obox, 6) && put(f, p->ct, 3) && put(f, off) && put(f, pol);
    else if (kind == HARE_OCTREE) ok = ok && put(f, p->oct.box) && put(f, p->oct.first_child) && put(f, p->oct.list_off) && put(f, p->oct.list_cnt) && put(f, p->oct.polys) && put(f, &p->oct.lost, 1);
    else ok = ok && put(f, p->kd.box) && put(f, p->kd.split) && put(f, p->kd.axis) && put(f, p->kd.left) && put(f, p->kd.list_off) && put(f, p->kd.list_cnt) && put(f, p->kd.polys);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(HARE_ERR_INVALID, std::string("hare_part_save: write failed: ") + path);
    return HARE_OK;
}

extern "C" int hare_part_load(hare_topo_t topo, const char* path, hare_part_t* out) {
    if (!topo || !path || !out) return fail(HARE_ERR_INVALID, "hare_part_load: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(HARE_ERR_INVALID, std::string("hare_part_load: cannot open ") + path);
    char magic[8]; uint32_t ver = 0; int32_t kind = 0; int64_t P = 0; uint64_t fp = 0;
    bool ok = fread(magic, 8, 1, f) == 1 && fread(&ver, 4, 1, f) == 1 && fread(&kind, 4, 1, f) == 1 && fread(&P, 8, 1, f) == 1 && fread(&fp, 8, 1, f) == 1;
    auto bail = [&](const std::string& why) { fclose(f); return fail(HARE_ERR_INVALID, "hare_part_load: " + why); };
    if (!ok || std::memcmp(magic, kMagic, 8) != 0) return bail("not a hare_b200 partition file");
    if (ver != kFileVersion) return bail("unsupported file version");
    if (P != topo->host.P || fp != topo_fingerprint(topo->host)) return bail("the file was written for a different Topology");
    const uint64_t lim = 1ull << 33;
    int rc;
    if (kind == HARE_VOXEL_GRID) {
        std::vector<double> obox; std::vector<int32_t> ct; std::vector<uint32_t> off, pol;
        if (!(get(f, obox, 6) && get(f, ct, 3) && get(f, off, lim) && get(f, pol, lim)) || obox.size() != 6 || ct.size() != 3) return bail("truncated Voxel_Grid record");
        if (ct[0] < 1 || ct[1] < 1 || ct[2] < 1 || (uint64_t)ct[0] * ct[1] * ct[2] + 1 != off.size() || off.back() != pol.size()) return bail("inconsistent Voxel_Grid record");
        fclose(f);
        rc = hare_voxelgrid_upload(topo, obox.data(), ct.data(), off.data(), pol.data(), out);
    } else if (kind == HARE_OCTREE) {
        OctTree t; std::vector<int64_t> lost;
        if (!(get(f, t.box, lim) && get(f, t.first_child, lim) && get(f, t.list_off, lim) && get(f, t.list_cnt, lim) && get(f, t.polys, lim) && get(f, lost, 1)) || lost.size() != 1)
            return bail("truncated Octree record");
        const size_t N = t.first_child.size();
        if (N < 1 || t.box.size() != 6 * N || t.list_off.size() != N || t.list_cnt.size() != N) return bail("inconsistent Octree record");
        fclose(f);
        rc = hare_octree_upload(topo, t.box.data(), t.first_child.data(), t.list_off.data(), t.list_cnt.data(), t.polys.data(), (int64_t)N, (int64_t)t.polys.size(), out);
        if (rc == HARE_OK) (*out)->oct.lost = lost[0];
    } else if (kind == HARE_KDTREE) {
        KdTree t;
        if (!(get(f, t.box, lim) && get(f, t.split, lim) && get(f, t.axis, lim) && get(f, t.left, lim) && get(f, t.list_off, lim) && get(f, t.list_cnt, lim) && get(f, t.polys, lim)))
            return bail("truncated KDTree record");
        const size_t N = t.axis.size();
        if (N < 1 || t.box.size() != 6 * N || t.split.size() != N || t.left.size() != N || t.list_off.size() != N || t.list_cnt.size() != N) return bail("inconsistent KDTree record");
        fclose(f);
        rc = hare_kdtree_upload(topo, t.box.data(), t.split.data(), t.axis.data(), t.left.data(), t.list_off.data(), t.list_cnt.data(), t.polys.data(), (int64_t)N, (int64_t)t.polys.size(), out);
    } else {
        return bail("unknown partition kind");
    }
    return rc;
}

extern "C" int hare_part_kind(hare_part_t p) { return p ? p->kind : HARE_ERR_INVALID; }
extern "C" int64_t hare_part_device_bytes(hare_part_t p) { return (p && !p->dev.empty()) ? (int64_t)p->dev[0].bytes : -1; }

extern "C" int hare_part_destroy(hare_part_t p) {
    if (!p) return HARE_OK;
    for (auto& d : p->dev) free_partdev(d);
    drop_part(p);
    return HARE_OK;
}

// ---------------------------------------------------------------------------------------
// Shoot
// ---------------------------------------------------------------------------------------
struct ShootArgs {
    const double *o, *d; const int32_t *o1, *o2, *rid; int64_t N;
    double *t, *xyz; int32_t* pid; double *uv, *om; unsigned long long* counters;
};

// Ray supply of a traversal launch of `total_warps` warps (RayFeed in vg_wave.cuh): block size and the counter the warps claim their
// rays from, zeroed in stream order
static int make_feed(int64_t N, int64_t total_warps, cudaStream_t st, RayFeedArgs* feed, int huge = HARE_FEED_BLOCK) {
    feed->block = feed_block_for(N, total_warps, huge);
    feed->first = total_warps * feed->block;
    feed->ctr = nullptr;
    CK(cudaMallocAsync(reinterpret_cast<void**>(&feed->ctr), sizeof(unsigned long long), st));
    cudaError_t e = cudaMemsetAsync(feed->ctr, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) { cudaFreeAsync(feed->ctr, st); CK(e); }
    return HARE_OK;
}

#ifndef HARE_WAVE_SLOTS
#define HARE_WAVE_SLOTS 64
#endif
#ifndef HARE_WAVE_WMAX
#define HARE_WAVE_WMAX 8
#endif
// Voxel_Grid: per-warp wavefront scheduler over shared-memory ray pools (vg_wave.cuh).
// One CTA of HARE_WAVE_WARPS warps per SM; dynamic shared memory = occupancy bitmap (when it fits) + the pools.
template <bool CHAIN, bool COUNT, bool OCC_SMEM>
static int launch_vg_wave2(const VGrid& g, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2,
                           const int32_t* rid, int64_t N, int order, const uint32_t* perm, const WalkOut& w, int warps, size_t smem, cudaStream_t st) {
    auto k = vg_wave_kernel<CHAIN, COUNT, OCC_SMEM, HARE_WAVE_SLOTS, HARE_WAVE_WMAX>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = warps * 32;
    int64_t blocks = std::min<int64_t>((N + threads - 1) / threads, (int64_t)d.sms);
    RayFeedArgs feed;
    int rc = make_feed(N, blocks * warps, st, &feed);
    if (rc) return rc;
    k<<<(unsigned)blocks, threads, smem, st>>>(g, d.polys, o, dd, o1, o2, rid, N, order, perm, feed, w);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(feed.ctr, st);
    CK(e);
    return HARE_OK;
}

static const size_t kSmemMax = 227 * 1024;

template <bool CHAIN>
static int launch_vg_wave(const VGrid& g, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2,
                          const int32_t* rid, int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    const size_t pool = WavePool<HARE_WAVE_SLOTS>::STRIDE;
    const size_t occ_bytes = ((((size_t)(g.nx + 2) * (g.ny + 2) * (g.nz + 2) + 31) / 32 + 3) & ~(size_t)3) * 4;   // padded grid, see vg_wave.cuh
    // the occupancy bitmap rides in shared memory next to the pools; a larger grid gives up warps for it (down to half),
    // and beyond that the bitmap is read through L1
    int warps = HARE_WAVE_WARPS;
    while (warps > HARE_WAVE_WARPS / 2 && occ_bytes + warps * pool > kSmemMax) --warps;
    const bool in_smem = occ_bytes + warps * pool <= kSmemMax;
    if (!in_smem) warps = HARE_WAVE_WARPS;
    const size_t smem = (in_smem ? occ_bytes : 0) + warps * pool;
    if (w.counters) {
        if (in_smem) return launch_vg_wave2<CHAIN, true, true>(g, d, o, dd, o1, o2, rid, N, order, perm, w, warps, smem, st);
        return launch_vg_wave2<CHAIN, true, false>(g, d, o, dd, o1, o2, rid, N, order, perm, w, warps, smem, st);
    }
    if (in_smem) return launch_vg_wave2<CHAIN, false, true>(g, d, o, dd, o1, o2, rid, N, order, perm, w, warps, smem, st);
    return launch_vg_wave2<CHAIN, false, false>(g, d, o, dd, o1, o2, rid, N, order, perm, w, warps, smem, st);
}

template <bool CHAIN>
static int launch_vg_walk(const VGrid& g, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2,
                          const int32_t* rid, int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    if (N <= 0) return HARE_OK;
    // the kernel packs ray numbers into 32 bits and the bounce into 16, and walks the border-padded bitmap (grids below 2^32 padded voxels)
    if (N >= (1LL << 32) || order >= 65536) return fail(HARE_ERR_INVALID, "Voxel_Grid Shoot: at most 2^32-1 rays and 65535 bounces per call");
    if (!g.occp) return fail(HARE_ERR_UNSUPPORTED, "Voxel_Grid Shoot: grid with 2^32 or more padded voxels");
    return launch_vg_wave<CHAIN>(g, d, o, dd, o1, o2, rid, N, order, perm, w, st);
}

#ifndef HARE_OCTW_SLOTS
#define HARE_OCTW_SLOTS 64
#endif
#ifndef HARE_OCTW_NMAX
#define HARE_OCTW_NMAX 4
#endif
#ifndef HARE_OCTW_FEED_HUGE
#define HARE_OCTW_FEED_HUGE 128   /* rays per RayFeed block for batches of ~39 M rays and more (100 M rays: 844 vs 829 Mrays/s with 64; 12.5 M: 708 vs 721) */
#endif

// Octree: per-warp wavefront scheduler over shared-memory ray pools (oct_wave.cuh), one CTA of HARE_OCTW_WARPS warps per SM.
// The spilled traversal frames live in a scratch area allocated stream-ordered for this launch (24 bytes per slot and level).
template <bool CHAIN, bool COUNT>
static int launch_oct_wave2(const OctDev& t, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2,
                            int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    auto k = oct_wave_kernel<CHAIN, COUNT, HARE_OCTW_SLOTS, HARE_OCTW_NMAX>;
    const size_t smem = (size_t)HARE_OCTW_WARPS * OctPool<HARE_OCTW_SLOTS>::STRIDE;
    static_assert((size_t)HARE_OCTW_WARPS * OctPool<HARE_OCTW_SLOTS>::STRIDE <= 227 * 1024, "ray pools exceed the shared memory of an SM");
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = HARE_OCTW_WARPS * 32;
    const int64_t blocks = std::min<int64_t>((N + threads - 1) / threads, (int64_t)d.sms);
    OctFrames F;
    F.depth = t.depth + 1;
    const size_t nfr = (size_t)blocks * HARE_OCTW_WARPS * HARE_OCTW_SLOTS * (size_t)F.depth;
    void* scratch = nullptr;
    CK(cudaMallocAsync(&scratch, nfr * (sizeof(double2) + sizeof(uint2)), st));
    F.ab = reinterpret_cast<double2*>(scratch); F.cq = reinterpret_cast<uint2*>(F.ab + nfr);
    RayFeedArgs feed;
    int rc = make_feed(N, blocks * HARE_OCTW_WARPS, st, &feed, HARE_OCTW_FEED_HUGE);
    if (rc) { cudaFreeAsync(scratch, st); return rc; }
    k<<<(unsigned)blocks, threads, smem, st>>>(t, F, d.polys, o, dd, o1, o2, N, order, perm, feed, w);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    cudaFreeAsync(feed.ctr, st);
    CK(e);
    return HARE_OK;
}

template <bool CHAIN>
static int launch_oct_walk(const OctDev& t, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2,
                           int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    if (N <= 0) return HARE_OK;
    if (N >= (1LL << 32) || order >= 65536) return fail(HARE_ERR_INVALID, "Octree Shoot: at most 2^32-1 rays and 65535 bounces per call");
    if (w.counters) return launch_oct_wave2<CHAIN, true>(t, d, o, dd, o1, o2, N, order, perm, w, st);
    return launch_oct_wave2<CHAIN, false>(t, d, o, dd, o1, o2, N, order, perm, w, st);
}

#ifndef HARE_KDW_SLOTS
#define HARE_KDW_SLOTS 64
#endif
#ifndef HARE_KDW_NMAX
#define HARE_KDW_NMAX 4
#endif

// KDTree: per-warp wavefront scheduler over shared-memory ray pools (kd_wave.cuh); the stacks of pending far children live in a
// scratch area allocated stream-ordered for this launch (4 bytes per slot and level)
template <bool CHAIN, bool COUNT>
static int launch_kd_wave2(const KdDev& t, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2, const int32_t* rid,
                           int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    auto k = kd_wave_kernel<CHAIN, COUNT, HARE_KDW_SLOTS, HARE_KDW_NMAX>;
    const size_t smem = (size_t)HARE_KDW_WARPS * KdPool<HARE_KDW_SLOTS>::STRIDE;
    static_assert((size_t)HARE_KDW_WARPS * KdPool<HARE_KDW_SLOTS>::STRIDE <= 227 * 1024, "ray pools exceed the shared memory of an SM");
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = HARE_KDW_WARPS * 32;
    const int64_t blocks = std::min<int64_t>((N + threads - 1) / threads, (int64_t)d.sms);
    KdStacks S;
    S.depth = 3 * (t.depth / 2 + 2) + 4;     // a record pushes at most three entries, one record per two levels
    const size_t n = (size_t)blocks * HARE_KDW_WARPS * HARE_KDW_SLOTS * (size_t)S.depth;
    void* scratch = nullptr;
    CK(cudaMallocAsync(&scratch, n * sizeof(uint4), st));
    S.st = reinterpret_cast<uint4*>(scratch);
    RayFeedArgs feed;
    int rc = make_feed(N, blocks * HARE_KDW_WARPS, st, &feed);
    if (rc) { cudaFreeAsync(scratch, st); return rc; }
    k<<<(unsigned)blocks, threads, smem, st>>>(t, S, d.polys, o, dd, o1, o2, rid, N, order, perm, feed, w);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    cudaFreeAsync(feed.ctr, st);
    CK(e);
    return HARE_OK;
}

template <bool CHAIN>
static int launch_kd_walk(const KdDev& t, const PartDev& d, const double* o, const double* dd, const int32_t* o1, const int32_t* o2, const int32_t* rid,
                          int64_t N, int order, const uint32_t* perm, const WalkOut& w, cudaStream_t st) {
    if (N <= 0) return HARE_OK;
    if (N >= (1LL << 32) || order >= 65536) return fail(HARE_ERR_INVALID, "KDTree Shoot: at most 2^32-1 rays and 65535 bounces per call");
    if (w.counters) return launch_kd_wave2<CHAIN, true>(t, d, o, dd, o1, o2, rid, N, order, perm, w, st);
    return launch_kd_wave2<CHAIN, false>(t, d, o, dd, o1, o2, rid, N, order, perm, w, st);
}

// Coherence pre-pass (ray_bin.cuh): a permutation of the batch grouped by origin cell and direction cell, in `*perm` (stream-ordered
// allocation, released by the caller after the traversal launch).  Small batches are shot in the caller's order.
static const int64_t kRayBinMin = 1 << 16;
static int bin_rays(hare_part_s* p, const PartDev& d, const double* o, const double* dd, int64_t N, cudaStream_t st, uint32_t** perm) {
    *perm = nullptr;
    if (N < kRayBinMin) return HARE_OK;
    const HostTopo& M = p->topo->host;
    RayBinGeom g;
    g.ox = (float)M.minmax[0]; g.oy = (float)M.minmax[1]; g.oz = (float)M.minmax[2];
    g.sx = 4.0f / std::max((float)(M.minmax[3] - M.minmax[0]), 1e-30f); g.sy = 4.0f / std::max((float)(M.minmax[4] - M.minmax[1]), 1e-30f);
    g.sz = 4.0f / std::max((float)(M.minmax[5] - M.minmax[2]), 1e-30f);
    g.dirbits = ray_bin_dirbits(N);
    StreamTemps tmp(st);
    uint32_t *keys = nullptr, *counts = nullptr, *offs = nullptr, *tiles = nullptr, *pm = nullptr;
    const int64_t nb = ray_bin_buckets(g.dirbits);
    CK(tmp.get(&keys, (size_t)N)); CK(tmp.get(&counts, (size_t)nb)); CK(tmp.get(&offs, (size_t)nb + 1));
    CK(tmp.get(&tiles, (size_t)((nb + HARE_SCAN_TILE - 1) / HARE_SCAN_TILE + 1))); CK(tmp.get(&pm, (size_t)N));
    CK(cudaMemsetAsync(counts, 0, (size_t)nb * 4, st));
    const unsigned blocks = (unsigned)std::min<int64_t>((N + 255) / 256, (int64_t)d.sms * 16);
    ray_bin_count<<<blocks, 256, 0, st>>>(o, dd, N, g, keys, counts);
    ++g_launches;
    int r = scan_u32(counts, nb, offs, tiles, st);
    if (r) return r;
    ray_bin_scatter<<<blocks, 256, 0, st>>>(keys, N, offs, pm);
    ++g_launches;
    CK(cudaGetLastError());
    tmp.keep(pm);
    *perm = pm;
    return HARE_OK;
}

static int launch_shoot(hare_part_s* p, const PartDev& d, const ShootArgs& a, cudaStream_t st) {
    if (a.N >= (1LL << 32)) return fail(HARE_ERR_INVALID, "Shoot: at most 2^32-1 rays per call");
    uint32_t* perm = nullptr;
    int rc = bin_rays(p, d, a.o, a.d, a.N, st, &perm);
    if (rc) return rc;
    WalkOut w = { a.t, a.xyz, a.pid, a.uv, a.om, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, a.counters, nullptr, nullptr };
    switch (p->kind) {
        case HARE_VOXEL_GRID:
            rc = launch_vg_walk<false>(make_vgrid(p, d), d, a.o, a.d, a.o1, a.o2, a.rid, a.N, 1, perm, w, st);
            break;
        case HARE_OCTREE: {
            OctDev t = { (const OctNode*)d.nodes, d.lists, d.cbox, d.gbox, d.pbox, d.nbox, p->oct.depth, p->oct_regular ? 1 : 0 };
            rc = launch_oct_walk<false>(t, d, a.o, a.d, a.o1, a.o2, a.N, 1, perm, w, st);
            break;
        }
        case HARE_KDTREE: {
            KdDev t = { d.kd_wide, d.kd_hot, (const KdNode*)d.nodes, d.lists, d.tbox, p->kd.depth, d.ref_box };
            rc = launch_kd_walk<false>(t, d, a.o, a.d, a.o1, a.o2, a.rid, a.N, 1, perm, w, st);
            break;
        }
        default: rc = fail(HARE_ERR_INVALID, "unknown partition kind");
    }
    if (perm) cudaFreeAsync(perm, st);
    return rc;
}

struct ChainArgs {
    const double *o, *d; int64_t N; int order;
    int32_t* ev_pid; double* ev_t; double *fin_o, *fin_d; int32_t* nshots;
    unsigned long long *total, *counters;
    double *ev_xyz, *ev_uv;
};

static int launch_chain(hare_part_s* p, const PartDev& d, const ChainArgs& a, cudaStream_t st) {
    switch (p->kind) {
        case HARE_VOXEL_GRID: {
            WalkOut w = { nullptr, nullptr, nullptr, nullptr, nullptr, a.ev_pid, a.ev_t, a.fin_o, a.fin_d, a.nshots, a.total, a.counters, a.ev_xyz, a.ev_uv };
            return launch_vg_walk<true>(make_vgrid(p, d), d, a.o, a.d, nullptr, nullptr, nullptr, a.N, a.order, nullptr, w, st);
        }
        case HARE_OCTREE: {
            OctDev t = { (const OctNode*)d.nodes, d.lists, d.cbox, d.gbox, d.pbox, d.nbox, p->oct.depth, p->oct_regular ? 1 : 0 };
            WalkOut w = { nullptr, nullptr, nullptr, nullptr, nullptr, a.ev_pid, a.ev_t, a.fin_o, a.fin_d, a.nshots, a.total, a.counters, a.ev_xyz, a.ev_uv };
            return launch_oct_walk<true>(t, d, a.o, a.d, nullptr, nullptr, a.N, a.order, nullptr, w, st);
        }
        case HARE_KDTREE: {
            KdDev t = { d.kd_wide, d.kd_hot, (const KdNode*)d.nodes, d.lists, d.tbox, p->kd.depth, d.ref_box };
            WalkOut w = { nullptr, nullptr, nullptr, nullptr, nullptr, a.ev_pid, a.ev_t, a.fin_o, a.fin_d, a.nshots, a.total, a.counters, a.ev_xyz, a.ev_uv };
            return launch_kd_walk<true>(t, d, a.o, a.d, nullptr, nullptr, nullptr, a.N, a.order, nullptr, w, st);
        }
    }
    return fail(HARE_ERR_INVALID, "unknown partition kind");
}

static const int64_t kChunk = 1 << 20;   // chains per pipelined chunk of hare_reflect_chain

// CUDA call inside a lambda that reports through an int status (the caller drains every stream before returning it)
#define CKS(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(HARE_ERR_CUDA, b_);                                                        \
        }                                                                                          \
    } while (0)

// Wait for every stream of every device: called on EVERY exit path of the host-buffer entry points, so that no asynchronous copy
// into the caller's arrays is still in flight when the call returns (the caller may free them right after an error).
static int drain_streams(hare_part_s* p) {
    int rc = HARE_OK;
    for (PartDev& dv : p->dev) {
        if (cudaSetDevice(dv.dev) != cudaSuccess) { rc = HARE_ERR_CUDA; continue; }
        for (int s = 0; s < kStreams; ++s) if (cudaStreamSynchronize(dv.stream[s]) != cudaSuccess) rc = HARE_ERR_CUDA;
    }
    return rc;
}

extern "C" int hare_shoot_batch(hare_part_t p, const double* o, const double* d,
                                const int32_t* origin1, const int32_t* origin2, const int32_t* ray_id, int64_t N,
                                double* t, double* xyz, int32_t* poly_id, double* uv, double* o_moved, uint64_t* counters) {
    if (!p || N < 0 || (N && (!o || !d || !poly_id))) return fail(HARE_ERR_INVALID, "hare_shoot_batch: bad argument");
    std::lock_guard<std::mutex> lk(p->mu);
    const int G = (int)p->dev.size();
    if (G == 0) return fail(HARE_ERR_CUDA, "hare_shoot_batch: host-only handle, no CUDA device (hare_b200 has no CPU fallback)");
    if (counters) std::memset(counters, 0, HARE_CNT_N * sizeof(uint64_t));
    // block-shard the batch over the devices; per device, pipeline chunks over kStreams streams
    auto enqueue = [&]() -> int {
        for (int g = 0; g < G; ++g) {
            PartDev& dv = p->dev[g];
            const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;
            if (r1 <= r0) continue;
            const std::vector<int64_t> sched = shoot_schedule(r1 - r0);
            int rc = ensure_staging(dv, *std::max_element(sched.begin(), sched.end()),
                                    (origin1 ? ST_O1 : 0u) | (origin2 ? ST_O2 : 0u) | (ray_id ? ST_RID : 0u) | (t ? ST_T : 0u) | (xyz ? ST_XYZ : 0u) |
                                    (uv ? ST_UV : 0u) | (o_moved ? ST_OM : 0u));
            if (rc) return rc;
            CKS(cudaSetDevice(dv.dev));
            if (counters) CKS(cudaMemsetAsync(dv.counters, 0, 8 * sizeof(unsigned long long), dv.stream[0]));
            if (counters) CKS(cudaStreamSynchronize(dv.stream[0]));
            int s = 0;
            int64_t c0 = r0;
            for (size_t k = 0; k < sched.size(); c0 += sched[k], ++k, s = (s + 1) % kStreams) {
                const int64_t n = sched[k];
                if (n <= 0) continue;
                cudaStream_t st = dv.stream[s];
                CKS(cudaMemcpyAsync(dv.s_o[s], o + 3 * c0, n * 24, cudaMemcpyHostToDevice, st));
                CKS(cudaMemcpyAsync(dv.s_d[s], d + 3 * c0, n * 24, cudaMemcpyHostToDevice, st));
                if (origin1) CKS(cudaMemcpyAsync(dv.s_o1[s], origin1 + c0, n * 4, cudaMemcpyHostToDevice, st));
                if (origin2) CKS(cudaMemcpyAsync(dv.s_o2[s], origin2 + c0, n * 4, cudaMemcpyHostToDevice, st));
                if (ray_id) CKS(cudaMemcpyAsync(dv.s_rid[s], ray_id + c0, n * 4, cudaMemcpyHostToDevice, st));
                ShootArgs a = { dv.s_o[s], dv.s_d[s], origin1 ? dv.s_o1[s] : nullptr, origin2 ? dv.s_o2[s] : nullptr, ray_id ? dv.s_rid[s] : nullptr, n,
                                t ? dv.s_t[s] : nullptr, xyz ? dv.s_xyz[s] : nullptr, dv.s_pid[s], uv ? dv.s_uv[s] : nullptr,
                                o_moved ? dv.s_om[s] : nullptr, counters ? dv.counters : nullptr };
                int rc2 = launch_shoot(p, dv, a, st);
                if (rc2) return rc2;
                CKS(cudaMemcpyAsync(poly_id + c0, dv.s_pid[s], n * 4, cudaMemcpyDeviceToHost, st));
                if (t) CKS(cudaMemcpyAsync(t + c0, dv.s_t[s], n * 8, cudaMemcpyDeviceToHost, st));
                if (xyz) CKS(cudaMemcpyAsync(xyz + 3 * c0, dv.s_xyz[s], n * 24, cudaMemcpyDeviceToHost, st));
                if (uv) CKS(cudaMemcpyAsync(uv + 2 * c0, dv.s_uv[s], n * 16, cudaMemcpyDeviceToHost, st));
                if (o_moved) CKS(cudaMemcpyAsync(o_moved + 3 * c0, dv.s_om[s], n * 24, cudaMemcpyDeviceToHost, st));
            }
        }
        return HARE_OK;
    };
    const int rc = enqueue();
    const int rs = drain_streams(p);
    if (rc) return rc;
    if (rs) return fail(HARE_ERR_CUDA, std::string("hare_shoot_batch: ") + cudaGetErrorString(cudaGetLastError()));
    if (counters) {
        for (int g = 0; g < G; ++g) {
            PartDev& dv = p->dev[g];
            CK(cudaSetDevice(dv.dev));
            unsigned long long h[4];
            CK(cudaMemcpy(h, dv.counters, sizeof h, cudaMemcpyDeviceToHost));
            for (int k = 0; k < 4; ++k) counters[k] += h[k];
        }
    }
    return HARE_OK;
}

extern "C" int hare_shoot_batch_device(hare_part_t p, const double* o, const double* d,
                                       const int32_t* origin1, const int32_t* origin2, const int32_t* ray_id, int64_t N,
                                       double* t, double* xyz, int32_t* poly_id, double* uv, double* o_moved,
                                       uint64_t* counters_device, void* cuda_stream) {
    if (!p || N < 0 || (N && (!o || !d || !poly_id))) return fail(HARE_ERR_INVALID, "hare_shoot_batch_device: bad argument");
    if (p->dev.size() != 1) return fail(p->dev.empty() ? HARE_ERR_CUDA : HARE_ERR_INVALID, "hare_shoot_batch_device: needs a single-device partition handle");
    PartDev& dv = p->dev[0];
    CK(cudaSetDevice(dv.dev));
    ShootArgs a = { o, d, origin1, origin2, ray_id, N, t, xyz, poly_id, uv, o_moved, (unsigned long long*)counters_device };
    return launch_shoot(p, dv, a, cuda_stream ? (cudaStream_t)cuda_stream : dv.stream[0]);
}

extern "C" int hare_reflect_chain(hare_part_t p, const double* o, const double* d, int64_t N, int order,
                                  int32_t* ev_poly_id, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots,
                                  uint64_t* total_shots, uint64_t* counters) {
    return hare_reflect_chain_events(p, o, d, N, order, ev_poly_id, ev_t, nullptr, nullptr, fin_o, fin_d, nshots, total_shots, counters);
}

extern "C" int hare_reflect_chain_events(hare_part_t p, const double* o, const double* d, int64_t N, int order,
                                         int32_t* ev_poly_id, double* ev_t, double* ev_xyz, double* ev_uv, double* fin_o, double* fin_d,
                                         int32_t* nshots, uint64_t* total_shots, uint64_t* counters) {
    if (!p || N < 0 || order < 1 || (N && (!o || !d))) return fail(HARE_ERR_INVALID, "hare_reflect_chain: bad argument");
    std::lock_guard<std::mutex> lk(p->mu);
    const int G = (int)p->dev.size();
    if (G == 0) return fail(HARE_ERR_CUDA, "hare_reflect_chain: host-only handle, no CUDA device (hare_b200 has no CPU fallback)");
    const bool events = ev_poly_id || ev_t || ev_xyz || ev_uv;
    const int64_t chunk = events ? std::max<int64_t>(1024, kChunk / order) : kChunk;
    if (counters) std::memset(counters, 0, HARE_CNT_N * sizeof(uint64_t));
    if (total_shots) *total_shots = 0;
    auto enqueue = [&]() -> int {
        for (int g = 0; g < G; ++g) {
            PartDev& dv = p->dev[g];
            const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;
            if (r1 <= r0) continue;
            int rc = ensure_staging(dv, std::min<int64_t>(chunk, r1 - r0), (fin_o ? ST_XYZ : 0u) | (fin_d ? ST_OM : 0u));
            if (rc) return rc;
            if (events) { rc = ensure_chain_staging(dv, std::min<int64_t>(chunk, r1 - r0), order); if (rc) return rc; }
            else if (nshots) { rc = ensure_chain_staging(dv, std::min<int64_t>(chunk, r1 - r0), 1); if (rc) return rc; }
            if (ev_xyz || ev_uv) { rc = ensure_chain_rows(dv, std::min<int64_t>(chunk, r1 - r0), order, ev_xyz != nullptr, ev_uv != nullptr); if (rc) return rc; }
            CKS(cudaSetDevice(dv.dev));
            CKS(cudaMemsetAsync(dv.counters, 0, 8 * sizeof(unsigned long long), dv.stream[0]));
            CKS(cudaStreamSynchronize(dv.stream[0]));
            int s = 0;
            for (int64_t c0 = r0; c0 < r1; c0 += chunk, s = (s + 1) % kStreams) {
                const int64_t n = std::min<int64_t>(chunk, r1 - c0);
                cudaStream_t st = dv.stream[s];
                CKS(cudaMemcpyAsync(dv.s_o[s], o + 3 * c0, n * 24, cudaMemcpyHostToDevice, st));
                CKS(cudaMemcpyAsync(dv.s_d[s], d + 3 * c0, n * 24, cudaMemcpyHostToDevice, st));
                ChainArgs a = { dv.s_o[s], dv.s_d[s], n, order, ev_poly_id ? dv.c_evpid[s] : nullptr, ev_t ? dv.c_evt[s] : nullptr,
                                fin_o ? dv.s_xyz[s] : nullptr, fin_d ? dv.s_om[s] : nullptr, nshots ? dv.c_ns[s] : nullptr,
                                dv.counters + 4, counters ? dv.counters : nullptr, ev_xyz ? dv.c_evxyz[s] : nullptr, ev_uv ? dv.c_evuv[s] : nullptr };
                int rc2 = launch_chain(p, dv, a, st);
                if (rc2) return rc2;
                if (ev_xyz) CKS(cudaMemcpyAsync(ev_xyz + 3 * c0 * order, dv.c_evxyz[s], (size_t)n * order * 24, cudaMemcpyDeviceToHost, st));
                if (ev_uv) CKS(cudaMemcpyAsync(ev_uv + 2 * c0 * order, dv.c_evuv[s], (size_t)n * order * 16, cudaMemcpyDeviceToHost, st));
                if (ev_poly_id) CKS(cudaMemcpyAsync(ev_poly_id + c0 * order, dv.c_evpid[s], (size_t)n * order * 4, cudaMemcpyDeviceToHost, st));
                if (ev_t) CKS(cudaMemcpyAsync(ev_t + c0 * order, dv.c_evt[s], (size_t)n * order * 8, cudaMemcpyDeviceToHost, st));
                if (fin_o) CKS(cudaMemcpyAsync(fin_o + 3 * c0, dv.s_xyz[s], n * 24, cudaMemcpyDeviceToHost, st));
                if (fin_d) CKS(cudaMemcpyAsync(fin_d + 3 * c0, dv.s_om[s], n * 24, cudaMemcpyDeviceToHost, st));
                if (nshots) CKS(cudaMemcpyAsync(nshots + c0, dv.c_ns[s], n * 4, cudaMemcpyDeviceToHost, st));
            }
        }
        return HARE_OK;
    };
    const int rc = enqueue();
    const int rs = drain_streams(p);
    if (rc) return rc;
    if (rs) return fail(HARE_ERR_CUDA, std::string("hare_reflect_chain: ") + cudaGetErrorString(cudaGetLastError()));
    for (int g = 0; g < G; ++g) {
        PartDev& dv = p->dev[g];
        CK(cudaSetDevice(dv.dev));
        unsigned long long h[5];
        CK(cudaMemcpy(h, dv.counters, sizeof h, cudaMemcpyDeviceToHost));
        if (counters) for (int k = 0; k < 4; ++k) counters[k] += h[k];
        if (total_shots) *total_shots += h[4];
    }
    return HARE_OK;
}

extern "C" int hare_reflect_chain_device(hare_part_t p, const double* o, const double* d, int64_t N, int order,
                                         int32_t* ev_poly_id, double* ev_t, double* fin_o, double* fin_d, int32_t* nshots,
                                         uint64_t* total_shots_device, uint64_t* counters_device, void* cuda_stream) {
    return hare_reflect_chain_events_device(p, o, d, N, order, ev_poly_id, ev_t, nullptr, nullptr, fin_o, fin_d, nshots, total_shots_device,
                                            counters_device, cuda_stream);
}

extern "C" int hare_reflect_chain_events_device(hare_part_t p, const double* o, const double* d, int64_t N, int order,
                                                int32_t* ev_poly_id, double* ev_t, double* ev_xyz, double* ev_uv, double* fin_o, double* fin_d,
                                                int32_t* nshots, uint64_t* total_shots_device, uint64_t* counters_device, void* cuda_stream) {
    if (!p || N < 0 || order < 1 || (N && (!o || !d)) || !total_shots_device) return fail(HARE_ERR_INVALID, "hare_reflect_chain_device: bad argument");
    if (p->dev.size() != 1) return fail(p->dev.empty() ? HARE_ERR_CUDA : HARE_ERR_INVALID, "hare_reflect_chain_device: needs a single-device partition handle");
    PartDev& dv = p->dev[0];
    CK(cudaSetDevice(dv.dev));
    ChainArgs a = { o, d, N, order, ev_poly_id, ev_t, fin_o, fin_d, nshots, (unsigned long long*)total_shots_device, (unsigned long long*)counters_device,
                    ev_xyz, ev_uv };
    return launch_chain(p, dv, a, cuda_stream ? (cudaStream_t)cuda_stream : dv.stream[0]);
}

// ---------------------------------------------------------------------------------------
// host / device buffers, IPC export of result buffers (include/hare_b200.h)
// ---------------------------------------------------------------------------------------
extern "C" int hare_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(HARE_ERR_INVALID, "hare_host_alloc: null argument");
    int rc = ensure_init(); if (rc) return rc;
    if (g_host_only) return fail(HARE_ERR_CUDA, "hare_host_alloc: host-only mode, no CUDA device");
    *out = nullptr;
    CK(cudaHostAlloc(out, std::max<size_t>(bytes, 1), cudaHostAllocPortable));
    return HARE_OK;
}
extern "C" int hare_host_free(void* p) { if (p) CK(cudaFreeHost(p)); return HARE_OK; }
extern "C" int hare_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(HARE_ERR_INVALID, "hare_host_register: null argument");
    int rc = ensure_init(); if (rc) return rc;
    if (g_host_only) return fail(HARE_ERR_CUDA, "hare_host_register: host-only mode, no CUDA device");
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return HARE_OK;
}
extern "C" int hare_host_unregister(void* p) { if (p) CK(cudaHostUnregister(p)); return HARE_OK; }
extern "C" int hare_host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

extern "C" int hare_device_alloc(int device, size_t bytes, void** out) {
    if (!out) return fail(HARE_ERR_INVALID, "hare_device_alloc: null argument");
    *out = nullptr;
    CK(cudaSetDevice(device));
    CK(cudaMalloc(out, std::max<size_t>(bytes, 1)));
    return HARE_OK;
}
extern "C" int hare_device_free(int device, void* p) { if (p) { CK(cudaSetDevice(device)); CK(cudaFree(p)); } return HARE_OK; }
extern "C" int hare_device_memcpy(void* dst, const void* src, size_t bytes, int kind, int device) {
    if (!dst || !src) return fail(HARE_ERR_INVALID, "hare_device_memcpy: null argument");
    if (kind < 1 || kind > 3) return fail(HARE_ERR_INVALID, "hare_device_memcpy: kind must be 1 (H2D), 2 (D2H) or 3 (D2D)");
    CK(cudaSetDevice(device));
    CK(cudaMemcpy(dst, src, bytes, kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice)));
    return HARE_OK;
}
extern "C" int hare_ipc_export(int device, void* dev_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!dev_ptr || !handle) return fail(HARE_ERR_INVALID, "hare_ipc_export: null argument");
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle, &h, 64);
    return HARE_OK;
}
extern "C" int hare_ipc_open(int device, const unsigned char handle[64], void** out) {
    if (!handle || !out) return fail(HARE_ERR_INVALID, "hare_ipc_open: null argument");
    *out = nullptr;
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return HARE_OK;
}
extern "C" int hare_ipc_close(int device, void* p) { if (p) { CK(cudaSetDevice(device)); CK(cudaIpcCloseMemHandle(p)); } return HARE_OK; }
