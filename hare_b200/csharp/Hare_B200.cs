// Hare_B200.cs -- binding of libhare_b200 (include/hare_b200.h) for the Hare_NC (.NET) build.
//
// Two ways to use it (INTEGRATION.md has the csproj lines):
//
//   * DROP-IN (define the compile symbol HARE_B200 and exclude Voxel_Grid.cs, "Octree - alt.cs" and KDTree.cs from the build):
//     this file then defines Hare.Geometry.Voxel_Grid, Octree and KDTree themselves -- same names, same constructors
//     (Voxel_Grid.cs:48, :128; "Octree - alt.cs":45; KDTree.cs:51), both Spatial_Partition.Shoot overloads
//     (Spatial_Partition.cs:32-33) and Voxel_Grid's kept public surface (Poly_Ray_ID, Voxel_Inv, VoxelCode / VoxelDecode,
//     PointInVoxel x2, Fill_Voxels, Xdim / Ydim / Zdim / MinPt; Voxel_Grid.cs:29, 33, 256-267, 273, 322-332, 763-791).
//     Pachyderm compiles unchanged and gains the batched overload
//         bool[] Shoot(Ray[] R, int top_index, out X_Event[] events, int[] poly_origin1 = null, int[] poly_origin2 = null).
//   * SIDE BY SIDE (symbol not defined): the same classes are named Gpu_Voxel_Grid, Gpu_Octree, Gpu_KDTree and live next to
//     the CPU ones.
//
// Host memory: a C# `fixed` block or GCHandle.Alloc(Pinned) pins an array for the garbage collector; it does NOT page-lock it
// for CUDA, and copies from pageable memory are staged and synchronous.  The classes below therefore keep their own reusable
// staging arrays, pinned for the GC AND registered with hare_host_register (page-locked) once per growth, so that the library's
// two-stream copy / compute pipeline really overlaps.
//
// Cliff to know about: Shoot(Ray, ...) for ONE ray is a one-element batch -- H2D copy, kernel launch, D2H copy and a stream
// synchronisation per call, ~100 us (bench.py --config C1 prints the measured figure), i.e. ~10^4 Shoots/s against the CPU
// reference's ~10^6 per thread.  It exists for compatibility; callers that loop over rays must move to the batched overload
// (or to hare_reflect_chain, which keeps whole reflection chains on the device).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no .NET toolchain); kept deliberately thin: it only flattens and marshals.  All
// arithmetic that decides a result lives behind the C ABI.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;

namespace Hare.Geometry.Native
{
    internal static class HareB200
    {
        const string Lib = "hare_b200";
        const CallingConvention CC = CallingConvention.Cdecl;

        public const int HARE_MISS = -1, HARE_RAY_FAULT = -2;

        [DllImport(Lib, CallingConvention = CC)] public static extern IntPtr hare_version();
        [DllImport(Lib, CallingConvention = CC)] public static extern IntPtr hare_last_error();
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_device_count();
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_init(int[] device_ids, int n_devices);

        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_topology_create(double[] verts, double[] normals, int[] vcount, long P, double[] minmax, out IntPtr topo);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_topology_destroy(IntPtr topo);

        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_voxelgrid_build(IntPtr topo, int domain, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_voxelgrid_build_adaptive(IntPtr topo, int max_domain_log2, int avg_polys, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_voxelgrid_upload(IntPtr topo, double[] obox, int[] ct, uint[] cell_offset, uint[] cell_poly, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_voxelgrid_info(IntPtr part, double[] obox, double[] voxeldims, int[] ct, out long npairs);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_voxelgrid_download(IntPtr part, uint[] cell_offset, uint[] cell_poly);

        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_octree_build(IntPtr topo, int maxDepth, int maxPolys, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_octree_upload(IntPtr topo, double[] node_box, int[] first_child, uint[] list_off, uint[] list_cnt, uint[] polys,
                                                    long n_nodes, long n_list, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_kdtree_build(IntPtr topo, int maxDepth, int maxPolys, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_kdtree_upload(IntPtr topo, double[] node_box, double[] split, int[] axis, int[] left, uint[] list_off, uint[] list_cnt,
                                                    uint[] polys, long n_nodes, long n_list, out IntPtr part);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_part_destroy(IntPtr part);
        [DllImport(Lib, CallingConvention = CC, CharSet = CharSet.Ansi)] public static extern int hare_part_save(IntPtr part, string path);
        [DllImport(Lib, CallingConvention = CC, CharSet = CharSet.Ansi)] public static extern int hare_part_load(IntPtr topo, string path, out IntPtr part);

        // page-locked host memory (include/hare_b200.h "host and device buffers")
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_host_alloc(UIntPtr bytes, out IntPtr p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_host_free(IntPtr p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_host_register(IntPtr p, UIntPtr bytes);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_host_unregister(IntPtr p);
        [DllImport(Lib, CallingConvention = CC)] public static extern int hare_host_is_pinned(IntPtr p);

        // pointer forms: the arrays are this file's own page-locked staging arrays
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_shoot_batch(IntPtr part, IntPtr o, IntPtr d, IntPtr origin1, IntPtr origin2, IntPtr ray_id, long N,
                                                  IntPtr t, IntPtr xyz, IntPtr poly_id, IntPtr uv, IntPtr o_moved, ulong[] counters);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_reflect_chain(IntPtr part, double[] o, double[] d, long N, int order, int[] ev_poly_id, double[] ev_t,
                                                    double[] fin_o, double[] fin_d, int[] nshots, out ulong total_shots, ulong[] counters);
        // + per-bounce X_Point (N x order x 3) and u, v (N x order x 2); any event array may be null
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_reflect_chain_events(IntPtr part, double[] o, double[] d, long N, int order, int[] ev_poly_id, double[] ev_t,
                                                           double[] ev_xyz, double[] ev_uv, double[] fin_o, double[] fin_d, int[] nshots,
                                                           out ulong total_shots, ulong[] counters);

        public static void Check(int rc, string what)
        {
            if (rc == 0) return;
            string msg = Marshal.PtrToStringAnsi(hare_last_error());
            if (rc == -3) throw new NotImplementedException(msg);   // same exception type as Hare_Geometry_Topology.cs:245-248
            throw new InvalidOperationException(what + " failed (" + rc + "): " + msg);
        }

        /// <summary>Flattens Model[0] the way the kernels want it: Polys[i].Points (4 x xyz, a triangle
        /// repeats vertex 2), Polys[i].Normal, Polys[i].VertextCT, Topology.Min/Max.
        /// Only Model[0] is used: Octree and KDTree index Model[0] themselves ("Octree - alt.cs":123,
        /// KDTree.cs:99) and Voxel_Grid's multi-topology bounds are inconsistent (Voxel_Grid.cs:67-72).</summary>
        public static IntPtr Flatten(Topology[] Model)
        {
            if (Model == null || Model.Length != 1) throw new NotSupportedException("hare_b200 flattens a single Topology (Model[0])");
            Topology T = Model[0];
            int P = T.Polygon_Count;
            double[] verts = new double[12 * P], normals = new double[3 * P], minmax = new double[6];
            int[] vcount = new int[P];
            for (int i = 0; i < P; i++)
            {
                Polygon poly = T.Polys[i];
                int n = poly.VertextCT;
                vcount[i] = n;
                for (int k = 0; k < 4; k++)
                {
                    Point p = poly.Points[Math.Min(k, n - 1)];
                    verts[12 * i + 3 * k] = p.x; verts[12 * i + 3 * k + 1] = p.y; verts[12 * i + 3 * k + 2] = p.z;
                }
                normals[3 * i] = poly.Normal.dx; normals[3 * i + 1] = poly.Normal.dy; normals[3 * i + 2] = poly.Normal.dz;
            }
            minmax[0] = T.Min.x; minmax[1] = T.Min.y; minmax[2] = T.Min.z; minmax[3] = T.Max.x; minmax[4] = T.Max.y; minmax[5] = T.Max.z;
            IntPtr h;
            Check(hare_topology_create(verts, normals, vcount, P, minmax, out h), "hare_topology_create");
            return h;
        }
    }

    /// <summary>A managed array that is pinned for the garbage collector AND page-locked for CUDA (hare_host_register), reused
    /// across batches; grows geometrically.</summary>
    internal sealed class PinnedArray<T> : IDisposable where T : struct
    {
        public T[] Data = new T[0];
        GCHandle handle; bool registered;
        public IntPtr Ptr { get { return Data.Length == 0 ? IntPtr.Zero : handle.AddrOfPinnedObject(); } }
        public void Ensure(int n)
        {
            if (n <= Data.Length) return;
            Release();
            Data = new T[Math.Max(n, 2 * Data.Length)];
            handle = GCHandle.Alloc(Data, GCHandleType.Pinned);
            registered = HareB200.hare_host_register(handle.AddrOfPinnedObject(), (UIntPtr)((ulong)Data.Length * (ulong)Marshal.SizeOf(typeof(T)))) == 0;
        }
        void Release()
        {
            if (Data.Length == 0) return;
            if (registered) HareB200.hare_host_unregister(handle.AddrOfPinnedObject());
            handle.Free(); registered = false; Data = new T[0];
        }
        public void Dispose() { Release(); }
    }
}

namespace Hare.Geometry
{
    using Hare.Geometry.Native;

    /// <summary>Common part of the three GPU partitions: handles, both reference Shoot overloads (routed through a
    /// one-element batch) and the batched overload.</summary>
    public abstract class Gpu_Spatial_Partition : Spatial_Partition, IDisposable
    {
        protected IntPtr topo = IntPtr.Zero, part = IntPtr.Zero;
        protected abstract bool HasUV { get; }
        readonly object staging_lock = new object();
        readonly PinnedArray<double> s_o = new PinnedArray<double>(), s_d = new PinnedArray<double>(), s_t = new PinnedArray<double>(),
                                     s_xyz = new PinnedArray<double>(), s_uv = new PinnedArray<double>(), s_om = new PinnedArray<double>();
        readonly PinnedArray<int> s_rid = new PinnedArray<int>(), s_pid = new PinnedArray<int>(), s_o1 = new PinnedArray<int>(), s_o2 = new PinnedArray<int>();

        public override bool Shoot(Ray R, int top_index, out X_Event Ret_event)
        {
            return Shoot(R, top_index, out Ret_event, -1, -1);
        }

        /// <summary>One ray = a one-element batch: ~100 us per call (see the header of this file).</summary>
        public override bool Shoot(Ray R, int top_index, out X_Event Ret_event, int poly_origin1, int poly_origin2 = -1)
        {
            X_Event[] ev;
            bool[] hit = Shoot(new Ray[] { R }, top_index, out ev, new int[] { poly_origin1 }, new int[] { poly_origin2 });
            Ret_event = ev[0];
            return hit[0];
        }

        /// <summary>The new batched overload (SURVEY.md 8(b)).  Rays are mutated exactly where the reference mutates
        /// them (Voxel_Grid moves a ray that starts outside the grid: AABB_Main.cs:255-257).</summary>
        public bool[] Shoot(Ray[] R, int top_index, out X_Event[] events, int[] poly_origin1 = null, int[] poly_origin2 = null)
        {
            if (top_index != 0) throw new IndexOutOfRangeException("top_index");
            int N = R.Length;
            bool[] hit = new bool[N];
            events = new X_Event[N];
            lock (staging_lock)     // the page-locked staging arrays are shared by the callers of this partition
            {
                s_o.Ensure(3 * N); s_d.Ensure(3 * N); s_t.Ensure(N); s_xyz.Ensure(3 * N); s_uv.Ensure(2 * N); s_om.Ensure(3 * N);
                s_rid.Ensure(N); s_pid.Ensure(N);
                if (poly_origin1 != null) { s_o1.Ensure(N); Array.Copy(poly_origin1, s_o1.Data, N); }
                if (poly_origin2 != null) { s_o2.Ensure(N); Array.Copy(poly_origin2, s_o2.Data, N); }
                double[] o = s_o.Data, d = s_d.Data;
                for (int i = 0; i < N; i++)
                {
                    o[3 * i] = R[i].x; o[3 * i + 1] = R[i].y; o[3 * i + 2] = R[i].z;
                    d[3 * i] = R[i].dx; d[3 * i + 1] = R[i].dy; d[3 * i + 2] = R[i].dz;
                    s_rid.Data[i] = R[i].Ray_ID;
                }
                HareB200.Check(HareB200.hare_shoot_batch(part, s_o.Ptr, s_d.Ptr, poly_origin1 != null ? s_o1.Ptr : IntPtr.Zero,
                                                         poly_origin2 != null ? s_o2.Ptr : IntPtr.Zero, s_rid.Ptr, N,
                                                         s_t.Ptr, s_xyz.Ptr, s_pid.Ptr, s_uv.Ptr, s_om.Ptr, null), "hare_shoot_batch");
                double[] t = s_t.Data, xyz = s_xyz.Data, uv = s_uv.Data, om = s_om.Data; int[] pid = s_pid.Data;
                for (int i = 0; i < N; i++)
                {
                    R[i].x = om[3 * i]; R[i].y = om[3 * i + 1]; R[i].z = om[3 * i + 2];
                    if (pid[i] == HareB200.HARE_RAY_FAULT) throw new IndexOutOfRangeException("ray " + i + ": entry voxel outside the grid (Voxel_Grid.cs:374-383)");
                    hit[i] = pid[i] >= 0;
                    events[i] = hit[i]
                        ? new X_Event(new Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), uv[2 * i], uv[2 * i + 1], t[i], pid[i])
                        : new X_Event();
                }
            }
            return hit;
        }

        /// <summary>Write the flattened partition to disk / read it back for the same Topology (hare_part_save / hare_part_load).</summary>
        public void Save(string path) { HareB200.Check(HareB200.hare_part_save(part, path), "hare_part_save"); }

        public void Dispose()
        {
            if (part != IntPtr.Zero) { HareB200.hare_part_destroy(part); part = IntPtr.Zero; }       // partitions first: they read the Topology's records
            if (topo != IntPtr.Zero) { HareB200.hare_topology_destroy(topo); topo = IntPtr.Zero; }
            s_o.Dispose(); s_d.Dispose(); s_t.Dispose(); s_xyz.Dispose(); s_uv.Dispose(); s_om.Dispose(); s_rid.Dispose(); s_pid.Dispose(); s_o1.Dispose(); s_o2.Dispose();
            GC.SuppressFinalize(this);
        }
        ~Gpu_Spatial_Partition() { Dispose(); }
    }

    /// <summary>Voxel_Grid(Topology[] Model_in, int Domain)  (Voxel_Grid.cs:48) and Voxel_Grid(Topology[], int MaxDomain, int Avg_polys)
    /// (:128): cell lists are built on the GPU; the public surface the reference's callers can see is kept.</summary>
#if HARE_B200
    public class Voxel_Grid : Gpu_Spatial_Partition
#else
    public class Gpu_Voxel_Grid : Gpu_Spatial_Partition
#endif
    {
        protected override bool HasUV { get { return false; } }
        protected int VoxelCtX, VoxelCtY, VoxelCtZ, XYTot;
        protected Point BoxDims, VoxelDims, OBoxMin, OBoxMax;
        protected double Epsilon = 0.001;

        /// <summary>The reference's per-thread mailboxes (Voxel_Grid.cs:29, 54-62).  Mailboxing never changes a result (strict
        /// t &lt; tmin), the kernels keep none; the field stays for source compatibility, with empty per-box arrays.</summary>
        public int[,][] Poly_Ray_ID;

        List<int>[,,,] voxel_inv;
        /// <summary>Voxel_Inv[x, y, z, m]: the polygon list of a voxel (Voxel_Grid.cs:33), materialised from the device lists
        /// (hare_voxelgrid_download) on first use.</summary>
        public List<int>[,,,] Voxel_Inv
        {
            get
            {
                if (voxel_inv == null)
                {
                    long ncells = (long)VoxelCtX * VoxelCtY * VoxelCtZ;
                    uint[] off = new uint[ncells + 1], pol = new uint[Math.Max(1, npairs)];
                    HareB200.Check(HareB200.hare_voxelgrid_download(part, off, pol), "hare_voxelgrid_download");
                    List<int>[,,,] v = new List<int>[VoxelCtX, VoxelCtY, VoxelCtZ, 1];
                    for (int x = 0; x < VoxelCtX; x++) for (int y = 0; y < VoxelCtY; y++) for (int z = 0; z < VoxelCtZ; z++)
                    {
                        long c = ((long)x * VoxelCtY + y) * VoxelCtZ + z;          // cell index of include/hare_b200.h
                        List<int> l = new List<int>((int)(off[c + 1] - off[c]));
                        for (uint k = off[c]; k < off[c + 1]; k++) l.Add((int)pol[k]);
                        v[x, y, z, 0] = l;
                    }
                    voxel_inv = v;
                }
                return voxel_inv;
            }
        }
        long npairs;

#if HARE_B200
        public Voxel_Grid(Topology[] Model_in, int Domain)
#else
        public Gpu_Voxel_Grid(Topology[] Model_in, int Domain)
#endif
        {
            Model = Model_in;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_voxelgrid_build(topo, Domain, out part), "hare_voxelgrid_build");
            After_Build();
        }

#if HARE_B200
        public Voxel_Grid(Topology[] Model_in, int MaxDomain, int Avg_polys)
#else
        public Gpu_Voxel_Grid(Topology[] Model_in, int MaxDomain, int Avg_polys)
#endif
        {
            Model = Model_in;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_voxelgrid_build_adaptive(topo, MaxDomain, Avg_polys, out part), "hare_voxelgrid_build_adaptive");
            After_Build();
        }

        void After_Build()
        {
            double[] obox = new double[6], vd = new double[3]; int[] ct = new int[3];
            HareB200.Check(HareB200.hare_voxelgrid_info(part, obox, vd, ct, out npairs), "hare_voxelgrid_info");
            VoxelCtX = ct[0]; VoxelCtY = ct[1]; VoxelCtZ = ct[2]; XYTot = VoxelCtX * VoxelCtY;
            OBoxMin = new Point(obox[0], obox[1], obox[2]); OBoxMax = new Point(obox[3], obox[4], obox[5]);
            BoxDims = new Point(obox[3] - obox[0], obox[4] - obox[1], obox[5] - obox[2]);                           // Voxel_Grid.cs:76
            VoxelDims = new Point(vd[0], vd[1], vd[2]);
            Char_Step = (vd[0] < vd[1]) ? ((vd[0] < vd[2]) ? vd[0] : vd[2]) : (vd[1] < vd[2] ? vd[1] : vd[2]);         // :90
            Poly_Ray_ID = new int[Model.Length, 0][];
        }

        // Voxel_Grid.cs:256-267, kept literally -- including that VoxelDecode returns X and Y swapped with respect to VoxelCode
        public void VoxelDecode(int Code, out int X, out int Y, out int Z)
        {
            Z = (int)Math.Floor((double)(Code / XYTot));
            Code -= Z * XYTot;
            Y = (int)Math.Floor((double)(Code / VoxelCtY));
            X = Code - Y * VoxelCtY;
        }

        public int VoxelCode(int X, int Y, int Z) { return XYTot * Z + VoxelCtY * X + Y; }

        /// <summary>Thread body of the reference's CPU build (Voxel_Grid.cs:273-304).  The lists are built on the GPU by the
        /// constructor; there is nothing left to fill.</summary>
        public void Fill_Voxels(object o) { throw new NotSupportedException("the voxel lists are built on the GPU by the constructor (hare_voxelgrid_build)"); }

        public void PointInVoxel(Point Pt, out int X, out int Y, out int Z)      // :322-327
        {
            X = (int)Math.Floor((Pt.x - OBoxMin.x) / VoxelDims.x);
            Y = (int)Math.Floor((Pt.y - OBoxMin.y) / VoxelDims.y);
            Z = (int)Math.Floor((Pt.z - OBoxMin.z) / VoxelDims.z);
        }

        public int PointInVoxel(Point Pt)                                         // :329-332
        {
            return VoxelCode((int)Math.Floor((Pt.x - OBoxMin.x) / VoxelDims.x), (int)Math.Floor((Pt.y - OBoxMin.y) / VoxelDims.y), (int)Math.Floor((Pt.z - OBoxMin.z) / VoxelDims.z));
        }

        public double Xdim { get { return BoxDims.x; } }                          // :763-784
        public double Ydim { get { return BoxDims.y; } }
        public double Zdim { get { return BoxDims.z; } }
        public Point MinPt { get { return OBoxMin; } }                            // :786-791
    }

    /// <summary>Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)  ("Octree - alt.cs":45).</summary>
#if HARE_B200
    public class Octree : Gpu_Spatial_Partition
#else
    public class Gpu_Octree : Gpu_Spatial_Partition
#endif
    {
        protected override bool HasUV { get { return true; } }
#if HARE_B200
        public Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
#else
        public Gpu_Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
#endif
        {
            Model = Model_In;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_octree_build(topo, maxDepth, maxPolygonsPerNode, out part), "hare_octree_build");
        }
    }

    /// <summary>KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)  (KDTree.cs:51).</summary>
#if HARE_B200
    public class KDTree : Gpu_Spatial_Partition
#else
    public class Gpu_KDTree : Gpu_Spatial_Partition
#endif
    {
        protected override bool HasUV { get { return true; } }
#if HARE_B200
        public KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
#else
        public Gpu_KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
#endif
        {
            Model = Model_In;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_kdtree_build(topo, maxDepth, maxPolygonsPerNode, out part), "hare_kdtree_build");
        }
    }
}
