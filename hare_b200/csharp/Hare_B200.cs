// Hare_B200.cs -- binding of libhare_b200 (include/hare_b200.h) for the Hare_NC (.NET) build.
//
// Drop this file next to Hare's sources (both csproj files glob *.cs) and ship libhare_b200.so
// beside the assembly.  It adds, without touching any existing member:
//   * Hare.Geometry.Native.HareB200          -- the [DllImport] block, one entry per C-ABI function;
//   * Hare.Geometry.Gpu_Voxel_Grid / Gpu_Octree / Gpu_KDTree -- Spatial_Partition subclasses with
//     the reference constructors' signatures, both reference Shoot overloads, and the new
//     batched overload  bool[] Shoot(Ray[] R, int top_index, out X_Event[] events,
//                                    int[] poly_origin1 = null, int[] poly_origin2 = null).
// Pachyderm code that says `new Voxel_Grid(Model, Domain)` switches by saying
// `new Gpu_Voxel_Grid(Model, Domain)`; everything it touches afterwards (Model, Char_Step, Shoot)
// is the abstract Spatial_Partition surface (Spatial_Partition.cs:27-35).
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no .NET toolchain); kept deliberately thin: it only
// flattens and marshals.  All arithmetic that decides a result lives behind the C ABI.
using System;
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

        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_shoot_batch(IntPtr part, double[] o, double[] d, int[] origin1, int[] origin2, int[] ray_id, long N,
                                                  double[] t, double[] xyz, int[] poly_id, double[] uv, double[] o_moved, ulong[] counters);
        [DllImport(Lib, CallingConvention = CC)]
        public static extern int hare_reflect_chain(IntPtr part, double[] o, double[] d, long N, int order, int[] ev_poly_id, double[] ev_t,
                                                    double[] fin_o, double[] fin_d, int[] nshots, out ulong total_shots, ulong[] counters);

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

        public override bool Shoot(Ray R, int top_index, out X_Event Ret_event)
        {
            return Shoot(R, top_index, out Ret_event, -1, -1);
        }

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
            double[] o = new double[3 * N], d = new double[3 * N], t = new double[N], xyz = new double[3 * N], uv = new double[2 * N], om = new double[3 * N];
            int[] rid = new int[N], pid = new int[N];
            for (int i = 0; i < N; i++)
            {
                o[3 * i] = R[i].x; o[3 * i + 1] = R[i].y; o[3 * i + 2] = R[i].z;
                d[3 * i] = R[i].dx; d[3 * i + 1] = R[i].dy; d[3 * i + 2] = R[i].dz;
                rid[i] = R[i].Ray_ID;
            }
            HareB200.Check(HareB200.hare_shoot_batch(part, o, d, poly_origin1, poly_origin2, rid, N, t, xyz, pid, uv, om, null), "hare_shoot_batch");
            bool[] hit = new bool[N];
            events = new X_Event[N];
            for (int i = 0; i < N; i++)
            {
                R[i].x = om[3 * i]; R[i].y = om[3 * i + 1]; R[i].z = om[3 * i + 2];
                if (pid[i] == HareB200.HARE_RAY_FAULT) throw new IndexOutOfRangeException("ray " + i + ": entry voxel outside the grid (Voxel_Grid.cs:374-383)");
                hit[i] = pid[i] >= 0;
                events[i] = hit[i]
                    ? new X_Event(new Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), uv[2 * i], uv[2 * i + 1], t[i], pid[i])
                    : new X_Event();
            }
            return hit;
        }

        public void Dispose()
        {
            if (part != IntPtr.Zero) { HareB200.hare_part_destroy(part); part = IntPtr.Zero; }
            if (topo != IntPtr.Zero) { HareB200.hare_topology_destroy(topo); topo = IntPtr.Zero; }
            GC.SuppressFinalize(this);
        }
        ~Gpu_Spatial_Partition() { Dispose(); }
    }

    /// <summary>Voxel_Grid(Topology[] Model_in, int Domain)  (Voxel_Grid.cs:48): cell lists are built on the GPU.</summary>
    public class Gpu_Voxel_Grid : Gpu_Spatial_Partition
    {
        protected override bool HasUV { get { return false; } }
        public Gpu_Voxel_Grid(Topology[] Model_in, int Domain)
        {
            Model = Model_in;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_voxelgrid_build(topo, Domain, out part), "hare_voxelgrid_build");
            Set_Char_Step();
        }

        /// <summary>Voxel_Grid(Topology[] Model_in, int MaxDomain, int Avg_polys)  (Voxel_Grid.cs:128).</summary>
        public Gpu_Voxel_Grid(Topology[] Model_in, int MaxDomain, int Avg_polys)
        {
            Model = Model_in;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_voxelgrid_build_adaptive(topo, MaxDomain, Avg_polys, out part), "hare_voxelgrid_build_adaptive");
            Set_Char_Step();
        }

        void Set_Char_Step()
        {
            double[] obox = new double[6], vd = new double[3]; int[] ct = new int[3]; long n;
            HareB200.Check(HareB200.hare_voxelgrid_info(part, obox, vd, ct, out n), "hare_voxelgrid_info");
            Char_Step = (vd[0] < vd[1]) ? ((vd[0] < vd[2]) ? vd[0] : vd[2]) : (vd[1] < vd[2] ? vd[1] : vd[2]);   // Voxel_Grid.cs:90
        }
    }

    /// <summary>Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)  ("Octree - alt.cs":45).</summary>
    public class Gpu_Octree : Gpu_Spatial_Partition
    {
        protected override bool HasUV { get { return true; } }
        public Gpu_Octree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
        {
            Model = Model_In;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_octree_build(topo, maxDepth, maxPolygonsPerNode, out part), "hare_octree_build");
        }
    }

    /// <summary>KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)  (KDTree.cs:51).</summary>
    public class Gpu_KDTree : Gpu_Spatial_Partition
    {
        protected override bool HasUV { get { return true; } }
        public Gpu_KDTree(Topology[] Model_In, int maxDepth, int maxPolygonsPerNode)
        {
            Model = Model_In;
            topo = HareB200.Flatten(Model);
            HareB200.Check(HareB200.hare_kdtree_build(topo, maxDepth, maxPolygonsPerNode, out part), "hare_kdtree_build");
        }
    }
}
