// HareOracle.cs -- harness that runs the REAL reference (PachydermAcoustic/Hare, unmodified) on inputs
// dumped by tools/csharp_golden.py and writes its X_Events as raw binary, so that "restatement-pinned"
// golden vectors can be upgraded to "reference-executed" ones on any machine with a .NET 7 SDK:
//
//     python tools/csharp_golden.py dump case.bin            # mesh + rays + partition arguments
//     dotnet new console -o HareOracle && cp oracle/csharp/HareOracle.cs HareOracle/Program.cs
//     cd HareOracle && dotnet add reference <path to>/Hare_NC.csproj && dotnet run -- ../case.bin ../case.out
//     python tools/csharp_golden.py compare case.bin case.out  # against the C++ oracle and, on a GPU box, the CUDA path
//
// NOT COMPILED IN THIS REPOSITORY'S IMAGE: no .NET toolchain exists here (SURVEY.md 8(c)).
//
// case.bin  : int32 P, int32 N, int32 kind (1 Voxel_Grid, 2 Octree, 3 KDTree), int32 arg0, int32 arg1,
//             double min[3], double max[3], then P x (int32 n, n x 3 doubles), then N x 6 doubles (o, d),
//             then N x 2 int32 (poly_origin1, poly_origin2)
// case.out  : N x (int32 Poly_id, double t, double x, y, z, u, v, double ox, oy, oz)   (ray origin after the call)
using System;
using System.IO;
using Hare.Geometry;

class HareOracle
{
    static void Main(string[] a)
    {
        using var br = new BinaryReader(File.OpenRead(a[0]));
        int P = br.ReadInt32(), N = br.ReadInt32(), kind = br.ReadInt32(), arg0 = br.ReadInt32(), arg1 = br.ReadInt32();
        Point mn = new Point(br.ReadDouble(), br.ReadDouble(), br.ReadDouble());
        Point mx = new Point(br.ReadDouble(), br.ReadDouble(), br.ReadDouble());
        Topology T = new Topology(mn, mx);
        for (int i = 0; i < P; i++)
        {
            int n = br.ReadInt32();
            Point[] poly = new Point[n];
            for (int k = 0; k < n; k++) poly[k] = new Point(br.ReadDouble(), br.ReadDouble(), br.ReadDouble());
            T.Add_Polygon(poly);
        }
        T.Finish_Topology();
        Topology[] Model = new Topology[] { T };
        Spatial_Partition part = kind == 1 ? new Voxel_Grid(Model, arg0)
                               : kind == 2 ? (Spatial_Partition)new Octree(Model, arg0, arg1)
                               : new KDTree(Model, arg0, arg1);
        Ray[] R = new Ray[N];
        for (int i = 0; i < N; i++)
            R[i] = new Ray(br.ReadDouble(), br.ReadDouble(), br.ReadDouble(), br.ReadDouble(), br.ReadDouble(), br.ReadDouble(), 0, i + 1);
        int[] o1 = new int[N], o2 = new int[N];
        for (int i = 0; i < N; i++) { o1[i] = br.ReadInt32(); o2[i] = br.ReadInt32(); }
        using var bw = new BinaryWriter(File.Create(a[1]));
        for (int i = 0; i < N; i++)      // sequential on purpose: one mailbox slot at a time (Voxel_Grid.cs:334-342)
        {
            X_Event ev; int pid; bool hit;
            try { hit = part.Shoot(R[i], 0, out ev, o1[i], o2[i]); pid = hit ? ev.Poly_id : -1; }
            catch (IndexOutOfRangeException) { hit = false; ev = new X_Event(); pid = -2; }
            bw.Write(pid); bw.Write(hit ? ev.t : 0.0);
            bw.Write(hit ? ev.X_Point.x : 0.0); bw.Write(hit ? ev.X_Point.y : 0.0); bw.Write(hit ? ev.X_Point.z : 0.0);
            bw.Write(hit ? ev.u : 0.0); bw.Write(hit ? ev.v : 0.0);
            bw.Write(R[i].x); bw.Write(R[i].y); bw.Write(R[i].z);
        }
    }
}
