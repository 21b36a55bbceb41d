"""CPU-side checks of the product library (no GPU): the C ABI loads and exports every symbol
include/hare_b200.h declares, the host-side construction steps (Topology ingest, Octree and
KDTree builders) agree with the oracle, arguments are validated, and every compute entry point
fails loudly when no device is available (there is no CPU fallback)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import hare_b200 as hb
from hare_b200 import _lib
from hare_b200.harness import meshes
from hare_b200.harness.meshes import Mesh
from oracle import hare_oracle as ho
from tests.util import canon_kdtree, canon_octree

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hare_golden.npz"))
NO_GPU = hb.lib().hare_device_count() == 0


@pytest.fixture(scope="module")
def host_only():
    """hare_init(NULL, -1): handles without a device replica (build-time tooling mode)."""
    hb.init(host_only=True)
    yield hb
    if not NO_GPU:
        hb.init([0])


def test_abi_exports_every_declared_symbol():
    syms = _lib.declared_symbols()
    assert len(syms) >= 28 and "hare_shoot_batch" in syms and "hare_voxelgrid_build" in syms
    L = hb.lib()
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(syms) <= exported
    assert not [s for s in exported if s.startswith("ho_")], "the oracle must not be linked into the product"
    assert b"sm_100a" in L.hare_version()


def test_shared_object_contains_sm100a_code_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.parametrize("level", ["tiny", "2k", "10k"])
def test_ingest_matches_oracle(host_only, level):
    mesh = meshes.hall(level)
    T = hb.Topology.from_mesh(mesh); To = ho.Topology.from_mesh(mesh)
    v, n, c, mm = To.arrays()
    assert np.array_equal(T.verts, v) and np.array_equal(T.normals, n) and np.array_equal(T._mm, mm)
    assert T.Vertex_Count == To.Vertex_Count and T.Polygon_Count == mesh.P


def test_ingest_matches_golden_python_restatement(host_only):
    for name in ("shoebox", "tiny"):
        mesh = Mesh(G[f"{name}_verts"], G[f"{name}_vcount"], G[f"{name}_minpt"], G[f"{name}_maxpt"], name)
        T = hb.Topology.from_mesh(mesh)
        assert np.array_equal(T.verts, G[f"{name}_topo_verts"]) and np.array_equal(T.normals, G[f"{name}_topo_normals"])
        assert np.array_equal(T._mm, G[f"{name}_topo_minmax"])


def test_ingest_rounds_and_welds(host_only):
    """Math.Round(x, 15) and the 1 mm weld (Hare_Geometry_Topology.cs:342-377)."""
    T = hb.Topology([0, 0, 0], [2, 2, 2])
    a = 0.1234567890123456789
    T.Add_Polygon([[a, 0, 0], [1, 0, 0], [1, 1, 0]])
    T.Add_Polygon([[a + 2e-4, 1e-4, 0], [1, 1, 0], [0, 1, 1]])        # first vertex shares the 1 mm cell of polygon 0's
    T.Finish_Topology()
    assert T.verts[0, 0, 0] == ho.round15(a) == 0.123456789012346
    assert np.array_equal(T.verts[1, 0], T.verts[0, 0])                # welded onto the first vertex seen
    assert T.Vertex_Count == 4
    with pytest.raises(NotImplementedError):
        T.Add_Polygon(np.zeros((5, 3)))


@pytest.mark.parametrize("level,args", [("tiny", (3, 4)), ("2k", (5, 8)), ("10k", (6, 16))])
def test_octree_build_matches_oracle(host_only, level, args):
    mesh = meshes.hall(level)
    t = hb.Octree([hb.Topology.from_mesh(mesh)], *args)
    o = ho.Octree(ho.Topology.from_mesh(mesh), *args)
    assert canon_octree(*t.arrays()) == canon_octree(*o.arrays())
    i = t.info()
    assert (i["nodes"], i["list_entries"], i["lost"]) == o.info() and i["depth"] <= args[0]


@pytest.mark.parametrize("level,args", [("tiny", (8, 4)), ("2k", (14, 8)), ("10k", (18, 16))])
def test_kdtree_build_matches_oracle(host_only, level, args):
    mesh = meshes.hall(level)
    t = hb.KDTree([hb.Topology.from_mesh(mesh)], *args)
    o = ho.KDTree(ho.Topology.from_mesh(mesh), *args)
    box, sp, ax, le, lo, lc, pol = t.arrays()
    assert canon_kdtree(box, sp, ax, le, le + 1, lo, lc, pol) == canon_kdtree(*o.arrays())


def test_upload_validation(host_only):
    T = hb.Topology.from_mesh(meshes.shoebox())
    with pytest.raises(hb.HareError):     # polygon index out of range
        hb.Voxel_Grid.from_lists([T], [0, 0, 0, 1, 1, 1], [1, 1, 1], [0, 1], [99])
    with pytest.raises(hb.HareError):     # child index out of range
        hb.Octree.from_nodes([T], np.zeros((1, 6)), [5], [0], [0], [0])
    with pytest.raises(hb.HareError):     # bad axis
        hb.KDTree.from_nodes([T], np.zeros((3, 6)), [0, 0, 0], [7, -1, -1], [1, -1, -1], [0, 0, 0], [0, 0, 0], [0])
    g = hb.Voxel_Grid.from_lists([T], [0, 0, 0, 2, 2, 2], [2, 2, 2], np.arange(9), np.arange(8) % 6)
    obox, vd, ct, n = g.info()
    assert vd.tolist() == [1.0, 1.0, 1.0] and n == 8 and g.Char_Step == 1.0


def test_upload_rejects_malformed_trees_and_offsets(host_only):
    """a cycle (first_child[0] = 0) used to hang the depth scan, a child index <= its parent silently broke the packing
    passes, and decreasing cell offsets made the kernels read out of bounds.  All are refused now -- also through hare_part_load."""
    T = hb.Topology.from_mesh(meshes.shoebox())
    z6 = lambda n: np.zeros((n, 6))
    with pytest.raises(hb.HareError, match="non-decreasing"):
        hb.Voxel_Grid.from_lists([T], [0, 0, 0, 2, 2, 2], [2, 2, 2], [0, 3, 2, 4, 5, 5, 6, 7, 8], np.arange(8) % 6)
    with pytest.raises(hb.HareError, match=r"cell_offset\[0\]"):
        hb.Voxel_Grid.from_lists([T], [0, 0, 0, 1, 1, 1], [1, 1, 1], [1, 2], [0, 1])
    with pytest.raises(hb.HareError, match="follow their parent"):          # the root is its own child: a cycle
        hb.Octree.from_nodes([T], z6(9), [0] + [-1] * 8, [0] * 9, [0] * 9, [0])
    with pytest.raises(hb.HareError, match="follow their parent"):          # child block in front of its parent
        hb.Octree.from_nodes([T], z6(17), [9, -1, -1, -1, -1, -1, -1, -1, -1, 1] + [-1] * 7, [0] * 17, [0] * 17, [0])
    with pytest.raises(hb.HareError, match="two parents"):                  # nodes 1 and 2 share the block 9..16
        hb.Octree.from_nodes([T], z6(17), [1, 9, 9] + [-1] * 14, [0] * 17, [0] * 17, [0])
    with pytest.raises(hb.HareError, match="not reachable"):                # 8 orphan nodes behind the root's children
        hb.Octree.from_nodes([T], z6(17), [1] + [-1] * 16, [0] * 17, [0] * 17, [0])
    with pytest.raises(hb.HareError, match="bad internal node"):            # kd cycle: left[0] = 0
        hb.KDTree.from_nodes([T], z6(3), [0.0] * 3, [0, -1, -1], [0, -1, -1], [0] * 3, [0] * 3, [0])
    with pytest.raises(hb.HareError, match="two parents"):
        hb.KDTree.from_nodes([T], z6(5), [0.0] * 5, [0, 1, 1, -1, -1], [1, 3, 3, -1, -1], [0] * 5, [0] * 5, [0])
    with pytest.raises(hb.HareError, match="not reachable"):
        hb.KDTree.from_nodes([T], z6(5), [0.0] * 5, [0, -1, -1, -1, -1], [1, -1, -1, -1, -1], [0] * 5, [0] * 5, [0])
    # a chain deeper than the kernels' frame budget
    n = 8 * 21 + 1
    fc = np.full(n, -1, np.int32); fc[0] = 1
    for k in range(1, 21):
        fc[8 * (k - 1) + 1] = 8 * k + 1
    with pytest.raises(NotImplementedError, match="deeper"):
        hb.Octree.from_nodes([T], z6(n), fc, np.zeros(n), np.zeros(n), [0])


def test_topology_outlives_its_partitions(host_only):
    """a partition reads its Topology's records; the library refuses to free a Topology with live partitions, and the
    Python mirror refuses to re-finish one."""
    T = hb.Topology.from_mesh(meshes.shoebox())
    t = hb.Octree([T], 3, 2)
    assert hb.lib().hare_topology_destroy(T._h) != 0 and b"still alive" in hb.lib().hare_last_error()
    with pytest.raises(hb.HareError, match="still alive"):
        T.Finish_Topology()
    assert t.info()["nodes"] > 1            # the partition is intact
    del t
    T.Finish_Topology()                     # fine once the partition is gone


def test_no_cpu_fallback(host_only):
    """Host-only handles cannot compute: Shoot, chains and the GPU grid build fail with HARE_ERR_CUDA."""
    T = hb.Topology.from_mesh(meshes.shoebox())
    with pytest.raises(hb.HareError, match="no CPU fallback"):
        hb.Voxel_Grid([T], 10)
    t = hb.Octree([T], 3, 2)
    with pytest.raises(hb.HareError, match="no CPU fallback"):
        t.Shoot_Batch(np.zeros((4, 3)), np.ones((4, 3)))
    with pytest.raises(hb.HareError, match="no CPU fallback"):
        t.Reflect_Chain(np.zeros((4, 3)), np.ones((4, 3)), 3)


@pytest.mark.skipif(not NO_GPU, reason="only meaningful on a box without a CUDA device")
def test_device_init_fails_loudly_without_gpu():
    with pytest.raises(hb.HareError, match="no CUDA device"):
        hb.init([0])


def test_missing_extension_raises(tmp_path):
    code = ("import os, sys; os.environ['HARE_B200_LIB'] = %r; sys.path.insert(0, %r)\n"
            "import hare_b200\n"
            "try:\n    hare_b200.lib()\nexcept hare_b200.HareError as e:\n    print('RAISED', e)\n") % (str(tmp_path / "nope.so"), _lib.ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "RAISED" in out and "no CPU fallback" in out


@pytest.mark.parametrize("kind,args", [("Octree", (5, 8)), ("KDTree", (12, 8))])
def test_partition_save_load_roundtrip(host_only, tmp_path, kind, args):
    """On-disk form of a flattened partition: load == what was saved, for the same Topology only."""
    mesh = meshes.hall("2k")
    T = hb.Topology.from_mesh(mesh)
    part = getattr(hb, kind)([T], *args)
    f = str(tmp_path / "part.hare")
    part.Save(f)
    back = getattr(hb, kind).Load([T], f)
    for a, b in zip(part.arrays(), back.arrays()):
        assert np.array_equal(a, b)
    assert part.info() == back.info()
    other = {"Octree": hb.KDTree, "KDTree": hb.Octree}[kind]
    with pytest.raises(ValueError):
        other.Load([T], f)
    T2 = hb.Topology.from_mesh(meshes.hall("tiny"))
    with pytest.raises(hb.HareError):
        getattr(hb, kind).Load([T2], f)
    with open(f, "r+b") as fh:                      # a truncated file is refused, not read past its end
        fh.truncate(200)
    with pytest.raises(hb.HareError):
        getattr(hb, kind).Load([T], f)


def test_sass_has_no_contracted_fp64_multiply_adds():
    """Bit-exactness against the reference's C# doubles needs every a*b+c to stay two roundings (nvcc -fmad=false).  Guard on the
    shipped SASS: the kernels that are pure mul/add chains (the SAT of the grid and Octree builds) contain NO DFMA at all although
    they hold > 100 DMUL and DADD each; in every other kernel DFMA only occurs in the IEEE division / reciprocal sequences
    (MUFU.RCP64H + Newton steps, <= 9 DFMA per division) and in the handful of explicit fma() calls of the conservative culls."""
    import re
    out = subprocess.run(["cuobjdump", "-sass", _lib.SO_PATH], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", out)[1:]
    assert len(parts) >= 30
    seen = set()
    for p in parts:
        name = p.split("\n", 1)[0].strip()
        dfma = len(re.findall(r"\bDFMA\b", p)); rcp = len(re.findall(r"MUFU\.RCP64H", p))
        dmul = len(re.findall(r"\bDMUL\b", p)); dadd = len(re.findall(r"\bDADD\b", p))
        if "vg_refine_kernelILi0" in name or "oct_mask_kernel" in name:
            seen.add(name[:40])
            assert dfma == 0 and dmul > 100 and dadd > 100, (name, dfma, dmul, dadd)
        assert dfma <= 9 * rcp + 8, (name, dfma, rcp)
    assert len(seen) == 2
    for k in ("vg_wave_kernel", "oct_wave_kernel", "kd_wave_kernel"):
        assert any(k in p.split("\n", 1)[0] for p in parts), k


def test_ray_order_and_sampling_helpers():
    """harness/rays.py: the direction of a ray depends on its number only; 'source-major' hands every source one contiguous run of the
    workload and a shard of it equals the same rays of the whole; sample_blocks / sample_rays pick evenly spaced runs that reach every
    source and can be generated without the rest of the batch."""
    from hare_b200.harness import rays_from_sources, sample_blocks, sample_rays, source_index
    S = meshes.sources(8)
    n = 100_003
    oi, di = rays_from_sources(n, S, stream=3)
    om, dm = rays_from_sources(n, S, stream=3, order="source-major", total=n)
    assert np.array_equal(di, dm)
    src = source_index(np.arange(n), 8, "source-major", n)
    assert np.all(np.diff(src) >= 0) and set(src) == set(range(8)) and np.array_equal(om, S[src])
    assert np.bincount(src).max() - np.bincount(src).min() <= 1
    assert np.array_equal(oi, S[np.arange(n) % 8])
    o2, d2 = rays_from_sources(30_001, S, stream=3, first=45_000, order="source-major", total=n)      # a rank's block
    assert np.array_equal(o2, om[45_000:75_001]) and np.array_equal(d2, dm[45_000:75_001])
    with pytest.raises(ValueError):
        rays_from_sources(10, S, order="source-major")
    idx = sample_blocks(n, 6_400)
    assert len(idx) == 6_400 and np.all(np.diff(idx) > 0) and idx[-1] < n and set(src[idx]) == set(range(8))
    assert np.array_equal(sample_blocks(50, 100), np.arange(50)) and np.array_equal(sample_blocks(1000, 10), np.arange(10))
    i3, o3, d3 = sample_rays(n, 6_400, S, stream=3, order="source-major")
    assert np.array_equal(i3, idx) and np.array_equal(o3, om[idx]) and np.array_equal(d3, dm[idx])


def test_q8_q9_a_single_topology_per_partition(host_only):
    """Q8 / Q9: the reference's Octree and KDTree overwrite `root` per topology and index Model[0] only ("Octree - alt.cs":63-88, 123;
    KDTree.cs:71-87, 99), and Voxel_Grid's multi-topology bounds are inconsistent (Voxel_Grid.cs:67-72): the boundary takes
    Model.Length == 1 and says so instead of guessing."""
    T = hb.Topology.from_mesh(meshes.shoebox())
    T2 = hb.Topology.from_mesh(meshes.shoebox())
    for ctor in (lambda M: hb.Octree(M, 3, 2), lambda M: hb.KDTree(M, 4, 1), lambda M: hb.Voxel_Grid(M, 10)):
        with pytest.raises(NotImplementedError, match="single Topology"):
            ctor([T, T2])
        with pytest.raises(NotImplementedError, match="single Topology"):
            ctor([])
    assert hb.Octree([T], 3, 2).info()["nodes"] > 1
