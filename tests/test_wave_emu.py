"""CPU replay of the wavefront scheduler (hare_b200/csrc/vg_wave.cuh) against the oracle.

vg_wave.cuh's per-slot phase functions and its scheduling policy are __host__ __device__; tests/emu/wave_emu.cu
compiles them for the host and drives them with a sequential copy of the kernel's trip loop.  These tests check the
state machine (slot packing, phase transitions, ray numbering, chain bookkeeping) without a GPU: every output must be
bit-identical to the oracle's Voxel_Grid.Shoot / reflection chain.  The kernel itself is covered by the -m gpu tests.
"""
import shutil

import numpy as np
import pytest

from hare_b200.harness import meshes, rays_from_sources
from oracle import hare_oracle as ho

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="the emulator is compiled with nvcc (host code only)")


@pytest.fixture(scope="module")
def hall():
    T = ho.Topology.from_mesh(meshes.hall("10k"))
    g = ho.Voxel_Grid(T, 24, "fast", nthreads=4)
    return T, g, T.arrays(), g.info(), g.csr()


@pytest.mark.parametrize("slots,wmax,warps", [(64, 4, 3), (48, 8, 1), (32, 4, 2), (96, 4, 5)])
def test_wave_shoot_matches_oracle(hall, slots, wmax, warps):
    from tests.emu import wave_emu
    T, g, ta, gi, csr = hall
    o, d = rays_from_sources(3000, meshes.sources(4), stream=2)
    o[::7] += np.array([60.0, -3.0, 2.0])          # some rays start outside the grid (origin moved, t offset)
    rid = np.arange(1, 3001, dtype=np.int32); rid[5::11] = 0   # Ray_ID == 0: blind rays
    o1 = np.full(3000, -1, np.int32)
    ref0 = g.Shoot(o, d, nthreads=4)
    o1[::3] = ref0["poly_id"][::3]                 # poly_origin1 = the polygon the plain Shoot hits
    # the reference's Ray_ID == 0 behaviour depends on what earlier rays left in the mailbox; the batched API defines the
    # fresh-mailbox case (nothing is tested -> miss), so blind rays are checked apart (as in tests/test_gpu_parity.py)
    blind = rid == 0
    ref = g.Shoot(o, d, origin1=o1, ray_id=np.where(blind, 1 << 30, rid).astype(np.int32), nthreads=4)
    got = wave_emu.run(ta, gi, csr, o, d, origin1=o1, ray_id=rid, slots=slots, wmax=wmax, n_warps=warps)
    for k in ("poly_id", "t", "xyz"):
        assert np.array_equal(got[k][~blind], ref[k][~blind]), k
    assert np.array_equal(got["o"], ref["o"])
    assert (got["poly_id"][blind] == -1).all() and not got["t"][blind].any() and not got["xyz"][blind].any()
    assert not got["uv"].any()


@pytest.mark.parametrize("slots,wmax,warps", [(64, 4, 2), (40, 4, 7)])
def test_wave_chain_matches_oracle(hall, slots, wmax, warps):
    from tests.emu import wave_emu
    T, g, ta, gi, csr = hall
    o, d = rays_from_sources(1500, meshes.sources(4), stream=3)
    ref = g.reflect_chain(o, d, 12, nthreads=4, points=True)
    got = wave_emu.run(ta, gi, csr, o, d, chain=True, order=12, slots=slots, wmax=wmax, n_warps=warps)
    for k in ("ev_poly_id", "ev_t", "ev_xyz", "ev_uv", "o", "d", "nshots"):     # ev_xyz / ev_uv: per-bounce X_Point and u, v rows
        assert np.array_equal(got[k], ref[k]), k
    assert int(got["total"][0]) == int(ref["nshots"].sum())


def test_wave_empty_and_tiny_batches(hall):
    from tests.emu import wave_emu
    T, g, ta, gi, csr = hall
    for n in (0, 1, 33):
        o, d = rays_from_sources(max(n, 1), meshes.sources(1), stream=4)
        o, d = o[:n], d[:n]
        got = wave_emu.run(ta, gi, csr, o, d, slots=64, wmax=4, n_warps=4)
        if n:
            ref = g.Shoot(o, d)
            assert np.array_equal(got["poly_id"], ref["poly_id"]) and np.array_equal(got["t"], ref["t"])


def test_wave_far_origins_and_scaled_directions(hall):
    """The box cull works in FP32 around the current voxel's exit point: rays shot from 1e5 m / 1e9 m away, and rays with very
    short or very long direction vectors, still give the oracle's events bit for bit."""
    from tests.emu import wave_emu
    T, g, ta, gi, csr = hall
    o, d = rays_from_sources(2000, meshes.sources(4), stream=6)
    for oo, dd in ((o - d * 1e5, d), (o - d * 1e9, d), (o, d * 1e-3), (o, d * 1e-30), (o, d * 1e6)):
        ref = g.Shoot(oo, dd, nthreads=4)
        got = wave_emu.run(ta, gi, csr, oo, dd, slots=64, wmax=8, n_warps=2)
        for k in ("poly_id", "t", "xyz", "o"):
            assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("scale,size", [(1.0, 0.05), (30.0, 1.0), (30.0, 40.0), (1000.0, 0.02), (1000.0, 300.0)])
def test_cull_box_never_rejects_a_polygon_the_ray_hits(scale, size):
    """cull_box is not in the reference, so it must be conservative: for rays constructed THROUGH a point of the polygon, from
    origins up to 2 km away, with the FP32 frame point anywhere between the origin and the far side of the model and with
    axis-parallel directions mixed in, the padded box is never rejected.  (hare_box_pad: 1e-3 m + 1e-5 extent + 1e-6 |coordinate|.)"""
    import ctypes as C
    from tests.emu import wave_emu
    L = wave_emu.lib()
    rng = np.random.default_rng(int(scale * 7 + size * 1000))
    n = 200_000
    c = rng.uniform(-scale, scale, (n, 1, 3))
    verts = c + rng.uniform(-size, size, (n, 4, 3))
    vcount = rng.integers(3, 5, n).astype(np.int32)
    verts[vcount == 3, 3] = verts[vcount == 3, 2]
    # a point of the polygon: convex combination of its first three vertices
    w = rng.dirichlet([1, 1, 1], n)
    hitp = (verts[:, :3] * w[:, :, None]).sum(axis=1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    ax = rng.random(n) < 0.2                                        # 20 %: axis-parallel rays (zero direction components)
    k = rng.integers(0, 3, n)
    d[ax] = 0.0; d[ax, k[ax]] = rng.choice([-1.0, 1.0], int(ax.sum()))
    d *= rng.choice([1.0, 0.37, 3.0], (n, 1))                       # reflected directions are not renormalised
    dist = rng.choice([0.0, 1.0, 50.0, 2000.0], n) * rng.random(n)
    o = hitp - d * dist[:, None]
    t_frame = rng.random(n) * (dist / np.linalg.norm(d, axis=1) + rng.choice([0.0, 3.0 * scale], n))
    out = np.zeros(n, np.uint8)
    L.emu_cull_box(verts.ctypes.data_as(C.c_void_p), vcount.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                   t_frame.ctypes.data_as(C.c_void_p), C.c_int64(n), out.ctypes.data_as(C.c_void_p))
    assert out.sum() == 0, f"{int(out.sum())} of {n} hit polygons rejected"
    # ... and the test has teeth: the same rays moved sideways by 20 polygon sizes are mostly rejected
    side = np.cross(d, rng.normal(size=(n, 3))); side /= np.linalg.norm(side, axis=1, keepdims=True) + 1e-300
    o2 = o + side * (20.0 * size + 1.0)
    L.emu_cull_box(verts.ctypes.data_as(C.c_void_p), vcount.ctypes.data_as(C.c_void_p), o2.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p),
                   t_frame.ctypes.data_as(C.c_void_p), C.c_int64(n), out.ctypes.data_as(C.c_void_p))
    assert out.mean() > 0.9


def test_wave_degenerate_rays(hall):
    """Zero, tiny and huge directions, -0.0 components (Q7), origins on vertices / on walls / far outside, rays that miss the
    grid or fault at its boundary: the wavefront state machine gives the oracle's literal IEEE behaviour."""
    from tests.emu import wave_emu
    T, g, ta, gi, csr = hall
    v = ta[0].reshape(-1, 3)
    o = np.array([[15.0, 6.0, 5.0]] * 6 + [v[10], v[500], v[1500], [15.0, 6.0, 0.0], [1e6, 1e6, 1e6], [-1e3, 20.0, 8.0], [15.0, 6.0, 5.0], [15.0, 6.0, 5.0],
                                          [-5.0, 6.0, 5.0], [15.0, 6.0, 5.0]], dtype=np.float64)
    d = np.array([[0, 0, 0], [1e-30, 0, 0], [1e30, 2e30, -1e30], [1e-12, 1e-12, 1], [3, -4, 12], [1e-300, 1e-300, 1e-300],
                  [0.3, 0.4, 0.5], [-0.3, 0.4, 0.5], [0.0, 0.0, 1.0], [0.6, 0.0, 0.8], [-1, -1, -1], [1, 0, 0], [0, -0.0, -1], [1, 1, 0],
                  [-1, 0, 0], [-0.0, 1, 0]], dtype=np.float64)
    assert o.shape == d.shape
    ref = g.Shoot(o, d)
    for slots, wmax in ((64, 8), (32, 4)):
        got = wave_emu.run(ta, gi, csr, o, d, slots=slots, wmax=wmax, n_warps=1)
        for k in ("poly_id", "t", "xyz", "o"):
            assert np.array_equal(got[k], ref[k]), k


def test_ray_bin_keys_stay_in_range_and_group_coherent_rays():
    """ray_bin.cuh: the coherence pre-pass scatters rays through per-bucket counters, so a key outside [0, buckets) would be an
    out-of-bounds atomic on the device -- whatever the ray holds (zero / NaN / Inf / huge components).  And the key does its job:
    rays from one source into a narrow cone share a bucket or neighbouring ones."""
    import ctypes as C
    from tests.emu import wave_emu
    L = wave_emu.lib()
    rng = np.random.default_rng(5)
    n = 200_000
    o = rng.uniform(-50, 80, (n, 3)); d = rng.normal(size=(n, 3))
    special = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1e308, -1e308, 1e-320, 1.0, -1.0])
    o[:5000] = rng.choice(special, (5000, 3)); d[5000:10000] = rng.choice(special, (5000, 3))
    d[10000:11000] = 0.0
    mm = np.array([0.0, 0.0, 0.0, 30.0, 40.0, 17.0])
    keys = np.zeros(n, np.uint32); nb = C.c_uint32()
    L.emu_ray_bin_keys(o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int64(n), mm.ctypes.data_as(C.c_void_p),
                       keys.ctypes.data_as(C.c_void_p), C.byref(nb))
    assert keys.max() < nb.value
    assert len(np.unique(keys[11000:])) > 10_000                      # isotropic rays from everywhere spread over the buckets
    # a 0.5-degree cone from one point: a handful of buckets
    axis = np.array([0.3, -0.5, 0.81]); axis /= np.linalg.norm(axis)
    dd = axis + 0.004 * rng.normal(size=(5000, 3)); oo = np.tile([15.0, 6.0, 5.0], (5000, 1))
    k2 = np.zeros(5000, np.uint32)
    L.emu_ray_bin_keys(oo.ctypes.data_as(C.c_void_p), np.ascontiguousarray(dd).ctypes.data_as(C.c_void_p), C.c_int64(5000), mm.ctypes.data_as(C.c_void_p),
                       k2.ctypes.data_as(C.c_void_p), C.byref(nb))
    assert len(np.unique(k2)) <= 16
    for n2 in (1, 70_000, 3_000_000):          # the key always fits the bucket count chosen for the batch size
        kk = np.zeros(min(n2, n), np.uint32)
        L.emu_ray_bin_keys(o.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.c_int64(len(kk)), mm.ctypes.data_as(C.c_void_p),
                           kk.ctypes.data_as(C.c_void_p), C.byref(nb))
        assert kk.max() < nb.value


@pytest.mark.parametrize("n,tw", [(0, 7), (1, 3), (31, 2), (1000, 40), (70_001, 9), (300_000, 148 * 2), (1_000_003, 64), (3_000_000, 16)])
def test_ray_feed_hands_out_every_ray_exactly_once(n, tw):
    """RayFeed (vg_wave.cuh): the warps claim blocks of the batch from a counter, one block always in reserve -- whatever the order
    in which they come for rays and however many they take per trip, rays 0..N-1 go out once each, and the claims past the end of
    the batch stay bounded."""
    import ctypes as C
    from tests.emu import wave_emu
    L = wave_emu.lib()
    L.emu_ray_feed.restype = C.c_longlong
    for seed in (1, 2, 3):
        counts = np.zeros(max(n, 1), np.uint32)
        covered = L.emu_ray_feed(C.c_int64(n), C.c_int64(tw), C.c_uint64(seed), counts.ctypes.data_as(C.c_void_p))
        assert covered >= n
        assert np.all(counts[:n] == 1)
        assert covered <= n + tw * 128 * 4


def test_shoot_batch_chunk_schedule_covers_the_batch():
    """schedule.hpp: the chunks of a hare_shoot_batch share add up to the share, are never empty (for a non-empty share) nor larger
    than the staging cap, start and end small (what is copied in first and out last overlaps with nothing) and are symmetric."""
    import ctypes as C
    from tests.emu import wave_emu
    L = wave_emu.lib()
    rng = np.random.default_rng(3)
    sizes = np.zeros(4096, np.int64)
    for n in [1, 2, 1000, 262_143, 262_144, 524_289, 1_000_000, 4_000_000, 12_500_000, 100_000_000, 4_294_967_295] + \
             [int(x) for x in rng.integers(1, 600_000_000, 300)]:
        k = L.emu_shoot_schedule(C.c_int64(n), sizes.ctypes.data_as(C.c_void_p), len(sizes))
        v = sizes[:k]
        assert 0 < k <= len(sizes) and int(v.sum()) == n
        assert v.min() > 0 and v.max() <= (1 << 24)
        assert v[0] <= max(1 << 20, n) and v[0] == v[-1] or k == 1
        if n >= (1 << 24):
            assert v[0] <= (1 << 20) and v[-1] <= (1 << 20)        # the un-overlapped copies stay small
        h = (k - 1) // 2
        assert h == 0 or np.abs(v[:h] - v[::-1][:h]).max() <= 1     # (the equal middle chunks differ by a ray)
    assert L.emu_shoot_schedule(C.c_int64(0), sizes.ctypes.data_as(C.c_void_p), len(sizes)) == 1 and sizes[0] == 0
