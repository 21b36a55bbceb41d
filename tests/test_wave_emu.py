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
    ref = g.reflect_chain(o, d, 12, nthreads=4)
    got = wave_emu.run(ta, gi, csr, o, d, chain=True, order=12, slots=slots, wmax=wmax, n_warps=warps)
    for k in ("ev_poly_id", "ev_t", "o", "d", "nshots"):
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
