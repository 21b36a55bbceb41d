"""Deterministic synthetic workloads for BASELINE.json's configs.

Harness code only: mesh and ray generators shared by tests and bench.py.
Nothing here touches the oracle; both the oracle and the CUDA path are fed the
same arrays.
"""
from .rng import splitmix64, uniform01, unit_directions  # noqa: F401
from .meshes import shoebox, hall, Mesh  # noqa: F401
from .rays import rays_from_sources, sample_blocks, sample_rays, source_index  # noqa: F401
