"""Multi-GPU plumbing for the ray-sharded Shoot path (one process per GPU, torch.distributed).

Geometry is replicated on every rank; the ray batch is block-partitioned; the only cross-rank
step is the gather of the per-ray results onto rank 0 (NCCL over NVLink on GPUs, gloo in the CPU
tests).  No collective touches the traversal itself.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block of ray indices owned by `rank`: [n*rank/world, n*(rank+1)/world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_rows(t, dst=0, sizes=None):
    """Gather row-blocks of `t` (same trailing shape on all ranks) to rank `dst`, in rank order.
    `sizes`: rows per rank (defaults to equal blocks).  Returns the concatenated tensor on dst, None elsewhere."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    if sizes is None:
        sizes = [t.shape[0]] * world
    if len(set(sizes)) == 1:
        out = torch.empty((sizes[0] * world,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) if rank == dst else None
        dist.gather(t.contiguous(), list(out.chunk(world)) if rank == dst else None, dst=dst)
        return out
    # ragged blocks: gather() needs equal sizes, so pad to the largest block and trim on dst
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def max_over_ranks(x, device):
    v = torch.tensor([float(x)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def sum_over_ranks(x, device):
    v = torch.tensor([int(x)], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
    return int(v.item())


# ---------------------------------------------------------------------------------------------------------------
# Fused result delivery: the traversal kernel of every rank stores its X_Event rows straight into rank `dst`'s
# buffers over NVLink (peer memory opened through CUDA IPC) -- no gather after the kernel.
# ---------------------------------------------------------------------------------------------------------------
import ctypes as _C

import numpy as _np


class DeviceArray:
    """A raw device allocation of the library (hare_device_alloc) or a peer mapping of one (hare_ipc_open), viewable as a torch
    tensor through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr, device, owner):
        self.ptr, self.shape, self.typestr, self.device, self._owner = int(ptr), tuple(shape), typestr, device, owner
        self.itemsize = _np.dtype(typestr).itemsize
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 2}

    def row_ptr(self, row):
        n = 1
        for s in self.shape[1:]:
            n *= s
        return self.ptr + row * n * self.itemsize

    def torch(self):
        return torch.as_tensor(self, device=torch.device("cuda", self.device))


class PeerResults:
    """X_Event SoA buffers (poly_id, t, xyz, uv) for `n_total` rays living on rank `dst`'s GPU and mapped into every other rank.

    Rank r passes `out_ptrs(first_ray_of_r)` to hare_shoot_batch_device: its kernel then writes rows [lo, hi) of rank dst's arrays
    directly (4 + 8 + 24 + 16 = 52 B per ray over NVLink).  `fence()` orders completion: after it returns on dst's stream every
    rank's kernel has finished, i.e. all rows have landed.
    """
    FIELDS = (("poly_id", (), "<i4"), ("t", (), "<f8"), ("xyz", (3,), "<f8"), ("uv", (2,), "<f8"))

    def __init__(self, n_total, device, dst=0):
        from ._lib import check, lib
        L = lib()
        self.n, self.device, self.dst = n_total, device, dst
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.arrays, self._base, self._opened = {}, {}, []
        dev = torch.device("cuda", device)
        handles = torch.zeros((len(self.FIELDS), 64), dtype=torch.uint8, device=dev)
        if self.rank == dst:
            hs = _np.zeros((len(self.FIELDS), 64), _np.uint8)
            for k, (name, tail, ts) in enumerate(self.FIELDS):
                nbytes = n_total * int(_np.prod(tail, dtype=_np.int64)) * _np.dtype(ts).itemsize
                p = _C.c_void_p()
                check(L.hare_device_alloc(device, nbytes, _C.byref(p)), "hare_device_alloc")
                check(L.hare_ipc_export(device, p, hs[k].ctypes.data_as(_C.c_void_p)), "hare_ipc_export")
                self._base[name] = p.value
            handles.copy_(torch.from_numpy(hs))
        dist.broadcast(handles, src=dst)
        hs = handles.cpu().numpy()
        for k, (name, tail, ts) in enumerate(self.FIELDS):
            if self.rank != dst:
                p = _C.c_void_p()
                check(L.hare_ipc_open(device, _np.ascontiguousarray(hs[k]).ctypes.data_as(_C.c_void_p), _C.byref(p)), "hare_ipc_open")
                self._base[name] = p.value
                self._opened.append(p.value)
            self.arrays[name] = DeviceArray(self._base[name], (n_total,) + tail, ts, device, self)
        self._flag = torch.zeros(1, dtype=torch.int32, device=dev)

    def out_ptrs(self, first_row):
        """(t, xyz, poly_id, uv) device pointers of row `first_row`, in hare_shoot_batch_device's argument order."""
        a = self.arrays
        return a["t"].row_ptr(first_row), a["xyz"].row_ptr(first_row), a["poly_id"].row_ptr(first_row), a["uv"].row_ptr(first_row)

    def fence(self):
        """Stream-ordered completion point: a 4-byte all-reduce behind this rank's kernel (NCCL, current stream)."""
        dist.all_reduce(self._flag)

    def close(self):
        from ._lib import lib
        L = lib()
        for p in self._opened:
            L.hare_ipc_close(self.device, _C.c_void_p(p))
        self._opened = []
        dist.barrier()
        if self.rank == self.dst:
            for p in self._base.values():
                L.hare_device_free(self.device, _C.c_void_p(p))
        self._base = {}
