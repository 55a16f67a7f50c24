"""Multi-GPU: node-partitioned feature table with the halo exchange fused into the gather.

One process per GPU (torch.distributed for bootstrap / barriers only).  Rank r owns the feature rows
of global nodes [r*R, (r+1)*R); every rank maps every shard through CUDA IPC (NVLink / NVSwitch peer
memory), and the gather kernels (gs_gather_mean_sharded / gs_gather_rows_sharded) resolve each id to
`base[id // R] + (id % R) * pitch` - remote rows are pulled by the consuming kernel itself, so no
staging buffer and no collective sits on the data path.  The reference is single-device
(supervised_train.py:59); this is new design.

Seeds are routed to their owner (owner-computes), so hop-0 self rows are always local; how many of the
hop-1/hop-2 rows are remote is decided by the partition (relabel nodes by community first:
`locality_order`).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import ShardedTable, check, lib
from .ops import pad_cols


def rows_per_shard(n_nodes, world):
    return (int(n_nodes) + world - 1) // world


def owner_of(ids, n_nodes, world):
    """Owner rank of each global id (numpy or torch); the dummy id n_nodes maps to -1 (every rank has a zero row)."""
    R = rows_per_shard(n_nodes, world)
    o = ids // R
    if torch.is_tensor(ids):
        return torch.where((ids < 0) | (ids >= n_nodes), torch.full_like(o, -1), o)
    return np.where((ids < 0) | (ids >= n_nodes), -1, o)


def locality_order(comm):
    """Permutation `order` (new id -> old id) that makes communities contiguous, so that contiguous
    equal ranges align with communities; returns (order, inverse) with inverse[old] = new."""
    order = np.argsort(comm, kind="stable")
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    return order, inv


def relabel_graph(indptr, indices, order, inv):
    """CSR of the graph with node ids renamed old -> inv[old] (rows reordered accordingly)."""
    indptr = np.asarray(indptr)
    deg = np.diff(indptr)[order]
    new_ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    src_start = indptr[:-1][order]
    ent = np.arange(new_ptr[-1]) - np.repeat(new_ptr[:-1], deg) + np.repeat(src_start, deg)
    return new_ptr, inv[np.asarray(indices)[ent]].astype(np.int32)


def route_seeds(seeds, n_nodes, group=None):
    """Owner-computes routing: every rank passes the seeds it was handed; returns the seeds this rank owns
    (all_to_all of variable-length id lists; works on gloo and nccl)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    seeds = seeds.reshape(-1)
    own = owner_of(seeds, n_nodes, world).clamp(min=0)
    order = torch.argsort(own, stable=True)
    send = seeds[order].contiguous()
    counts = torch.bincount(own, minlength=world)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    out = torch.empty(int(recv_counts.sum().item()), dtype=seeds.dtype, device=seeds.device)
    dist.all_to_all_single(out, send, output_split_sizes=recv_counts.tolist(), input_split_sizes=counts.tolist(),
                           group=group)
    return out


class _CudaView(object):
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 3, "strides": None}


class ShardedFeatures(object):
    """This rank's shard of a node-partitioned [N+1, F] fp32 feature table plus peer mappings of all others.

    local_rows: float32 [n_local, F] rows of global nodes [rank*R, rank*R + n_local) (numpy or tensor).
    The shard buffer is [R + 1, pitch] with a zero row at index R (the dummy row, reference
    supervised_train.py:133-135, kept local on every rank).
    """

    def __init__(self, local_rows, n_nodes, group=None, device=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_nodes = int(n_nodes)
        self.R = rows_per_shard(n_nodes, self.world)
        local_rows = torch.as_tensor(local_rows, dtype=torch.float32)
        F = local_rows.shape[1]
        lo = self.rank * self.R
        n_local = max(0, min(self.R, self.n_nodes - lo))
        if local_rows.shape[0] != n_local:
            raise ValueError("rank %d must pass %d rows (got %d)" % (self.rank, n_local, local_rows.shape[0]))
        self.shape = (self.n_nodes + 1, F)
        self.pitch = pad_cols(F)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nbytes = (self.R + 1) * self.pitch * 4
        p = ctypes.c_void_p()
        check(lib().gs_shard_alloc(nbytes, ctypes.byref(p)))
        self._own_ptr = p.value
        self.local = torch.as_tensor(_CudaView(p.value, (self.R + 1, self.pitch)), device=self.device)
        self.local.zero_()
        self.local[:n_local, :F] = local_rows.to(self.device)
        torch.cuda.synchronize()
        # exchange IPC handles
        handle = ctypes.create_string_buffer(64)
        check(lib().gs_ipc_export(p, handle))
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        else:
            handles[0] = bytes(handle.raw)
        self._peer_ptrs = []
        self._table = ShardedTable()
        for r in range(self.world):
            if r == self.rank:
                self._table.base[r] = p.value
            else:
                q = ctypes.c_void_p()
                check(lib().gs_ipc_import(handles[r], ctypes.byref(q)))
                self._peer_ptrs.append(q.value)
                self._table.base[r] = q.value
        self._table.n_shards = self.world
        self._table.my_shard = self.rank
        self._table.rows_per_shard = self.R
        self._table.n_global_rows = self.n_nodes + 1
        if self.world > 1:
            dist.barrier(group=group)

    def c_table(self):
        return ctypes.byref(self._table)

    def remote_fraction(self, ids):
        """Fraction of the given global ids whose feature row lives on another rank."""
        own = owner_of(ids.reshape(-1), self.n_nodes, self.world)
        return float(((own >= 0) & (own != self.rank)).float().mean().item())

    def close(self):
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized():
            dist.barrier(group=self.group)          # nobody may still be reading our shard
        for q in self._peer_ptrs:
            check(lib().gs_ipc_close(ctypes.c_void_p(q)))
        self._peer_ptrs = []
        if self._own_ptr:
            self.local = None
            check(lib().gs_shard_free(ctypes.c_void_p(self._own_ptr)))
            self._own_ptr = None
