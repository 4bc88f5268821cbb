"""Multi-GPU: one process per GPU.  A binding of the library's exchange entry points (cg_comm_*, NCCL over
NVLink / NVSwitch inside libcitus_gpu.so) plus the bootstrap that ships the communicator id to every rank.

The path shards the way the reference does (one task per shard, planner/multi_physical_planner.c:2757):
shard s is scanned by rank s mod world_size with no data-path collective.  The exchange steps are
  * cg_comm_combine: the coordinator-side combine of the partial aggregates
    (planner/multi_logical_optimizer.c:1807-1885, 2231-2275 over executor/adaptive_executor.c:3964-4189)
  * cg_comm_repartition_exchange: the map-output fetch of a repartition join
    (executor/partitioned_intermediate_results.c:115-298, executor/intermediate_results.c:789-1045)
Both live in C (citus_b200/csrc/cg_comm.cu).  torch.distributed is used for one thing only: handing rank 0's
128-byte NCCL id to the other ranks at start-up (gloo; in PostgreSQL the coordinator's connections would carry it).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib

_world = 1
_rank = 0


def shards_of_rank(nshards: int, rank: int, world: int):
    """shard s -> rank s mod world"""
    return [s for s in range(nshards) if s % world == rank]


def init(rank: int, world: int, bootstrap_group=None):
    """cg_comm_init on every rank.  world > 1 needs torch.distributed initialised (any backend) to ship the id."""
    global _world, _rank
    ident = (C.c_uint8 * capi.CG_COMM_ID_BYTES)()
    if world > 1:
        import torch
        import torch.distributed as dist
        if rank == 0:
            check(lib().cg_comm_unique_id(ident))
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0, group=bootstrap_group)
        ident = (C.c_uint8 * capi.CG_COMM_ID_BYTES).from_buffer_copy(box[0])
    check(lib().cg_comm_init(ident, rank, world))
    _world, _rank = world, rank


def destroy():
    check(lib().cg_comm_destroy())


def world_size() -> int:
    return _world


def rank() -> int:
    return _rank


def barrier():
    check(lib().cg_comm_barrier())


def peer_window() -> bool:
    """True when the combine and the repartition exchange run through the IPC-mapped peer window (NVLink loads and
    stores by the library's kernels); False when NCCL carries the data"""
    return bool(lib().cg_comm_peer_window())


def allreduce(values, op="max"):
    """host-side agreement over ranks: list of python ints -> list of python ints"""
    a = np.asarray(values, np.int64).copy()
    check(lib().cg_comm_allreduce_i64(a.ctypes.data, a.shape[0], {"sum": capi.CG_COMM_SUM, "min": capi.CG_COMM_MIN,
                                                                   "max": capi.CG_COMM_MAX}[op]))
    return [int(x) for x in a]


def combine_partials(agg, dst: int = 0, local_status: int = 0):
    """coord_combine over the ranks (cg_comm_combine): after the call rank `dst`'s partial holds the combined
    aggregate.  Every rank calls it, also one whose scans failed (local_status = its error code)."""
    check(lib().cg_comm_combine(agg.h, dst, local_status))


def synthetic_intervals(partition_count: int):
    """GenerateSyntheticShardIntervalArray (planner/multi_physical_planner.c:4667-4701): uniform
    int4 token ranges, the last one widened to INT32_MAX"""
    inc = (1 << 32) // partition_count
    mins = np.array([-(1 << 31) + i * inc for i in range(partition_count)], dtype=np.int64)
    maxs = mins + inc - 1
    maxs[-1] = (1 << 31) - 1
    return mins.astype(np.int32), maxs.astype(np.int32)


def repartition_exchange(slot: int, col_ptrs, n: int, partition_count: int, key_nulls_ptr=None, key_len: int = 8):
    """worker_partition_query_result + fetch over NCCL for this rank's rows (column 0 = key): enqueues routing,
    scatter and the grouped all-to-all; returns the rows this rank will hold.  exchange_result(slot) gives the
    device columns; exchange_wait(slot) orders the library's compute stream behind the exchange."""
    mins, maxs = synthetic_intervals(partition_count)
    cols = (C.c_void_p * len(col_ptrs))(*col_ptrs)
    got = C.c_int64()
    check(lib().cg_comm_repartition_exchange(slot, cols, key_nulls_ptr, n, len(col_ptrs), key_len, partition_count,
                                             mins.ctypes.data, maxs.ctypes.data, C.byref(got)))
    return got.value


def exchange_wait(slot: int):
    check(lib().cg_comm_exchange_wait(slot))


def exchange_result(slot: int, ncols: int, timing=False):
    """dict: cols (device pointers), nrows, part_counts [nlocal][world], sent_bytes, exchange_ms"""
    ptrs = (C.c_void_p * 8)()
    nrows, nlocal, sent, ms = C.c_int64(), C.c_int32(), C.c_uint64(), C.c_double()
    check(lib().cg_comm_exchange_result(slot, ptrs, C.byref(nrows), None, C.byref(nlocal), C.byref(sent), None))
    counts = np.zeros((max(nlocal.value, 1), _world), np.int64)
    check(lib().cg_comm_exchange_result(slot, None, None, counts.ctypes.data, None, None, C.byref(ms) if timing else None))
    return dict(cols=[ptrs[i] for i in range(ncols)], nrows=nrows.value, part_counts=counts[:nlocal.value],
                sent_bytes=sent.value, exchange_ms=ms.value if timing else None)


def peer_plan(partition_count: int, world: int, rank_: int, counts):
    """cg_comm_peer_plan: (pos_begin[world + 1], total[world], adj[world]) -- where this rank's rows land when the scatter
    stores straight into the owners' receive buffers"""
    counts = np.ascontiguousarray(counts, np.int64)
    pos_begin, total, adj = np.zeros(world + 1, np.int32), np.zeros(world, np.int64), np.zeros(world, np.int64)
    check(lib().cg_comm_peer_plan(partition_count, world, rank_, counts.ctypes.data, pos_begin.ctypes.data, total.ctypes.data,
                                  adj.ctypes.data))
    return pos_begin, total, adj


def exchange_plan(partition_count: int, world: int, rank_: int, counts=None):
    """cg_comm_exchange_plan: (position[P], send_rows[world], recv_rows[world], local_part_counts[nlocal][world])"""
    pos = np.zeros(partition_count, np.int32)
    nlocal = C.c_int32()
    if counts is None:
        check(lib().cg_comm_exchange_plan(partition_count, world, rank_, None, pos.ctypes.data, None, None, None, C.byref(nlocal)))
        return pos, None, None, None
    counts = np.ascontiguousarray(counts, np.int64)
    send, recv = np.zeros(world, np.int64), np.zeros(world, np.int64)
    check(lib().cg_comm_exchange_plan(partition_count, world, rank_, None, pos.ctypes.data, None, None, None, C.byref(nlocal)))
    local = np.zeros((max(nlocal.value, 1), world), np.int64)
    check(lib().cg_comm_exchange_plan(partition_count, world, rank_, counts.ctypes.data, pos.ctypes.data, send.ctypes.data,
                                      recv.ctypes.data, local.ctypes.data, C.byref(nlocal)))
    return pos, send, recv, local[:nlocal.value]
