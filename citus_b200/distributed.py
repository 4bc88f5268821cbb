"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch;
gloo on CPU for the host-logic tests).

The path shards the way the reference does (one task per shard, planner/
multi_physical_planner.c:2757): shard s is scanned by rank s mod world_size, with no
data-path collective.  The only exchange is the coordinator-side combine of the partial
aggregates (planner/multi_logical_optimizer.c:1807-1885, 2231-2275), done for the
commutative / associative built-ins only:

  * direct-indexed partials whose accumulator words are all additive have the same layout on
    every rank -> one in-place reduce (sum, int64) of the accumulator array to the
    coordinator rank;
  * anything else -> all ranks export their occupied rows (key, NULL flag, accumulator
    words), the rows are all-gathered, and the coordinator rank merges them with the
    combine kernel (cg_partial_merge_rows).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .capi import CG_WORD_ADD


def shards_of_rank(nshards: int, rank: int, world: int):
    """shard s -> rank s mod world"""
    return [s for s in range(nshards) if s % world == rank]


class DeviceWords:
    """zero-copy torch view of device memory owned by libcitus_gpu.so"""

    def __init__(self, ptr: int, nwords: int):
        self.__cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def device_view(ptr: int, nwords: int) -> torch.Tensor:
    return torch.as_tensor(DeviceWords(ptr, nwords), device="cuda")


def reduce_dense_words(words: torch.Tensor, dst: int = 0, group=None):
    """in-place sum of identically laid out accumulator arrays to rank dst"""
    if dist.get_world_size(group) > 1:
        dist.reduce(words, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return words


def allgather_rows(keys: torch.Tensor, key_nulls: torch.Tensor, words: torch.Tensor, nwords: int, group=None):
    """all-gather variable-length partial rows; returns the concatenation over ranks
    (own rows included) as (keys, key_nulls, words)"""
    world = dist.get_world_size(group)
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)

    def padded(t, width):
        out = torch.zeros(cap * width, dtype=t.dtype, device=t.device)
        out[: t.numel()] = t.reshape(-1)
        return out

    gk = [torch.empty(cap, dtype=keys.dtype, device=keys.device) for _ in range(world)]
    gn = [torch.empty(cap, dtype=key_nulls.dtype, device=keys.device) for _ in range(world)]
    gw = [torch.empty(cap * nwords, dtype=words.dtype, device=keys.device) for _ in range(world)]
    dist.all_gather(gk, padded(keys, 1), group=group)
    dist.all_gather(gn, padded(key_nulls, 1), group=group)
    dist.all_gather(gw, padded(words, nwords), group=group)
    ks = torch.cat([gk[r][: counts[r]] for r in range(world)])
    ns = torch.cat([gn[r][: counts[r]] for r in range(world)])
    ws = torch.cat([gw[r][: counts[r] * nwords] for r in range(world)])
    return ks, ns, ws, counts


def library_uses_torch_stream() -> bool:
    from . import columnar
    return getattr(columnar, "_on_torch_stream", None) == torch.cuda.current_stream().cuda_stream


def combine_partials(agg, dst: int = 0, group=None):
    """coord_combine over the ranks: after the call rank `dst`'s partial holds the combined
    aggregate.  `agg` is a columnar.GpuColumnarAgg."""
    if not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    rank = dist.get_rank(group)
    nw, ops, dense, cap = agg.layout()
    if (dense or agg.desc.ngroup_cols == 0) and all(o == CG_WORD_ADD for o in ops):
        if library_uses_torch_stream():
            # one queue: table maintenance, the collective and whatever reads the result are stream ordered, so the
            # host only waits once, after the collective is on its way (every rank: a worker error fails the query)
            ptr, total, stride = agg.dense_words_enqueue()
            reduce_dense_words(device_view(ptr, total), dst, group)
            agg.check()
            return
        ptr, total, stride = agg.dense_words()          # drains + verifies packed words, syncs the library's stream
        reduce_dense_words(device_view(ptr, total), dst, group)
        # the collective runs on torch's stream; the library is on its own stream
        torch.cuda.current_stream().synchronize()
        return
    n = agg.ngroups()
    keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    kn = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    words = torch.empty(max(n, 1) * nw, dtype=torch.int64, device="cuda")
    n = agg.export_device(keys.data_ptr(), kn.data_ptr(), words.data_ptr(), n)
    ks, ns, ws, counts = allgather_rows(keys[:n], kn[:n], words[: n * nw], nw, group)
    torch.cuda.current_stream().synchronize()           # gathered rows complete before the merge kernel reads them
    if rank == dst:
        # own rows are already in the table: merge everybody else's
        start = 0
        for r in range(world):
            if r != rank and counts[r]:
                agg.merge_rows(ks[start:start + counts[r]].data_ptr(), ns[start:start + counts[r]].data_ptr(),
                               ws[start * nw:(start + counts[r]) * nw].data_ptr(), counts[r])
            start += counts[r]


def repartition_all_to_all(keys: torch.Tensor, key_nulls, payload: list, partition_count: int, group=None,
                           key_len: int = 8):
    """Hash repartition of this rank's rows (the MAP_TASK + MAP_OUTPUT_FETCH_TASK pair of a dual
    repartition join, executor/partitioned_intermediate_results.c:115-298 and
    executor/intermediate_results.c:890): rows are routed by worker_partition_query_result's rule
    (hashint8 -> binary search of the synthetic token ranges, planner/multi_physical_planner.c:
    4667-4701), scattered into partition-contiguous order on the GPU, and partition p is sent to
    rank p mod world with one NCCL all-to-all per column (instead of P files + COPY over libpq).

    Returns (columns received [key, *payload], per-partition row counts of the received rows as a
    [partitions owned by this rank][world] tensor).  Partition p's rows from all source ranks are
    NOT merged into one run: the consumer (join build/probe) only needs them partition-complete.
    """
    from . import columnar as cg
    import numpy as np

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = keys.shape[0]
    P = partition_count
    mins, maxs = synthetic_intervals(P)
    idx = torch.empty(n, dtype=torch.int32, device=keys.device)
    counts = torch.empty(P, dtype=torch.int64, device=keys.device)
    cg.worker_partition_query_result(keys.data_ptr(), key_nulls.data_ptr() if key_nulls is not None else None,
                                     n, key_len, "hash", mins, maxs, idx.data_ptr(), counts.data_ptr())
    cols = [keys] + list(payload)
    # destination-major order: partitions owned by rank 0 first, then rank 1, ...  (p -> p mod world)
    order = [p for r in range(world) for p in range(P) if p % world == r]
    position = np.empty(P, np.int32)
    position[np.asarray(order)] = np.arange(P, dtype=np.int32)
    outs = [torch.empty_like(c) for c in cols]
    offs = cg.partition_scatter(idx.data_ptr(), n, P, [c.data_ptr() for c in cols], [o.data_ptr() for o in outs],
                                order=position)
    sizes_by_part = np.diff(offs)                          # rows per (reordered) partition
    per_rank = [len([p for p in range(P) if p % world == r]) for r in range(world)]
    bounds = np.cumsum([0] + per_rank)
    send_rows = [int(sizes_by_part[bounds[r]:bounds[r + 1]].sum()) for r in range(world)]
    if world == 1:
        return outs, torch.from_numpy(sizes_by_part.reshape(1, -1).T.copy())
    # exchange per-partition counts, then the rows
    send_counts = torch.from_numpy(sizes_by_part.astype(np.int64)).to(keys.device)
    mine = per_rank[rank]
    recv_counts = torch.empty(world * mine, dtype=torch.int64, device=keys.device)
    dist.all_to_all_single(recv_counts, send_counts, output_split_sizes=[mine] * world,
                           input_split_sizes=per_rank, group=group)
    recv_counts = recv_counts.reshape(world, mine)
    recv_rows = [int(x) for x in recv_counts.sum(1).tolist()]
    received = []
    for o in outs:
        buf = torch.empty(sum(recv_rows), dtype=o.dtype, device=o.device)
        dist.all_to_all_single(buf, o, output_split_sizes=recv_rows, input_split_sizes=send_rows, group=group)
        received.append(buf)
    return received, recv_counts.T.contiguous()


def synthetic_intervals(partition_count: int):
    """GenerateSyntheticShardIntervalArray (planner/multi_physical_planner.c:4667-4701): uniform
    int4 token ranges, the last one widened to INT32_MAX"""
    import numpy as np
    inc = (1 << 32) // partition_count
    mins = np.array([-(1 << 31) + i * inc for i in range(partition_count)], dtype=np.int64)
    maxs = mins + inc - 1
    maxs[-1] = (1 << 31) - 1
    return mins.astype(np.int32), maxs.astype(np.int32)
