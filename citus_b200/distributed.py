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
        ptr, total, stride = agg.dense_words()
        reduce_dense_words(device_view(ptr, total), dst, group)
        return
    n = agg.ngroups()
    keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    kn = torch.empty(max(n, 1), dtype=torch.uint8, device="cuda")
    words = torch.empty(max(n, 1) * nw, dtype=torch.int64, device="cuda")
    n = agg.export_device(keys.data_ptr(), kn.data_ptr(), words.data_ptr(), n)
    ks, ns, ws, counts = allgather_rows(keys[:n], kn[:n], words[: n * nw], nw, group)
    if rank == dst:
        # own rows are already in the table: merge everybody else's
        start = 0
        for r in range(world):
            if r != rank and counts[r]:
                agg.merge_rows(ks[start:start + counts[r]].data_ptr(), ns[start:start + counts[r]].data_ptr(),
                               ws[start * nw:(start + counts[r]) * nw].data_ptr(), counts[r])
            start += counts[r]
