"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard -> rank assignment, the
dense-table reduce and the variable-length row all-gather of the combine step
(citus_b200/distributed.py).  The merge kernel itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from citus_b200 import distributed as cgd
        # 1. every shard is owned by exactly one rank (shard s -> rank s mod world)
        mine = cgd.shards_of_rank(32, rank, world)
        owners = [torch.zeros(32, dtype=torch.int64) for _ in range(world)]
        flags = torch.zeros(32, dtype=torch.int64)
        flags[mine] = 1
        dist.all_gather(owners, flags)
        assert torch.stack(owners).sum(0).eq(1).all()
        assert all(s % world == rank for s in mine)

        # 2. dense accumulator arrays with identical layout reduce by plain int64 sums; the
        #    two-word form of a 128-bit sum needs no carries between the words
        rng = np.random.default_rng(100 + rank)
        v = rng.integers(-2**62, 2**62, size=1000)
        lo = torch.from_numpy((v & 0xffffffff).astype(np.int64))
        hi = torch.from_numpy((v >> 32).astype(np.int64))
        words = torch.stack([torch.ones(1000, dtype=torch.int64), lo, hi], dim=1).reshape(-1).contiguous()
        cgd.reduce_dense_words(words, dst=0)
        allv = [torch.zeros(1000, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allv, torch.from_numpy(v))
        if rank == 0:
            w = words.reshape(1000, 3).numpy()
            for i in range(0, 1000, 97):
                want = sum(int(a[i]) for a in allv)
                got = int(w[i, 1]) + (int(w[i, 2]) << 32)
                assert got == want and w[i, 0] == world

        # 3. all-gather of partial rows of different lengths
        n = 5 + 3 * rank
        nw = 3
        keys = torch.arange(n, dtype=torch.int64) + 1000 * rank
        nulls = torch.zeros(n, dtype=torch.uint8)
        nulls[0] = 1
        w = (torch.arange(n * nw, dtype=torch.int64) + 7 * rank)
        ks, ns, ws, counts = cgd.allgather_rows(keys, nulls, w, nw)
        assert counts == [5 + 3 * r for r in range(world)]
        start = 0
        for r in range(world):
            c = counts[r]
            assert ks[start:start + c].tolist() == (torch.arange(c) + 1000 * r).tolist()
            assert ws[start * nw:(start + c) * nw].tolist() == (torch.arange(c * nw) + 7 * r).tolist()
            assert int(ns[start]) == 1 and int(ns[start + 1:start + c].sum()) == 0
            start += c
        out.put((rank, "ok"))
    except Exception as e:          # noqa
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
