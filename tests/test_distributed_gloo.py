"""world_size-2 gloo tests (CPU) of the host logic of the multi-GPU path: shard -> rank assignment, the algebra
that makes a plain int64 sum-reduce an exact combine (two-limb 128-bit sums, packed count+sum words with the
rows-added check), and the exchange plan of the repartition (cg_comm_exchange_plan, C, no GPU needed): what
rank a plans to send to rank b is what b plans to receive from a.  The NCCL collectives themselves run in the
-m gpu tests / tools/multigpu_check.py."""
import os
import socket

import numpy as np
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
        from citus_b200 import build
        build.build()
        from citus_b200 import distributed as cgd
        # 1. every shard is owned by exactly one rank (shard s -> rank s mod world)
        mine = cgd.shards_of_rank(32, rank, world)
        owners = [torch.zeros(32, dtype=torch.int64) for _ in range(world)]
        flags = torch.zeros(32, dtype=torch.int64)
        flags[mine] = 1
        dist.all_gather(owners, flags)
        assert torch.stack(owners).sum(0).eq(1).all()
        assert all(s % world == rank for s in mine)

        # 2. what ncclReduce(sum, int64) does to the accumulator words: the two-limb form of a 128-bit sum needs
        #    no carries between the words; packed (sum << C) + count words add up exactly while counts stay < 2^C,
        #    and the rows-added word that travels with them proves it
        rng = np.random.default_rng(100 + rank)
        v = rng.integers(-2**62, 2**62, size=1000)
        lo = torch.from_numpy((v & 0xffffffff).astype(np.int64))
        hi = torch.from_numpy((v >> 32).astype(np.int64))
        words = torch.stack([torch.ones(1000, dtype=torch.int64), lo, hi], dim=1).reshape(-1).contiguous()
        dist.reduce(words, dst=0, op=dist.ReduceOp.SUM)
        allv = [torch.zeros(1000, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allv, torch.from_numpy(v))
        if rank == 0:
            w = words.reshape(1000, 3).numpy()
            for i in range(0, 1000, 97):
                want = sum(int(a[i]) for a in allv)
                got = int(w[i, 1]) + (int(w[i, 2]) << 32)
                assert got == want and w[i, 0] == world
        Cbits = 16
        cnt = rng.integers(0, 300, size=500)
        sums = rng.integers(-10**9, 10**9, size=500) * cnt
        packed = torch.from_numpy(((sums << Cbits) + cnt).astype(np.int64))
        tail = torch.tensor([int(cnt.sum())], dtype=torch.int64)
        buf = torch.cat([packed, tail])
        dist.reduce(buf, dst=0, op=dist.ReduceOp.SUM)
        both = [torch.zeros(1000, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(both, torch.from_numpy(np.concatenate([sums, cnt])))
        if rank == 0:
            w = buf.numpy()
            n = w[:500] & ((1 << Cbits) - 1)
            s = (w[:500] - n) >> Cbits
            assert np.array_equal(n, sum(b.numpy()[500:] for b in both))
            assert np.array_equal(s, sum(b.numpy()[:500] for b in both))
            assert int(n.sum()) == int(w[500])          # drained counts == rows added: no count field overflowed

        # 3. the repartition's exchange plan, computed by the C library on every rank from the same counts matrix
        P = 7
        counts_mine = torch.from_numpy(rng.integers(0, 1000, size=P))
        gathered = [torch.zeros(P, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, counts_mine)
        counts = torch.stack(gathered).numpy()
        pos, send, recv, local = cgd.exchange_plan(P, world, rank, counts)
        owners = [p % world for p in range(P)]
        assert sorted(pos.tolist()) == list(range(P))
        order = np.argsort(pos)                                   # partitions in send-buffer order: destination-major
        assert [owners[p] for p in order] == sorted(owners)
        for r in range(world):
            assert send[r] == sum(int(counts[rank, p]) for p in range(P) if p % world == r)
            assert recv[r] == sum(int(counts[r, p]) for p in range(P) if p % world == rank)
        plans = [torch.zeros(2 * world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(plans, torch.from_numpy(np.concatenate([send, recv])))
        for a in range(world):
            for b in range(world):
                assert int(plans[a][b]) == int(plans[b][world + a])        # a sends to b what b receives from a
        mine_parts = [p for p in range(P) if p % world == rank]
        assert local.shape == (len(mine_parts), world)
        for i, p in enumerate(mine_parts):
            assert local[i].tolist() == counts[:, p].tolist()
        # 4. the peer-window form: every rank scatters straight into the owners' receive buffers.  Each rank computes
        # where its rows land (cg_comm_peer_plan) and "stores" row labels into a model of the buffers, exchanged here
        # over gloo; every buffer must come out exactly filled, in (source rank, local partition) order -- the
        # layout of the send/recv form -- with the column stride cg_comm_peer_plan reports
        pos_begin, total, adj = cgd.peer_plan(P, world, rank, counts)
        assert pos_begin[0] == 0 and pos_begin[world] == P
        sends = []                                                # per destination: (index in that rank's column, label)
        idx_in_send_order = 0
        for q in range(P):                                        # output positions = send order, destination-major
            p = int(order[q])
            d = int(np.searchsorted(pos_begin, q, side="right") - 1)
            assert d == p % world
            for k in range(int(counts[rank, p])):
                sends.append((d, int(adj[d]) + idx_in_send_order, rank * 10**9 + p * 10**6 + k))
                idx_in_send_order += 1
        box = [None] * world
        dist.all_gather_object(box, sends)
        got = {}
        for lst in box:
            for d, i, label in lst:
                if d == rank:
                    assert i not in got, "two rows stored to one place"
                    got[i] = label
        assert total[rank] == recv.sum() and sorted(got) == list(range(int(total[rank])))
        want = [r * 10**9 + p * 10**6 + k for r in range(world) for p in mine_parts for k in range(int(counts[r, p]))]
        assert [got[i] for i in range(len(want))] == want
        for d in range(world):
            assert total[d] == sum(int(counts[r, p]) for r in range(world) for p in range(P) if p % world == d)
        out.put((rank, "ok"))
    except Exception as e:          # noqa
        import traceback
        out.put((rank, traceback.format_exc()))
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
