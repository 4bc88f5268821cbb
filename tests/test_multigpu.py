"""Two-GPU tests of the exchange steps behind the C-ABI (cg_comm_* inside libcitus_gpu.so): the combine in its three
forms (packed words, wide table, row gather + merge), error agreement (a failed rank fails every rank instead of
hanging the others), and the repartition exchange + merge-side join -- all bit-exact against the oracle, once through
the IPC-mapped peer window (the library's own kernels load and store over NVLink) and once with NCCL as the data path.
Needs >= 2 GPUs: skipped on the single-GPU box (run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`)."""
import os
import socket
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _groups_equal(got, want, naggs):
    assert set(got) == set(want), (len(got), len(want))
    for k in want:
        for i in range(naggs):
            assert got[k][i]["sum"] == want[k][i]["sum"] and got[k][i]["count"] == want[k][i]["count"], (k, i)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        from citus_b200 import capi, columnar as cg, distributed as cgd
        from oracle import oracle as orc
        orc.build()
        cg.init(rank)
        cgd.init(rank, world)
        assert cgd.allreduce([rank + 1, -rank], "sum") == [world * (world + 1) // 2, -(world * (world - 1) // 2)]
        assert cgd.allreduce([rank], "max") == [world - 1]
        nshards = 4
        mine = cgd.shards_of_rank(nshards, rank, world)

        def run(cols, quals, group, aggs_fn, oaggs, force_hash=False, expect=None):
            per = 120_000
            rels = {s: cg.Relation.generate(cols, per, seed=99, first_row=s * per, stripe_row_limit=30000, chunk_row_limit=5000)
                    for s in range(nshards)}                      # every rank can build every shard: the oracle needs all
            aggs = aggs_fn()
            d = cg.make_desc(quals, group, aggs)
            kmin, kmax, bounds = None, None, [0] * len(aggs)
            for r in rels.values():
                a, b, bs, _ = cg.relation_bounds(r, d)
                kmin = a if kmin is None else min(kmin, a)
                kmax = b if kmax is None else max(kmax, b)
                bounds = [max(x, y) for x, y in zip(bounds, bs)]
            for a, b in zip(aggs, bounds):
                a.term_abs_bound = b
            d = cg.make_desc(quals, group, aggs)
            if force_hash:
                kmin, kmax = 0, -1
            agg = cg.GpuColumnarAgg(d, rels[0].column_descs(), kmin, kmax, per * nshards)
            for s in mine:
                agg.scan_shard(cg.Shard(rels[s]), want_stats=False)
            cgd.combine_partials(agg, dst=0)
            if rank == 0:
                want = None
                for s in range(nshards):
                    t = orc.Table.attach(rels[s].pages(), rels[s].stripes_bytes(), rels[s].nodes_bytes(), [c[0] for c in cols],
                                         chunk_row_limit=5000)
                    want = t.scan(quals, group, oaggs, into=want)
                got = agg.groups()
                w = want.groups()
                if not group:
                    got = {0: list(got.values())[0]}
                _groups_equal(got, w, len(aggs))
            cgd.barrier()

        for mode in (1, 0, 1):                       # peer window, NCCL, and back (the slots' buffers change hands)
          cg.set_option("peer_window", mode)
          if mode:
              assert cgd.peer_window(), "CUDA IPC mapping of the peer window was refused on this box"
          else:
              assert not cgd.peer_window()
          c2 = [(8, 0, 0, 3000, 0), (8, 0, 0, 100, 0), (8, 0, -10**9, 10**9, 0)]
          c2n = [(8, 0, 0, 3000, 20000), (8, 0, 0, 100, 0), (8, 0, -10**9, 10**9, 50000)]
          oa = [orc.sum_(2), orc.count_star()]
          run(c2, [(1, "<", 50)], [0], lambda: [cg.sum_(2), cg.count_star()], oa)                    # packed words only
          run(c2n, [(1, "<", 50)], [0], lambda: [cg.sum_(2), cg.count_star()], oa)                   # NULLs: wide table reduce
          run(c2n, [(1, "<", 50)], [0], lambda: [cg.sum_(2), cg.count_star()], oa, force_hash=True)  # hash: row gather + merge
          run(c2, [(1, "<", 50)], [], lambda: [cg.sum_(2), cg.count_star()], oa)                     # plain aggregate

          # error agreement: rank 1 reports a failed scan; EVERY rank gets an error, nobody hangs
          d = cg.make_desc([(1, "<", 50)], [0], [cg.sum_(2), cg.count_star()])
          agg = cg.GpuColumnarAgg(d, [(8, 0)] * 3, 0, 2999, 1000)
          try:
              cgd.combine_partials(agg, dst=0, local_status=capi.CG_ECORRUPT if rank == 1 else 0)
              raise AssertionError("combine succeeded although a rank failed")
          except capi.CitusGpuError as e:
              assert e.code == capi.CG_ECORRUPT
          cgd.barrier()

          # repartition exchange + merge-side join against the oracle
          g = torch.Generator(device="cuda")
          P, n = 8, 200_000
          tabs = {}
          for name, seed in (("r", 1), ("s", 2)):
              g.manual_seed(seed * 100 + rank)
              k = torch.randint(0, 40_000, (n,), dtype=torch.int64, device="cuda", generator=g)
              v = torch.randint(-(1 << 40), 1 << 40, (n,), dtype=torch.int64, device="cuda", generator=g)
              tabs[name] = (k, v)
          cgd.repartition_exchange(0, [tabs["r"][0].data_ptr(), tabs["r"][1].data_ptr()], n, P)
          cgd.repartition_exchange(1, [tabs["s"][0].data_ptr(), tabs["s"][1].data_ptr()], n, P)
          cgd.exchange_wait(0)
          cgd.exchange_wait(1)
          r0, r1 = cgd.exchange_result(0, 2, timing=True), cgd.exchange_result(1, 2, timing=True)
          joined, jsum = cg.join_count_sum(r0["cols"][0], r0["cols"][1], r0["nrows"], r1["cols"][0], r1["cols"][1], r1["nrows"])
          # every rank's input, gathered on the host, through the oracle's routing and row-at-a-time join
          allk = {}
          for name in ("r", "s"):
              ks = [torch.zeros(n, dtype=torch.int64) for _ in range(world)]
              vs = [torch.zeros(n, dtype=torch.int64) for _ in range(world)]
              dist.all_gather(ks, tabs[name][0].cpu())
              dist.all_gather(vs, tabs[name][1].cpu())
              allk[name] = (torch.cat(ks).numpy(), torch.cat(vs).numpy())
          mins, maxs = cgd.synthetic_intervals(P)
          wi_r, _ = orc.partition_rows(allk["r"][0], None, 8, "h", mins, maxs)
          wi_s, _ = orc.partition_rows(allk["s"][0], None, 8, "h", mins, maxs)
          sel_r, sel_s = (wi_r % world) == rank, (wi_s % world) == rank
          assert r0["nrows"] == int(sel_r.sum()) and r1["nrows"] == int(sel_s.sum())
          # per local partition, rows by source rank
          mine_parts = [p for p in range(P) if p % world == rank]
          for i, p in enumerate(mine_parts):
              for src in range(world):
                  assert r0["part_counts"][i, src] == int(((wi_r == p) & (np.arange(world * n) // n == src)).sum())
          wj, ws = orc.join_count_sum(allk["r"][0][sel_r], allk["r"][1][sel_r], allk["s"][0][sel_s], allk["s"][1][sel_s])
          assert (joined, jsum) == (wj, ws), (joined, wj)
          # the received rows are exactly the oracle's rows of this rank's partitions (as multisets)
          rk = torch.as_tensor(np.sort(allk["r"][0][sel_r]))
          class V:
              def __init__(s, p, n):
                  s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (p, False), "version": 2}
          got_k = torch.as_tensor(V(r0["cols"][0], r0["nrows"]), device="cuda").cpu()
          got_v = torch.as_tensor(V(r0["cols"][1], r0["nrows"]), device="cuda").cpu()
          assert torch.equal(got_k.sort().values, rk)
          want_pairs = np.sort(allk["r"][0][sel_r] * 1000003 ^ allk["r"][1][sel_r])       # key and payload stay together
          assert np.array_equal(np.sort(got_k.numpy() * 1000003 ^ got_v.numpy()), want_pairs)
          # rows of one local partition from one source rank are contiguous, in (source rank, partition) order
          gi, _ = orc.partition_rows(got_k.numpy(), None, 8, "h", mins, maxs)
          off = 0
          for src in range(world):
              for i, p in enumerate(mine_parts):
                  c = int(r0["part_counts"][i, src])
                  assert (gi[off:off + c] == p).all(), (src, p)
                  off += c
          assert off == r0["nrows"]
        cgd.destroy()
        out.put((rank, "ok"))
    except Exception:          # noqa
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_gpu_combine_and_repartition():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from citus_b200 import build
    build.build()
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
