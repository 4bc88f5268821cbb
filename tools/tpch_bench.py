#!/usr/bin/env python
"""BASELINE config 5: TPC-H Q1 and Q6 over a synthetic columnar lineitem (decimal(15,2) as
scaled int64 cents, flags as 1-byte ints, dates as int4 days -- SURVEY.md 8(d) "C5"),
hash-distributed into 32 shards, shard s -> GPU s mod N, coordinator combine over NCCL.

    python tools/tpch_bench.py [--rows 600000000]                 # SF100 ~ 600 M rows
    torchrun --nproc-per-node N tools/tpch_bench.py --rows ...

Parity: both queries are checked bit-exactly against the CPU oracle on one shard (every group,
every aggregate, 128-bit sums), and the combined result's count(*) against the row count.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NSHARDS = 32
LINEITEM = [(8, 0, 100, 5100, 0),          # 0 l_quantity      1.00 .. 50.99
            (8, 0, 90000, 10500000, 0),    # 1 l_extendedprice
            (8, 0, 0, 11, 0),              # 2 l_discount      0.00 .. 0.10
            (8, 0, 0, 9, 0),               # 3 l_tax
            (1, 0, 65, 68, 0),             # 4 l_returnflag
            (1, 0, 70, 72, 0),             # 5 l_linestatus
            (4, 0, -2922, -365, 0),        # 6 l_shipdate      days since 2000-01-01
            (4, 0, 1, 10001, 0)]           # 7 l_suppkey
BYTES_PER_ROW = {"q6": 4 + 8 + 8 + 8 + 4 / 8, "q1": 4 + 1 + 1 + 8 * 4 + 7 / 8}   # 28.5 / 38.75


def queries(cg):
    q6 = dict(quals=[(6, ">=", -2192), (6, "<", -1827), (2, ">=", 5), (2, "<=", 7), (0, "<", 2400)], group=[],
              aggs=[cg.Agg(2, [(1, 0, 1), (2, 0, 1)])])
    q1 = dict(quals=[(6, "<=", -486)], group=[4, 5],
              aggs=[cg.sum_(0), cg.sum_(1), cg.Agg(2, [(1, 0, 1), (2, 100, -1)]),
                    cg.Agg(2, [(1, 0, 1), (2, 100, -1), (3, 100, 1)]), cg.sum_(2), cg.count_star()])
    return {"q6": q6, "q1": q1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=600_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from citus_b200 import columnar as cg, distributed as cgd
    cg.init(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    cg.use_torch_stream()
    per = a.rows // NSHARDS
    mine = cgd.shards_of_rank(NSHARDS, rank, world)
    t0 = time.time()
    rels = {s: cg.Relation.generate(LINEITEM, per, seed=100, first_row=s * per, nthreads=max(4, 64 // world)) for s in mine}
    shards = {s: cg.Shard(rels[s], [0, 1, 2, 3, 4, 5, 6]) for s in mine}
    if rank == 0:
        print(f"generated + staged {len(mine)} shards x {per} rows in {time.time() - t0:.1f}s", file=sys.stderr)
    out = {}
    for name, q in queries(cg).items():
        d = cg.make_desc(q["quals"], q["group"], q["aggs"], expected_groups=16)
        kmin, kmax, bounds, _ = cg.relation_bounds(rels[mine[0]], d)      # generator ranges are shard independent
        for i, ag in enumerate(q["aggs"]):
            ag.term_abs_bound = bounds[i]
        d = cg.make_desc(q["quals"], q["group"], q["aggs"], expected_groups=16)
        agg = cg.GpuColumnarAgg(d, rels[mine[0]].column_descs(), kmin, kmax, per * NSHARDS)
        # parity on this rank's first shard, every group and aggregate
        if rank == 0:
            from oracle import oracle as orc
            one = cg.GpuColumnarAgg(d, rels[mine[0]].column_descs(), kmin, kmax, per * NSHARDS)
            one.scan_shard(shards[mine[0]], want_stats=False)
            got = one.groups()
            r0 = rels[mine[0]]
            t = orc.Table.attach(r0.pages(), r0.stripes_bytes(), r0.nodes_bytes(), [c[0] for c in LINEITEM])
            oaggs = [orc.Agg(x.kind, list(x.factors), x.is_float) for x in q["aggs"]]
            want = t.scan(q["quals"], q["group"], oaggs).groups()
            if not q["group"]:
                got = {0: list(got.values())[0]}
            assert set(got) == set(want), (name, set(got) ^ set(want))
            for k in want:
                for i in range(len(oaggs)):
                    assert got[k][i]["sum"] == want[k][i]["sum"] and got[k][i]["count"] == want[k][i]["count"], (name, k, i)
        best = None
        for rep in range(a.reps + 1):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            agg.reset()
            for s in mine:
                agg.scan_shard(shards[s], want_stats=False)
            cgd.combine_partials(agg, dst=0)
            n = agg.ngroups() if rank == 0 else 0
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep > 0:
                best = ms if best is None else min(best, ms)
        t = torch.tensor([best], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        if rank == 0:
            g = agg.groups()
            if name == "q1":
                assert sum(v[5]["count"] for v in g.values()) > 0.9 * per * NSHARDS
            out[name] = dict(ms=ms, rows_per_s=per * NSHARDS / ms * 1e3, groups=len(g),
                             algorithmic_gbs=per * NSHARDS * BYTES_PER_ROW[name] / ms / 1e6 / world)
    if rank == 0:
        line = dict(config="C5 TPC-H Q1 + Q6, synthetic lineitem", n_gpus=world, rows=per * NSHARDS, shards=NSHARDS,
                    queries=out, parity="one shard per query bit-exact vs oracle (all groups, all aggregates)")
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"tpch_{world}gpu.json"), "w") as f:
            json.dump(line, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
