#!/usr/bin/env python
"""N-GPU parity check (run under torchrun): every rank scans its shards (shard s -> rank s mod N),
the partials are combined with NCCL (dense reduce and row all-gather + merge kernel), and rank 0
compares the combined result with the CPU oracle over all shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from citus_b200 import columnar as cg, distributed as cgd
    cg.init(local)
    cg.use_torch_stream()
    nshards, rows = 8, 200_000
    cols = [(8, 0, 0, 5000, 0), (8, 0, 0, 100, 0), (8, 0, -10**9, 10**9, 20000)]
    ok = True
    for mode in ("dense_add", "dense_minmax_rows", "hash_rows", "plain"):
        aggs = [cg.sum_(2), cg.count_star()] + ([cg.min_(2), cg.max_(2)] if mode == "dense_minmax_rows" else [])
        group = [] if mode == "plain" else [0]
        desc = cg.make_desc([(1, "<", 50)], group, aggs, expected_groups=6000)
        mine = cgd.shards_of_rank(nshards, rank, world)
        rels = {s: cg.Relation.generate(cols, rows, seed=5, first_row=s * rows) for s in range(nshards)}
        kmin, kmax = (0, 4999) if mode.startswith("dense") else (0, -1)
        agg = cg.GpuColumnarAgg(desc, rels[0].column_descs(), kmin, kmax, nshards * rows)
        for s in mine:
            agg.scan_shard(cg.Shard(rels[s]), want_stats=False)
        cgd.combine_partials(agg, dst=0)
        if rank == 0:
            from oracle import oracle as orc
            oaggs = [orc.Agg(a.kind, list(a.factors), a.is_float) for a in aggs]
            want = orc.Result(oaggs)
            for s in range(nshards):
                t = orc.Table.attach(rels[s].pages(), rels[s].stripes_bytes(), rels[s].nodes_bytes(), [8, 8, 8])
                t.scan([(1, "<", 50)], group, oaggs, into=want)
            got, w = agg.groups(), want.groups()
            if not group:
                got = {0: list(got.values())[0]}
            same = set(got) == set(w)
            for k in w:
                for a, spec in enumerate(aggs):
                    g, x = got[k][a], w[k][a]
                    same &= g["count"] == x["count"]
                    if spec.kind == 2:
                        same &= g["sum"] == x["sum"]
                    if spec.kind == 3 and x["count"]:
                        same &= g["minmax"] == x["min"]
                    if spec.kind == 4 and x["count"]:
                        same &= g["minmax"] == x["max"]
            print(f"[{world} GPUs] {mode}: {'bit-exact' if same else 'MISMATCH'} ({len(w)} groups)", flush=True)
            ok &= bool(same)
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
